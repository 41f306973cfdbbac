import sys; sys.path.insert(0, '.')
from bayesianinference_b200 import engine, configs as cfg
engine.init()
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
R = int(sys.argv[3]) if len(sys.argv) > 3 else 1
c = cfg.ALL[name]()
gp = engine.Problem.from_config(c)
o = engine.default_options(pool_size=c.pool_size, batch_k=K, mc_steps=200, max_iter=10**9, min_iter=10**9, seed=3, n_runs=R)
run = engine.RunGroup(gp, o)
for _ in range(4):
    run.advance(1)
print(run.timing())
