import sys; sys.path.insert(0, '.')
from bayesianinference_b200 import engine, configs as cfg
engine.init()
c = cfg.c1_gaussian()
gp = engine.Problem.from_config(c)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
o = engine.default_options(pool_size=100, batch_k=1, mc_steps=S, max_iter=600, min_iter=600, seed=7)
run = engine.RunGroup(gp, o)
run.advance(0)
print(run.sizes(0))
