import sys; sys.path.insert(0, '.')
from bayesianinference_b200 import engine, configs as cfg
engine.init()
c = cfg.c1_gaussian()
gp = engine.Problem.from_config(c)
for seed in (1, 2, 3):
    o = engine.default_options(pool_size=100, batch_k=1, mc_steps=200, max_iter=10**6, seed=seed)
    run = engine.RunGroup(gp, o)
    run.advance(0)
    print(run.walk_path(), run.sizes(0), run.timing())
    run.close()
