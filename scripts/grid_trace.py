"""Timeline of the persistent walk kernel (BINEST_GRID_TRACE=1): one launch, CTAs 0 and G/2, first 16 steps."""
import os, sys
os.environ["BINEST_GRID_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianinference_b200 import configs as cfg, engine
engine.init()
c = cfg.c2_polyreg()
gp = engine.Problem.from_config(c)
o = engine.default_options(pool_size=1024, batch_k=256, mc_steps=200, max_iter=10**9, min_iter=10**9, seed=3)
run = engine.RunGroup(gp, o)
run.advance(1)
