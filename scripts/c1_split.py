"""C1 (N = 100, n = 100, K = 1) on the device-resident loop: wall clock per iteration as a function of the walk
length S -> cost of the update (intercept) and of a walk step (slope)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from bayesianinference_b200 import engine, configs as cfg
engine.init()
c = cfg.c1_gaussian()
gp = engine.Problem.from_config(c)
iters = 800
res = []
for S in (1, 25, 50, 100, 200, 400):
    best = 1e9
    for rep in range(3):
        o = engine.default_options(pool_size=100, batch_k=1, mc_steps=S, max_iter=iters, min_iter=iters, seed=7 + rep)
        run = engine.RunGroup(gp, o)
        t0 = time.perf_counter(); run.advance(0); dt = time.perf_counter() - t0
        acc = np.nanmean(run.fetch(0)["acc"])
        path = run.walk_path()
        run.close()
        best = min(best, dt)
    res.append((S, best / iters * 1e6, acc))
    print(f"S = {S:4d}: {best / iters * 1e6:8.2f} us per iteration ({path}), mean acceptance {acc:.3f}")
S = np.array([r[0] for r in res], float); t = np.array([r[1] for r in res])
slope, icpt = np.polyfit(S, t, 1)
print(f"update ~ {icpt:.1f} us per iteration, walk ~ {slope * 1e3:.0f} ns per step")
