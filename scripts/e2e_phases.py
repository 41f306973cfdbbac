import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from bayesianinference_b200 import engine, configs as cfg
engine.init()
c = cfg.c2_polyreg()
gp0 = engine.Problem.from_config(c)
sp = gp0.sample_prior(1024, 5).reshape(1, 1024, 5)
o = engine.default_options(pool_size=1024, batch_k=256, mc_steps=200, max_iter=10**9, min_iter=10**9, seed=11)
for it in range(4):
    t = [time.perf_counter()]
    p2 = engine.Problem(c.op, c.inputs, c.outputs, c.iparam, c.kinds, c.lo, c.hi, c.p0, c.p1); t.append(time.perf_counter())
    r2 = engine.RunGroup(p2, o, sp); t.append(time.perf_counter())
    r2.advance(1); t.append(time.perf_counter())
    res = r2.fetch(0); t.append(time.perf_counter())
    r2.close(); p2.close(); t.append(time.perf_counter())
    print(it, ["%.1f ms" % (1e3 * (b - a)) for a, b in zip(t, t[1:])])
