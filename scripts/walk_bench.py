"""Device time of one S-step walk block (binest_run_timing) for the walk paths of engine.cu, selected by env:
  default            persistent grid-resident kernel, two alternating walker sets (walk_grid.cuh)
  BINEST_GRID_SETS=1 persistent kernel, one set (barriers + chain logic on the critical path)
  BINEST_NO_GRID=1   stepped CUDA graph [walk_step, loglike_stream] x S
usage: python scripts/walk_bench.py [C2] [runs_per_gpu] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bayesianinference_b200 import configs as cfg  # noqa: E402
from bayesianinference_b200 import engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
n_runs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 8
fac = {"C1": (cfg.c1_gaussian, 32, 3), "C2": (cfg.c2_polyreg, 256, 9), "C3": (cfg.c3_logistic, 512, 81),
       "C4": (cfg.c4_gbm, 64, 4)}[name]
c = fac[0]()
engine.init()
peak = engine.fp64_peak()
gp = engine.Problem.from_config(c)
opts = engine.default_options(pool_size=c.pool_size, batch_k=fac[1], mc_steps=200, max_iter=10**9, min_iter=10**9,
                              seed=2026, n_runs=n_runs)
run = engine.RunGroup(gp, opts)
run.advance(3)
t0 = run.timing()
run.advance(iters)
t1 = run.timing()
ms = (t1["walk_ms"] - t0["walk_ms"]) / (t1["walk_graphs"] - t0["walk_graphs"])
rows = c.inputs.shape[0] - (1 if c.op == cfg.OP_GBM else 0)
flop = fac[2] * rows * n_runs * fac[1] * 200
s = run.fetch(0)
print(f"{name} runs={n_runs} env={ {k: v for k, v in os.environ.items() if k.startswith('BINEST')} } "
      f"walk block {ms:.3f} ms = {ms / 200 * 1e3:.2f} us/step, {flop / ms / 1e9:.2f} TFLOP/s = {flop / ms / 1e9 / peak:.3f} of "
      f"fp64 peak {peak:.1f}; M={s['M']} crude_logZ={s['crude_logZ']:.6f} logL[-1]={s['logL'][-1]:.9f}")
chk = gp.loglike(s["points"])
rel = np.abs(chk - s["logL"]) / np.abs(chk)
print(f"  stored logL vs operator at the stored point: max rel err {rel.max():.3e} (argmax {rel.argmax()} of {rel.size}, "
      f"n_deleted={s['n_deleted']}); logL quantiles {np.percentile(s['logL'], [0, 25, 50, 75, 100])}")
print("  acc mean", np.nanmean(s["acc"]), "first dead logL", s["logL"][:3], "live min", s["logL"][s["n_deleted"]])
