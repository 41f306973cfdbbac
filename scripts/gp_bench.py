import sys, time; sys.path.insert(0, '.')
import numpy as np
from bayesianinference_b200 import engine, configs as cfg
engine.init()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
c = cfg.c5_gp(N=N)
gp = engine.Problem.from_config(c)
th = gp.sample_prior(B, 905)
gp.loglike(th)  # warm-up with the full batch: the workspace (B matrices) is allocated here, not in the timed call
import torch
REPS = 3
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(REPS):
    out = gp.loglike(th)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / REPS
flop = B * (N**3 / 3 + N * N * 26 / 2)
print(f"GP N={N} B={B}: {dt*1e3:.1f} ms, {B/dt:.1f} evals/s, {flop/dt/1e12:.2f} TFLOP/s (Cholesky N^3/3 + fill)", out[:3])
if len(sys.argv) > 3:  # predictFromGaussianProcess: Q prediction inputs appended to the same sweep
    Q = int(sys.argv[3])
    xs = np.linspace(0.0, 10.0, Q)
    gp.gp_predict(th, xs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(REPS):
        m, s = gp.gp_predict(th, xs)
    torch.cuda.synchronize(); dt2 = (time.perf_counter() - t0) / REPS
    flop2 = flop + B * (N * N * Q)  # + the Q x N triangular solve (N^2 Q) riding on the sweep
    print(f"GP predict N={N} B={B} Q={Q}: {dt2*1e3:.1f} ms ({dt2/dt:.2f}x the likelihood), {flop2/dt2/1e12:.2f} TFLOP/s", m[0, :2], s[0, :2])
