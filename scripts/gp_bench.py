import sys, time; sys.path.insert(0, '.')
import numpy as np
from bayesianinference_b200 import engine, configs as cfg
engine.init()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
c = cfg.c5_gp(N=N)
gp = engine.Problem.from_config(c)
th = gp.sample_prior(B, 905)
gp.loglike(th[:2])
import torch
torch.cuda.synchronize(); t0 = time.perf_counter()
out = gp.loglike(th)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
flop = B * (N**3 / 3 + N * N * 26 / 2)
print(f"GP N={N} B={B}: {dt*1e3:.1f} ms, {B/dt:.1f} evals/s, {flop/dt/1e12:.2f} TFLOP/s (Cholesky N^3/3 + fill)", out[:3])
