"""C3 (softmax classification, N = 1e6) streaming-likelihood kernel: ms per launch and fraction of the measured fp64 peak
(81 flop per datum-eval).  BINEST_LIB=<path> selects an A/B build of the library (csrc/Makefile)."""
import sys; sys.path.insert(0, '.')
from bayesianinference_b200 import engine, configs as cfg
engine.init()
peak = engine.fp64_peak()
c3 = cfg.c3_logistic()
g3 = engine.Problem.from_config(c3)
for P in (64, 512, 2048):
    k, t = g3.bench_loglike(P, 10, 3, True)
    tf = 81.0 * 1e6 * P / (k * 1e-3) / 1e12
    print(f"C3 P={P}: kernel {k:.4f} ms, total {t:.4f} ms, {tf:.2f} TFLOP/s = {tf / peak:.3f} of the measured fp64 peak ({peak:.2f})")
