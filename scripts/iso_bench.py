import sys; sys.path.insert(0,'.')
from bayesianinference_b200 import engine, configs as cfg
engine.init()
pk = engine.fp64_peak()
for name, fac, flop, Ps in (("C2", cfg.c2_polyreg, 9, (256, 1024)), ("C3", cfg.c3_logistic, 81, (512, 2048)), ("C4", cfg.c4_gbm, 4, (4096,))):
    c = fac(); gp = engine.Problem.from_config(c)
    rows = c.inputs.shape[0] - (1 if c.op == cfg.OP_GBM else 0)
    for P in Ps:
        k, t = gp.bench_loglike(P, 20, 3, True)
        print(f"{name} P={P}: kernel {k*1e3:.1f} us = {flop*rows*P/k/1e9:.2f} TF = {flop*rows*P/k/1e9/pk:.3f} of fp64 peak {pk:.1f}")
