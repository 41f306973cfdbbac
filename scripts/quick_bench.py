import sys, time; sys.path.insert(0,'.')
import numpy as np
from bayesianinference_b200 import engine, configs as cfg
engine.init()
print("fp64 peak TF:", engine.fp64_peak())
c = cfg.c2_polyreg()
gp = engine.Problem.from_config(c)
for P in (1, 32, 256, 1024):
    for fl in (False, True):
        print("P", P, "flush", fl, gp.bench_loglike(P, 20, 3, fl))
c3 = cfg.c3_logistic()
g3 = engine.Problem.from_config(c3)
for P in (256, 2048):
    print("C3 P", P, g3.bench_loglike(P, 5, 2, True))
c4 = cfg.c4_gbm()
g4 = engine.Problem.from_config(c4)
for P in (256, 4096):
    print("C4 P", P, g4.bench_loglike(P, 20, 3, True))
