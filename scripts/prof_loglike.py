import sys; sys.path.insert(0, '.')
from bayesianinference_b200 import engine, configs as cfg
engine.init()
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
c = cfg.ALL[name]()
gp = engine.Problem.from_config(c)
print(name, P, gp.bench_loglike(P, 3, 2, False))
