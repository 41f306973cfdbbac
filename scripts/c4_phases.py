"""Where the wall clock of a complete C4 run goes (64 parallel runs x 512 live points): problem definition, run group
creation, the device loop, fetching the 64 sample lists, host merge (combineRuns) and evidenceSampling."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
from bayesianinference_b200 import api, engine, configs as cfg
engine.init()
c = cfg.c4_gbm()
t = [time.perf_counter()]
obj = api.defineInferenceProblem(
    Data=(c.inputs[:, 0], c.outputs[:, 0]), GeneratingDistribution=api.GeometricBrownianMotionProcess("mu", "sigma", 100.0),
    Parameters=[(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)], PriorDistribution=["LocationParameter", "ScaleParameter"])
t.append(time.perf_counter())
a = obj.Normal()
R, n, K = 64, 512, 64
for rep in range(2):
    t = [time.perf_counter()]
    o = api._ns_options(dict(SamplePoolSize=n, BatchSize=K, MaxIterations=10**6, Seed=2026 + rep), dict(api.NS_DEFAULTS, ParallelRuns=4))
    grp = engine.RunGroup(a["_problem"], api._engine_options(engine, o, n_runs=R, first_run_id=0), None)
    t.append(time.perf_counter())
    grp.advance(0)
    t.append(time.perf_counter())
    tm = grp.timing()
    local = [grp.fetch(i) for i in range(R)]
    t.append(time.perf_counter())
    grp.close()
    runs = []
    for s in local:
        ra = dict(a); ra.update(api._result_assoc(s, n)); runs.append(api.inferenceObject(ra))
    t.append(time.perf_counter())
    res = api.combineRuns(*runs, PostProcessSamplingRuns=100, Seed=1)
    t.append(time.perf_counter())
    names = ["run group create", "advance (device loop)", "fetch x64", "result assoc", "combineRuns + evidenceSampling"]
    gen = res["GeneratedNestedSamples"]
    print(f"rep {rep}: " + ", ".join(f"{nm} {1e3 * (b - a_):.0f} ms" for nm, a_, b in zip(names, t, t[1:])) +
          f" | walk graphs {tm['walk_graphs']}, device walk {tm['walk_ms']:.0f} ms, batches {tm['batches']}, replacements {gen}, "
          f"{gen / (t[2] - t[1]):.0f} replacements/s in the device loop, logZ {res['LogEvidence']}")
