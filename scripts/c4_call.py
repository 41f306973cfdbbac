"""One parallelNestedSampling call on config C4 as BASELINE states it (64 runs x 512 live points, K = 64) — for the ncu
launch list of the whole call (walks, updates, merge, evidenceSampling)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianinference_b200 import api, configs as cfg  # noqa: E402

c = cfg.c4_gbm()
obj = api.defineInferenceProblem(
    Data=(c.inputs[:, 0], c.outputs[:, 0]), GeneratingDistribution=api.GeometricBrownianMotionProcess("mu", "sigma", 100.0),
    Parameters=[(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)], PriorDistribution=["LocationParameter", "ScaleParameter"])
res = api.parallelNestedSampling(obj, ParallelRuns=64, SamplePoolSize=512, BatchSize=64, MaxIterations=10**6, Seed=2026,
                                 PostProcessSamplingRuns=100)
print(res["LogEvidence"], res["TotalSamples"], res.Normal()["_Timing"])
