"""Host vs device combineRuns on C4-shaped synthetic runs (64 runs x (57 x 64 + 512) samples): wall clock of the pieces.
BINEST_MERGE_DEBUG=1 prints the device phases (csrc/merge.cu)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

from bayesianinference_b200 import api, engine  # noqa: E402
from test_gpu_merge import _as_objects, _synthetic_runs  # noqa: E402

R, n, K, iters = 64, 512, 64, 57
runs = _synthetic_runs(R, n, K, iters, seed=1, dup_frac=0.001, tie_frac=0.0005)
objs = _as_objects(runs, n)
engine.init()
for rep in range(3):
    t0 = time.perf_counter()
    tabs = [api._sorted_run_table(o.Normal()["Samples"], n) for o in objs]
    t1 = time.perf_counter()
    j = engine._join_runs(tabs)
    t2 = time.perf_counter()
    res = engine.combine_runs(tabs, False, R * n, 100, 1)
    t3 = time.perf_counter()
    print(f"rep {rep}: sorted tables {1e3 * (t1 - t0):.1f} ms, join {1e3 * (t2 - t1):.1f} ms, combine_runs (incl. join) {1e3 * (t3 - t2):.1f} ms")
for rep in range(2):
    t0 = time.perf_counter()
    dev = api.combineRuns(*objs, PostProcessSamplingRuns=100)
    t1 = time.perf_counter()
    api._HOST_MERGE = True
    host = api.combineRuns(*objs, PostProcessSamplingRuns=100)
    api._HOST_MERGE = False
    t2 = time.perf_counter()
    print(f"api.combineRuns device {1e3 * (t1 - t0):.1f} ms, host {1e3 * (t2 - t1):.1f} ms; logZ {dev['LogEvidence']} {host['LogEvidence']}; "
          f"equal order: {np.array_equal(dev['Samples']['LogLikelihood'], host['Samples']['LogLikelihood'])}")
