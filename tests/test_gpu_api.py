"""The reference-facing API (bayesianinference_b200.api) on the GPU engine, read like the reference's own
worked examples: define the problem, run nestedSampling / parallelNestedSampling, inspect LogEvidence."""
import json
import os

import numpy as np
import pytest

from bayesianinference_b200 import api
from bayesianinference_b200 import configs as cfg

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pins.json")))


def test_c1_example_style_run():
    c = cfg.c1_gaussian()
    obj = api.defineInferenceProblem(
        Data=c.inputs[:, 0], GeneratingDistribution=api.NormalDistribution("mu", "sigma"),
        Parameters=[("mu", -10, 10), ("sigma", 0.01, 10)], PriorDistribution=["LocationParameter", "ScaleParameter"])
    assert api.inferenceObjectQ(obj)
    res = api.nestedSampling(obj, Seed=12)  # reference defaults: 100 live points, 200 steps, K = 1
    z = res["LogEvidence"]
    assert abs(z["Mean"] - GOLD["c1_logZ_quadrature"]) < 3.5 * z["StandardError"], z
    assert 0.15 < z["StandardError"] < 0.45
    pe = res["ParameterExpectedValues"]
    assert abs(pe["mu"]["Mean"] - 1.5558) < 0.03 and abs(pe["sigma"]["Mean"] - 0.713) < 0.03
    assert 5.5 < res["RelativeEntropy"]["Mean"] < 9.5  # H ~ 7.4 nats
    assert res["TotalSamples"] == res["GeneratedNestedSamples"] + 100


def test_c4_parallel_runs_merge_and_evidence():
    c = cfg.c4_gbm()
    obj = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), GeneratingDistribution=api.GeometricBrownianMotionProcess("mu", "sigma", 100.0),
        Parameters=[("mu", -1, 1), ("sigma", 0.01, 2)], PriorDistribution=["LocationParameter", "ScaleParameter"])
    res = api.parallelNestedSampling(obj, ParallelRuns=8, SamplePoolSize=128, BatchSize=16, MaxIterations=100000, Seed=4)
    assert res["SamplePoolSize"] == 8 * 128
    z = res["LogEvidence"]
    assert abs(z["Mean"] - GOLD["c4_logZ_quadrature"]) < 4 * z["StandardError"] + 0.05, z
    assert z["StandardError"] < 0.2  # sqrt(H / (8*128)) ~ 0.09


def test_argument_errors_are_status_codes_not_crashes():
    """API misuse returns LibraryLink-style status codes (include/binest.h); operators never fail (logzero)."""
    from bayesianinference_b200 import _lib, engine
    c = cfg.c4_gbm(T=64)
    gp = engine.Problem(c.op, c.inputs, c.outputs, c.iparam, c.kinds, c.lo, c.hi)
    with pytest.raises(_lib.BinestError) as e:
        engine.RunGroup(gp, engine.default_options(pool_size=1))
    assert e.value.code == 3  # DIMENSION
    with pytest.raises(_lib.BinestError) as e:
        engine.Problem(c.op, c.inputs, c.outputs, c.iparam, c.kinds, c.hi, c.lo)  # lo > hi
    assert e.value.code == 3
    with pytest.raises(_lib.BinestError) as e:
        engine.Problem(99, c.inputs, c.outputs, c.iparam, c.kinds, c.lo, c.hi)
    assert e.value.code == 6  # FUNCTION: not in the operator table
    th = np.array([[0.1, np.nan], [0.1, -2.0]])
    assert np.all(gp.loglike(th) == _lib.LOGZERO)  # RuntimeErrorHandler -> logzero (BS:589-592)
