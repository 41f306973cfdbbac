"""Host logic of bayesianinference_b200.api (the mirror of the WL API) on CPU, served by the test-only
oracle backend: problem definition and its failure modes, result associations, combineRuns, and
parallelNestedSampling run-sharding under torch.distributed (gloo, world_size 2)."""
import os
import sys
import warnings

import numpy as np
import pytest

from bayesianinference_b200 import api
from bayesianinference_b200 import configs as cfg

sys.path.insert(0, os.path.dirname(__file__))
import oracle_backend as OB  # noqa: E402


def _c1_obj():
    c = cfg.c1_gaussian()
    return api.defineInferenceProblem(
        Data=c.inputs[:, 0], GeneratingDistribution=api.NormalDistribution("mu", "sigma"),
        Parameters=[("mu", -10, 10), ("sigma", 0.01, 10)], PriorDistribution=["LocationParameter", "ScaleParameter"],
        _backend_override=OB)


def test_define_inference_problem_and_properties():
    obj = _c1_obj()
    assert api.inferenceObjectQ(obj)
    assert obj["ParameterSymbols"] == ["mu", "sigma"]
    ll = obj["LogLikelihoodFunction"](np.array([[1.5, 0.7], [0.0, -1.0]]))
    assert np.isfinite(ll[0]) and ll[1] < -1e300  # constraint violation -> logzero
    assert not obj["Nope"]  # Missing["KeyAbsent", ...]
    assert "LogPriorPDFFunction" in obj.keys()


def test_definition_failures_return_failed_object():
    c = cfg.c1_gaussian()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        bad = api.defineInferenceProblem(Data=c.inputs[:, 0], GeneratingDistribution="PoissonDistribution[mu]",
                                         Parameters=[("mu", 0, 10)], PriorDistribution=["LocationParameter"],
                                         _backend_override=OB)
        assert bad.failed and repr(bad) == "inferenceObject[$Failed]"
        assert any("logLike" in str(x.message) for x in w)
        miss = api.defineInferenceProblem(Data=c.inputs[:, 0], _backend_override=OB)
        assert miss.failed
    assert api.nestedSampling(bad).failed and api.combineRuns(bad).failed


def test_polynomial_descriptor_maps_to_operator():
    c = cfg.c2_polyreg(N=500)
    obj = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), IndependentVariables=["x"],
        GeneratingDistribution=api.NormalDistribution(api.Polynomial("x", ("c0", "c1", "c2", "c3")), "sigma"),
        Parameters=[(n, lo, hi) for n, lo, hi in zip(c.names, c.lo, c.hi)],
        PriorDistribution=["LocationParameter"] * 4 + ["ScaleParameter"], _backend_override=OB)
    assert api.inferenceObjectQ(obj)
    th = np.array([[0.5, -1.2, 0.8, 0.3, 0.25]])
    x, y = c.inputs[:, 0], c.outputs[:, 0]
    ref = (-(y - (0.5 - 1.2 * x + 0.8 * x**2 + 0.3 * x**3)) ** 2 / (2 * 0.25**2) - np.log(0.25) - 0.5 * np.log(2 * np.pi)).sum()
    assert abs(obj["LogLikelihoodFunction"](th)[0] - ref) < 1e-9 * abs(ref)


def test_nested_sampling_result_association():
    obj = _c1_obj()
    res = api.nestedSampling(obj, SamplePoolSize=60, MonteCarloSteps=60, MaxIterations=400, Seed=3, BatchSize=4)
    for k in ("Samples", "SamplePoolSize", "GeneratedNestedSamples", "TotalSamples", "ParameterRanges", "LogEvidence",
              "CrudeLogEvidence", "LogLikelihoodMaximum", "LogEstimatedMissingEvidence", "CrudeRelativeEntropy",
              "ParameterExpectedValues", "RelativeEntropy", "EmpiricalPosteriorDistribution"):
        assert k in res, k
    S = res["Samples"]
    M = res["TotalSamples"]
    assert M == res["GeneratedNestedSamples"] + 60 and S["Point"].shape == (M, 2)
    assert np.all(np.diff(S["CrudeLogPosteriorWeight"]) <= 1e-12)  # sorted by -weight, BS:1241
    assert abs(S["CrudePosteriorWeight"].sum() - 1) < 1e-9
    assert np.isnan(S["AcceptanceRate"]).sum() == 60  # Missing["InitialSample"], BS:911
    assert set(res["ParameterExpectedValues"]) == {"mu", "sigma"}
    assert abs(res["ParameterExpectedValues"]["mu"]["Mean"] - 1.5558) < 0.1
    with pytest.raises(TypeError):
        api.nestedSampling(obj, NotAnOption=1)
    # supplied starting points fix the pool size (BS:1116-1131)
    sp = obj["_problem"].sample_prior(30, 4)
    res2 = api.nestedSampling(obj, StartingPoints=sp, MonteCarloSteps=20, MaxIterations=50, MinIterations=50)
    assert res2["SamplePoolSize"] == 30 and res2["GeneratedNestedSamples"] == 50
    # PostProcessSamplingRuns -> None skips the Monte-Carlo part (BS:1195-1197)
    res3 = api.nestedSampling(obj, SamplePoolSize=30, MonteCarloSteps=20, MaxIterations=50, PostProcessSamplingRuns=None)
    assert "LogEvidence" not in res3 and "CrudeLogEvidence" in res3


def test_combine_runs_pool_and_counts():
    obj = _c1_obj()
    runs = [api.nestedSampling(obj, SamplePoolSize=40, MonteCarloSteps=30, MaxIterations=150, Seed=5 + i,
                               PostProcessSamplingRuns=None) for i in range(3)]
    m = api.combineRuns(*runs, PostProcessSamplingRuns=20)
    assert m["SamplePoolSize"] == 120
    total = sum(r["TotalSamples"] for r in runs)
    uniq = np.unique(np.concatenate([r["Samples"]["Point"] for r in runs]), axis=0).shape[0]
    assert m["TotalSamples"] == uniq <= total and m["GeneratedNestedSamples"] == uniq - 120
    assert m["LogLikelihoodMaximum"] == max(r["LogLikelihoodMaximum"] for r in runs)
    assert np.isfinite(m["LogEvidence"]["Mean"]) and m["LogEvidence"]["StandardError"] > 0


def _parallel(seed=9):
    obj = _c1_obj()
    return api.parallelNestedSampling(obj, ParallelRuns=4, SamplePoolSize=30, MonteCarloSteps=25, MaxIterations=120,
                                      Seed=seed, BatchSize=2, PostProcessSamplingRuns=10)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = _parallel()
    q.put((rank, res["TotalSamples"], res["SamplePoolSize"], float(res["LogEvidence"]["Mean"]),
           res["Samples"]["LogLikelihood"].sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_parallel_nested_sampling_sharded_gloo_matches_single_process():
    """Runs are keyed by (seed, run id): sharding them over 2 ranks must give the same merged object."""
    import torch.multiprocessing as tmp
    single = _parallel()
    assert single["SamplePoolSize"] == 120  # 4 runs x 30 (BS:1307)
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=240) for _ in procs)
    [p.join(60) for p in procs]
    for _, M, n, z, s in out:
        assert (M, n) == (single["TotalSamples"], 120)
        assert abs(z - single["LogEvidence"]["Mean"]) < 1e-9
        assert abs(s - single["Samples"]["LogLikelihood"].sum()) < 1e-6


def test_shard_plan():
    for R, W in [(64, 8), (4, 3), (1, 2), (5, 5)]:
        parts = [api._shard(R, r, W) for r in range(W)]
        assert sum(c for _, c in parts) == R
        assert parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(W - 1))
