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


def test_numeric_loglikelihood_maximum_drives_termination():
    """"LogLikelihoodMaximum" -> number (BS:847): the estimate of the missing evidence uses it instead of the running
    maximum of the live set (BS:925-932), so a larger value keeps the run going and a smaller one ends it at
    MinIterations."""
    obj = _c1_obj()
    kw = dict(SamplePoolSize=40, MonteCarloSteps=30, MaxIterations=3000, MinIterations=50, Seed=4, PostProcessSamplingRuns=None)
    auto = api.nestedSampling(obj, **kw)
    lmax = auto["LogLikelihoodMaximum"]
    same = api.nestedSampling(obj, LogLikelihoodMaximum="Automatic", **kw)
    assert same["TotalSamples"] == auto["TotalSamples"]
    longer = api.nestedSampling(obj, LogLikelihoodMaximum=lmax + 8.0, **kw)
    shorter = api.nestedSampling(obj, LogLikelihoodMaximum=lmax - 200.0, **kw)
    assert longer["GeneratedNestedSamples"] > auto["GeneratedNestedSamples"] + 200  # ~8 nats more shrinkage at n = 40
    assert shorter["GeneratedNestedSamples"] == 50
    # the reported maximum stays the observed one (BS:1183-1189)
    assert abs(longer["LogLikelihoodMaximum"] - lmax) < 0.5 and shorter["LogLikelihoodMaximum"] < lmax


def test_evidence_sampling_on_a_finished_result():
    """evidenceSampling[obj, opts] re-post-processes a finished run (BS:1158-1160): the derived per-sample columns
    (dict-valued SampledLogX / LogPosteriorWeight included) are recomputed, not re-sorted as data."""
    obj = _c1_obj()
    res = api.nestedSampling(obj, SamplePoolSize=40, MonteCarloSteps=30, MaxIterations=150, Seed=6)
    again = api.evidenceSampling(res, PostProcessSamplingRuns=50, Seed=3)
    assert not again.failed
    assert again["TotalSamples"] == res["TotalSamples"]
    assert abs(again["CrudeLogEvidence"] - res["CrudeLogEvidence"]) < 1e-12
    assert abs(again["LogEvidence"]["Mean"] - res["LogEvidence"]["Mean"]) < 3 * res["LogEvidence"]["StandardError"]
    assert again["Samples"]["SampledLogX"]["Mean"].shape == (res["TotalSamples"],)
    assert abs(again["Samples"]["CrudePosteriorWeight"].sum() - 1) < 1e-9
    # same options, same seed: idempotent
    twice = api.evidenceSampling(again, PostProcessSamplingRuns=50, Seed=3)
    assert twice["LogEvidence"] == again["LogEvidence"]
    np.testing.assert_array_equal(twice["Samples"]["Point"], again["Samples"]["Point"])
    # crude-only form keeps working on a post-processed object
    crude = api.evidenceSampling(again, PostProcessSamplingRuns=None)
    assert abs(crude["CrudeLogEvidence"] - res["CrudeLogEvidence"]) < 1e-12


def _literal_reference_merge_logz(runs):
    """BS:1293-1315 + BS:785-799 + BS:756-771 evaluated literally in numpy: Join, DeleteDuplicatesBy Point, SortBy
    {logL, Point}; X from the constant Total[SamplePoolSize]; trapezoid weights; logSumExp."""
    from oracle import oracle as O
    pts = np.concatenate([r["Samples"]["Point"] for r in runs])
    L = np.concatenate([r["Samples"]["LogLikelihood"] for r in runs])
    _, first = np.unique(pts, axis=0, return_index=True)
    keep = np.sort(first)
    pts, L = pts[keep], L[keep]
    o = np.lexsort(tuple(pts[:, j] for j in range(pts.shape[1] - 1, -1, -1)) + (L,))
    L = L[o]
    n_tot = sum(int(r["SamplePoolSize"]) for r in runs)
    nd = L.size - n_tot
    lx = np.concatenate([-np.arange(1, nd + 1) / n_tot, np.log(np.arange(n_tot, 0, -1)) - np.log(n_tot + 1) - nd / n_tot])
    return O.logsumexp(O.trapezoid_log(lx) + L), lx


def test_combine_runs_merge_schemes():
    """combineRuns: K = 1 runs merge with the reference's literal X sequence (constant Total[SamplePoolSize], then
    n_tot..1 — BS:1307, 785-799); the summed-pool-size scheme stays close to it; K > 1 runs select the summed
    scheme, monotone to the end (no mixed sequences)."""
    obj = _c1_obj()
    runs = [api.nestedSampling(obj, SamplePoolSize=40, MonteCarloSteps=40, MaxIterations=400, Seed=15 + i,
                               PostProcessSamplingRuns=None) for i in range(4)]
    m = api.combineRuns(*runs, PostProcessSamplingRuns=30)
    assert m["MergeScheme"] == "Reference"
    M, n_tot = m["TotalSamples"], 160
    want_z, want_lx = _literal_reference_merge_logz(runs)
    assert abs(m["CrudeLogEvidence"] - want_z) < 1e-10
    o = np.argsort(m["Samples"]["LogLikelihood"], kind="stable")
    np.testing.assert_allclose(m["Samples"]["LogX"][o], want_lx, rtol=1e-13)
    pool = m["Samples"]["PoolSize"][o]
    assert np.all(pool[:M - n_tot] == n_tot) and np.array_equal(pool[M - n_tot:], np.arange(n_tot, 0, -1))
    # the dynamic-merge scheme on the same runs: same samples, different X; evidences agree well inside the error
    mp = api.combineRuns(*runs, PostProcessSamplingRuns=30, MergeScheme="PoolSizes")
    assert mp["MergeScheme"] == "PoolSizes" and mp["TotalSamples"] == M
    pp = mp["Samples"]["PoolSize"][np.argsort(mp["Samples"]["LogLikelihood"], kind="stable")]
    assert np.all(np.diff(pp) <= 0) and pp[0] == n_tot and pp[-1] == 1  # monotone, ends as a live set
    assert abs(mp["CrudeLogEvidence"] - m["CrudeLogEvidence"]) < 0.5 * m["LogEvidence"]["StandardError"]
    assert abs(mp["LogEvidence"]["Mean"] - m["LogEvidence"]["Mean"]) < 1.0 * m["LogEvidence"]["StandardError"]
    # a merged object merges again (pool structure is the reference's)
    mm = api.combineRuns(m, runs[0], PostProcessSamplingRuns=10)
    assert mm["SamplePoolSize"] == 200
    # K > 1 runs: pool sizes n, n-1, .. per batch -> summed scheme by default
    kr = [api.nestedSampling(obj, SamplePoolSize=40, MonteCarloSteps=40, MaxIterations=400, Seed=25 + i, BatchSize=4,
                             PostProcessSamplingRuns=None) for i in range(3)]
    mk = api.combineRuns(*kr, PostProcessSamplingRuns=30)
    assert mk["MergeScheme"] == "PoolSizes" and mk["SamplePoolSize"] == 120
    pk = mk["Samples"]["PoolSize"][np.argsort(mk["Samples"]["LogLikelihood"], kind="stable")]
    assert pk[-1] == 1 and pk.max() <= 120 and np.isfinite(mk["LogEvidence"]["Mean"])
    lx = mk["Samples"]["LogX"][np.argsort(mk["Samples"]["LogLikelihood"], kind="stable")]
    assert np.all(np.diff(lx) < 0)
    with pytest.raises(TypeError):
        api.combineRuns(*runs, MergeScheme="nope")


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


def _merge_reference(tables, pool_sizes):
    """combineRuns' merge the slow, literal way (BS:1293-1297): one searchsorted per run."""
    lex = lambda P, L: np.lexsort(tuple(P[:, j] for j in range(P.shape[1] - 1, -1, -1)) + (L,))
    pts = np.concatenate([t["Point"] for t in tables])
    L = np.concatenate([t["LogLikelihood"] for t in tables])
    _, first = np.unique(pts, axis=0, return_index=True)
    keep = np.sort(first)
    pts, L = pts[keep], L[keep]
    o = lex(pts, L)
    pts, L = pts[o], L[o]
    pool = np.zeros(L.size, dtype=np.int64)
    for t, n in zip(tables, pool_sizes):
        oo = lex(t["Point"], t["LogLikelihood"])
        tl, tp = t["LogLikelihood"][oo], t["PoolSize"][oo]
        idx = np.searchsorted(tl, L, side="left")
        pool += np.where(idx < tl.size, tp[np.minimum(idx, tl.size - 1)], 0)
    return pts, L, pool


def test_merge_samples_matches_literal_merge():
    """The O(M log M) merge (one cumulative sum over all samples) equals the per-run formulation: duplicate points
    (first kept), ties in logL within and across runs, K > 1 pool patterns, unsorted input tables."""
    rng = np.random.default_rng(4)
    for R, M, n, K in [(3, 60, 10, 1), (6, 300, 32, 8), (9, 1000, 100, 25)]:
        tables = []
        for r in range(R):
            L = np.sort(np.round(rng.normal(size=M), 2 if r % 2 else 12))
            P = np.round(rng.normal(size=(M, 2)), 1 if r % 3 == 0 else 9)
            for _ in range(4):
                a, b = rng.integers(0, M, 2)
                P[b] = P[a]
            nd = M - n
            pool = np.concatenate([np.tile(np.arange(n, n - K, -1), nd // K + 1)[:nd], np.arange(n, 0, -1)])
            perm = rng.permutation(M)
            tables.append({"Point": P[perm], "LogLikelihood": L[perm], "LogPriorPDF": rng.normal(size=M),
                           "AcceptanceRate": rng.uniform(size=M), "PoolSize": pool[perm]})
        got = api._merge_samples(tables, [n] * R)
        pts, L, pool = _merge_reference(tables, [n] * R)
        np.testing.assert_array_equal(got["Point"], pts)
        np.testing.assert_array_equal(got["LogLikelihood"], L)
        np.testing.assert_array_equal(got["PoolSize"], pool)
        assert got["PoolSize"][0] == R * n  # below every sample all runs are complete: K9, pool = sum n_r


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


def test_predictive_distribution_host_logic():
    """predictiveDistribution (BS:1373-1483) on the oracle backend: i.i.d., regression, point estimates, failures."""
    obj = _c1_obj()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert api.predictiveDistribution(obj) == api.FAILED
    assert any("unsampled" in str(x.message) for x in w)
    res = api.nestedSampling(obj, SamplePoolSize=30, MaxIterations=60, MinIterations=60, MonteCarloSteps=20, PostProcessSamplingRuns=5)
    mix = api.predictiveDistribution(res)
    S = res["Samples"]
    wts = S["CrudePosteriorWeight"] / S["CrudePosteriorWeight"].sum()
    np.testing.assert_allclose(mix.mean(), wts @ S["Point"][:, 0])
    ml = api.predictiveDistribution(res, point_estimate="MaximumLikelihood")
    i = int(np.argmax(S["LogLikelihood"]))
    assert ml.means.shape == (1,) and ml.means[0] == S["Point"][i, 0] and ml.sds[0] == S["Point"][i, 1]
    assert api.predictiveDistribution(res, [1.0]) == api.FAILED  # no independent variables

    c = cfg.c2_polyreg(N=300)
    reg = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), IndependentVariables=["x"],
        GeneratingDistribution=api.NormalDistribution(api.Polynomial("x", tuple(c.names[:4])), "sigma"),
        Parameters=[(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)],
        PriorDistribution=["LocationParameter"] * 4 + ["ScaleParameter"], _backend_override=OB)
    rr = api.nestedSampling(reg, SamplePoolSize=40, BatchSize=8, MaxIterations=30, MinIterations=30, MonteCarloSteps=20,
                            PostProcessSamplingRuns=5)
    xs = np.array([-0.5, 0.0, 0.75])
    pd_ = api.predictiveDistribution(rr, xs)
    P = rr["Samples"]["Point"]
    assert pd_.components.shape == (P.shape[0], 3, 2) and list(pd_.keys()) == [-0.5, 0.0, 0.75]
    np.testing.assert_allclose(pd_.components[:, 2, 0], P[:, 0] + P[:, 1] * 0.75 + P[:, 2] * 0.75**2 + P[:, 3] * 0.75**3, rtol=1e-13)
    np.testing.assert_array_equal(pd_.components[:, 1, 1], P[:, 4])
    named = api.predictiveDistribution(rr, xs, keys=["a", "b", "c"], point_estimate="MAP")
    assert list(named.keys()) == ["a", "b", "c"] and named["a"].means.shape == (1,)
    assert api.predictiveDistribution(rr, xs, keys=["a"]) == api.FAILED
    assert api.predictiveDistribution(rr) == api.FAILED


def test_approximate_evidence_laplace_matches_quadrature_scale():
    """approximateEvidence / laplaceLogEvidence (LA:22-30, 177-238) on the oracle backend: C1 has the quadrature pin
    -114.641064; the Laplace value of this skewed 2-parameter posterior sits within ~0.1 of it, and the mode is the MAP
    of mu (sample mean) and of sigma under the 1/sigma prior."""
    assert abs(api.laplaceLogEvidence(-3.0, [[2.0, 0.0], [0.0, 8.0]]) - (-3.0 + np.log(2 * np.pi) - 0.5 * np.log(16.0))) < 1e-14
    assert not api.laplaceLogEvidence(0.0, [[1.0, 2.0], [2.0, 1.0]])  # Missing[]: determinant < 0
    obj = _c1_obj()
    res = api.approximateEvidence(obj)  # no samples: starts from the best of 4096 prior draws
    c = cfg.c1_gaussian()
    x = c.inputs[:, 0]
    N = x.size
    assert abs(res["Mean"][0] - x.mean()) < 1e-6
    assert abs(res["Mean"][1] - np.sqrt(((x - x.mean()) ** 2).sum() / (N + 1))) < 1e-6  # d/ds [-(N+1) log s - S/(2 s^2)] = 0
    np.testing.assert_allclose(res["PrecisionMatrix"][0, 0], N / res["Mean"][1] ** 2, rtol=1e-5)
    assert abs(res["LogEvidence"] - c.truth["logZ"]) < 0.1
    assert res["Parameters"] == ["mu", "sigma"] and set(res["Maximum"][1]) == {"mu", "sigma"}
    # from a finished run the best sample is the start and the posterior spread sets the stencil
    run = api.nestedSampling(obj, SamplePoolSize=40, MaxIterations=300, MonteCarloSteps=30, PostProcessSamplingRuns=5)
    res2 = api.approximateEvidence(run)
    np.testing.assert_allclose(res2["Mean"], res["Mean"], atol=1e-6)
    np.testing.assert_allclose(res2["LogEvidence"], res["LogEvidence"], atol=1e-6)


def test_create_mcmc_chain_and_iterate_host_logic():
    """createMCMCChain / iterateMCMC (BS:651-703) on the oracle backend: option forms, thinning, failure modes."""
    obj = _c1_obj()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert not api.inferenceObjectQ(api.createMCMCChain(obj))  # createMCMCChain::start
    assert any("createMCMCChain::start" in str(x.message) for x in w)
    ch = api.createMCMCChain(obj, [1.5, 0.7], InitialCovariance=0.01, Seed=3)
    a = api.iterateMCMC(ch, 40)
    b = api.iterateMCMC(ch, (10, 5))  # 10 states, one every 5 steps
    assert a.shape == (40, 2) and b.shape == (10, 2)
    ref = api.iterateMCMC(api.createMCMCChain(obj, [1.5, 0.7], InitialCovariance=[0.01, 0.01], Seed=3), 90)
    np.testing.assert_array_equal(a, ref[:40])
    np.testing.assert_array_equal(b, ref[40:][4::5])
    x, t, mean, cov = ch["StateData"]
    assert t == 91 and x.shape == (2,) and cov.shape == (2, 2)
    np.testing.assert_array_equal(x, ref[-1])
    assert 0.0 < ch["AcceptanceRate"] <= 1.0
    # starting points from the object (BS:657-658), several chains
    sp = api.generateStartingPoints(obj, 3, seed=2)
    ch3 = api.createMCMCChain(sp, Chains=3, InitialCovariance=np.eye(2) * 0.01)
    assert api.iterateMCMC(ch3, 7).shape == (7, 3, 2)
    with pytest.raises(TypeError):
        api.createMCMCChain(obj, [1.5, 0.7], Nope=1)
    with pytest.raises(ValueError):
        api.iterateMCMC(api.createMCMCChain(obj, [1.5, -0.7]), 3)  # start outside the box


def test_predict_from_gaussian_process_host_logic():
    """predictFromGaussianProcess (GP:332-393) on the oracle backend: grid construction, weights, mixture moments."""
    c = cfg.c5_gp(N=40)
    pars = [(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)]
    obj = api.defineGaussianProcess((c.inputs[:, 0], c.outputs[:, 0]), api.SquaredExponentialGP(*c.names), pars,
                                    ["ScaleParameter"] * 3, _backend_override=OB)
    assert api.inferenceObjectQ(obj) and "GaussianProcessData" in obj
    assert api.predictFromGaussianProcess(obj, 5) == api.FAILED  # no "Samples" yet: the definition does not match
    res = api.nestedSampling(obj, SamplePoolSize=20, BatchSize=5, MonteCarloSteps=10, MaxIterations=8, MinIterations=8,
                             PostProcessSamplingRuns=5, Seed=3)
    pred = api.predictFromGaussianProcess(res, 6)
    x = c.inputs[:, 0]
    np.testing.assert_allclose(pred.points[:, 0], np.linspace(x.min(), x.max(), 6))
    assert list(pred.keys()) == [float(v) for v in pred.points[:, 0]]
    S = res["Samples"]
    M = S["Point"].shape[0]
    assert pred.means.shape == (M, 6) and pred.sds.shape == (M, 6)
    w = S["CrudePosteriorWeight"] / S["CrudePosteriorWeight"].sum()
    np.testing.assert_allclose(pred.weights, w)
    # one component by hand (closed form with numpy): sample 0, input 2
    sf, ell, sn = S["Point"][0]
    K = sf**2 * np.exp(-(x[:, None] - x[None]) ** 2 / (2 * ell**2)) + sn**2 * np.eye(x.size)
    ks = sf**2 * np.exp(-(x - pred.points[2, 0]) ** 2 / (2 * ell**2))
    np.testing.assert_allclose(pred.means[0, 2], ks @ np.linalg.solve(K, c.outputs[:, 0]), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(pred.sds[0, 2], np.sqrt(sf**2 + sn**2 - ks @ np.linalg.solve(K, ks)), rtol=1e-7)
    mix = pred[float(pred.points[2, 0])]
    np.testing.assert_allclose(mix.mean(), w @ pred.means[:, 2])
    np.testing.assert_allclose(mix.variance(), w @ (pred.sds[:, 2] ** 2 + pred.means[:, 2] ** 2) - mix.mean() ** 2, rtol=1e-9)
    assert abs(mix.cdf(mix.quantile(0.3)) - 0.3) < 1e-9
    xs = np.linspace(mix.mean() - 8 * mix.sd(), mix.mean() + 8 * mix.sd(), 4001)
    assert abs(np.trapezoid(mix.pdf(xs), xs) - 1.0) < 1e-6
    # explicit inputs, duplicates collapse like association keys; wrong dimension fails
    p2 = api.predictFromGaussianProcess(res, [1.0, 2.5, 1.0])
    assert list(p2.keys()) == [1.0, 2.5]
    assert api.predictFromGaussianProcess(res, np.zeros((3, 2))) == api.FAILED
    assert api.predictFromGaussianProcess(res, 1) == api.FAILED


def test_shard_plan():
    for R, W in [(64, 8), (4, 3), (1, 2), (5, 5)]:
        parts = [api._shard(R, r, W) for r in range(W)]
        assert sum(c for _, c in parts) == R
        assert parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(W - 1))


# ---------------------------------------------------------------------------------------------------
# data-sharded mode: host logic on CPU — row plan, and "DataSharding" under gloo, world_size 2
def test_shard_rows_plan():
    from bayesianinference_b200.engine import shard_rows
    for n, W in [(10, 2), (11, 3), (1_000_000, 8), (7, 7)]:
        parts = [shard_rows(n, r, W) for r in range(W)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(W - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    # series of increments (GBM): consecutive shards share exactly one point, every increment counted once
    for n, W in [(16385, 8), (9, 4)]:
        parts = [shard_rows(n, r, W, overlap=1) for r in range(W)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] - 1 == parts[i + 1][0] for i in range(W - 1))
        assert sum(b - a - 1 for a, b in parts) == n - 1
    with pytest.raises(ValueError):
        shard_rows(3, 0, 4)


def _sharded_objs():
    c2 = cfg.c2_polyreg(N=4001)
    o2 = api.defineInferenceProblem(
        Data=(c2.inputs[:, 0], c2.outputs[:, 0]), GeneratingDistribution=api.NormalDistribution(
            api.Polynomial("x", ("c0", "c1", "c2", "c3")), "sigma"),
        Parameters=[(n, lo, hi) for n, lo, hi in zip(c2.names, c2.lo, c2.hi)], IndependentVariables=["x"],
        PriorDistribution=["LocationParameter"] * 4 + ["ScaleParameter"], DataSharding="Automatic",
        _backend_override=OB)
    c4 = cfg.c4_gbm(T=1000)
    o4 = api.defineInferenceProblem(
        Data=(c4.inputs[:, 0], c4.outputs[:, 0]), GeneratingDistribution=api.GeometricBrownianMotionProcess("mu", "sigma"),
        Parameters=[("mu", -1, 1), ("sigma", 0.01, 2)], PriorDistribution=["LocationParameter", "ScaleParameter"],
        DataSharding="Automatic", _backend_override=OB)
    th2 = OB.Problem(c2.op, c2.inputs, c2.outputs, c2.iparam, c2.kinds, c2.lo, c2.hi).sample_prior(6, 3)
    th2[5, 4] = -1.0  # constraint violation -> logzero on every rank
    th4 = np.array([[0.1, 0.3], [-0.5, 1.0]])
    return o2, th2, o4, th4


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o2, th2, o4, th4 = _sharded_objs()
    assert o2["DataSharding"].world == world
    q.put((rank, o2["LogLikelihoodFunction"](th2), o4["LogLikelihoodFunction"](th4)))
    dist.barrier()
    dist.destroy_process_group()


def _batch_shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = cfg.c5_gp(N=24)
    pars = [(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)]
    obj = api.defineGaussianProcess((c.inputs[:, 0], c.outputs[:, 0]), api.SquaredExponentialGP(*c.names), pars,
                                    ["ScaleParameter"] * 3, DataSharding="Automatic", _backend_override=OB)
    th = obj["_problem"].sample_prior(7, 5)  # 7 vectors over 2 ranks: slices of 4 and 3
    res = api.nestedSampling(obj, SamplePoolSize=12, BatchSize=5, MonteCarloSteps=6, MaxIterations=10, MinIterations=10,
                             PostProcessSamplingRuns=None, Seed=2)
    q.put((rank, obj["LogLikelihoodFunction"](th), res["Samples"]["LogLikelihood"]))
    dist.barrier()
    dist.destroy_process_group()


def test_batch_sharding_gloo_matches_unsharded():
    """Batch-sharded mode host logic, world_size 2 on CPU: the GP problem splits every theta batch
    into contiguous slices, one per rank, and every rank ends with all values — likelihood batches (ragged: 7 over 2)
    and a short nested-sampling run equal the unsharded ones exactly."""
    import torch.multiprocessing as tmp
    c = cfg.c5_gp(N=24)
    pars = [(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)]
    obj = api.defineGaussianProcess((c.inputs[:, 0], c.outputs[:, 0]), api.SquaredExponentialGP(*c.names), pars,
                                    ["ScaleParameter"] * 3, _backend_override=OB)
    th = obj["_problem"].sample_prior(7, 5)
    want = obj["LogLikelihoodFunction"](th)
    res = api.nestedSampling(obj, SamplePoolSize=12, BatchSize=5, MonteCarloSteps=6, MaxIterations=10, MinIterations=10,
                             PostProcessSamplingRuns=None, Seed=2)
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 30700 + os.getpid() % 2000
    procs = [ctx.Process(target=_batch_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted((q.get(timeout=240) for _ in procs), key=lambda t: t[0])
    [p.join(60) for p in procs]
    for _, ll, L in out:
        np.testing.assert_array_equal(ll, want)
        np.testing.assert_array_equal(L, res["Samples"]["LogLikelihood"])


def test_data_sharding_gloo_matches_unsharded():
    import torch.multiprocessing as tmp
    o2, th2, o4, th4 = _sharded_objs()  # no process group here: "Automatic" -> unsharded
    assert o2["DataSharding"] is None
    ref2, ref4 = o2["LogLikelihoodFunction"](th2), o4["LogLikelihoodFunction"](th4)
    assert ref2[5] < -1e300
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=240) for _ in procs)
    [p.join(60) for p in procs]
    for _, l2, l4 in out:
        np.testing.assert_allclose(l2[:5], ref2[:5], rtol=1e-12)
        assert l2[5] < -1e300
        np.testing.assert_allclose(l4, ref4, rtol=1e-12)
    assert np.array_equal(out[0][1], out[1][1])  # same numbers in the same order on every rank


def test_gp_group_schedule_covers_every_panel_column_pair_once():
    """The update schedules of the blocked Cholesky sweep (csrc/gp.cu, gp_sweep): panels in groups of nblk, one
    full-width update per group, and inside a group either left-looking (the default: right before panel k is factored
    its column block receives all earlier panels of the group in one pass) or the binary schedule (after o panels of the
    group the last lowbit(o) update the next lowbit(o) column blocks).  Replayed here on block indices: when a column
    block is factored it has received every earlier panel exactly once, and nothing is ever applied twice.  (Mirror of
    the C++ loop; the numerical result is covered by the GPU parity tests.)"""
    for left in (True, False):
        for T in (1, 2, 5, 8, 13, 32):
            for nblk in (1, 2, 3, 4, 8, 16):
                got = {c: [] for c in range(T)}  # column block -> panels applied so far
                for kb in range(0, T, nblk):
                    kend = min(kb + nblk, T)
                    for k in range(kb, kend):
                        if left and k > kb:
                            got[k] += list(range(kb, k))
                        assert sorted(got[k]) == list(range(k)), (left, T, nblk, k, got[k])  # complete when factored
                        o = k - kb + 1
                        if not left and o < kend - kb:
                            w = o & -o
                            cols = min(w, kend - (k + 1))
                            for c in range(k + 1, k + 1 + cols):
                                got[c] += list(range(k + 1 - w, k + 1))
                    for c in range(kend, T):
                        got[c] += list(range(kb, kend))
                assert all(len(v) == len(set(v)) for v in got.values())


def test_merging_partial_merges_equals_merging_all_runs():
    """The multi-GPU path of parallelNestedSampling merges per-GPU merges (BS:1293-1297 applied twice).  Summed pool sizes
    are step functions of the likelihood level and add up, so the two-level merge must equal the flat one — here with the
    numpy merge (the device merge is checked against it in tests/test_gpu_merge.py), and both against the oracle's
    independent formulation (np.unique + one searchsorted per run)."""
    from types import SimpleNamespace

    from oracle import oracle as O
    rng = np.random.default_rng(12)
    runs = []
    for r in range(5):
        n, K, iters = 24, 3, 10
        M = iters * K + n
        L = np.sort(rng.normal(size=M)) * 10 - 50
        pts = rng.normal(size=(M, 2))
        if r:  # a few copies of earlier runs' samples (same point => same likelihood) and ties with different points
            src = runs[r - 1]
            pick = rng.choice(src["LogLikelihood"].size, 4, replace=False)
            where = rng.choice(M, 4, replace=False)
            L[where], pts[where] = src["LogLikelihood"][pick], src["Point"][pick]
        L[5] = L[4]
        o = np.lexsort((pts[:, 1], pts[:, 0], L))
        pool = np.concatenate([np.tile(np.arange(n, n - K, -1), iters), np.arange(n, 0, -1)]).astype(np.int64)
        runs.append({"Point": pts[o], "LogLikelihood": L[o], "LogPriorPDF": rng.normal(size=M), "AcceptanceRate": rng.random(M),
                     "PoolSize": pool})
    flat = api._merge_samples(runs, [24] * 5)
    a = api._merge_samples(runs[:3], [24] * 3)
    b = api._merge_samples(runs[3:], [24] * 2)
    two = api._merge_samples([a, b], [72, 48])
    for k in ("Point", "LogLikelihood", "LogPriorPDF", "AcceptanceRate", "PoolSize"):
        assert np.array_equal(two[k], flat[k]), k
    want = O.combine_runs([SimpleNamespace(points=t["Point"], logL=t["LogLikelihood"], logPrior=t["LogPriorPDF"],
                                           acc=t["AcceptanceRate"], pool=t["PoolSize"], n=24) for t in runs])
    assert np.array_equal(flat["Point"], want["points"]) and np.array_equal(flat["PoolSize"], want["pool"])
    assert np.array_equal(flat["RunIndex"], want["run_id"])
