"""combineRuns (BS:1293-1315) on the device (csrc/merge.cu) against the host merge of bayesianinference_b200.api
(numpy; itself pinned to a literal evaluation of the reference formula in tests/test_api_host.py).  Index work is exact:
the merged lists must be equal sample for sample."""
import os

import numpy as np
import pytest

from bayesianinference_b200 import api
from bayesianinference_b200 import configs as cfg

pytestmark = pytest.mark.gpu


def _synthetic_runs(R, n, K, iters, seed, d=2, dup_frac=0.02, tie_frac=0.02):
    """R run tables shaped like the engine's: sorted by {logL, point}, pool sizes n, n-1, .. n-K+1 per iteration then
    n..1; some samples are copies of samples of OTHER runs or of the SAME run (unmoved walkers), some share a logL
    with a different point."""
    rng = np.random.default_rng(seed)
    runs, bank = [], []
    for r in range(R):
        M = iters * K + n
        L = np.sort(rng.normal(size=M)) * 50 - 1000
        pts = rng.normal(size=(M, d))
        m = max(1, int(tie_frac * M))
        where = rng.choice(M - 1, m, replace=False)
        L[where + 1] = L[where]                     # equal likelihood, different point
        w2 = rng.choice(M - 1, max(1, m // 2), replace=False)
        L[w2 + 1] = L[w2]
        pts[w2 + 1] = pts[w2]                       # a duplicate inside the run
        if bank and dup_frac > 0:                   # copies of other runs' samples (same point => same likelihood)
            src = np.concatenate(bank)
            m = max(1, int(dup_frac * M))
            pick = rng.choice(len(src), m, replace=False)
            where = rng.choice(M, m, replace=False)
            L[where] = src[pick, 0]
            pts[where] = src[pick, 1:]
        o = np.lexsort(tuple(pts[:, j] for j in range(d - 1, -1, -1)) + (L,))
        L, pts = L[o], pts[o]
        bank.append(np.column_stack([L, pts]))
        pool = np.concatenate([np.tile(np.arange(n, n - K, -1), iters), np.arange(n, 0, -1)]).astype(np.int64)
        runs.append({"Point": pts, "LogLikelihood": L, "LogPriorPDF": rng.normal(size=M),
                     "AcceptanceRate": rng.random(M), "PoolSize": pool})
    return runs


@pytest.mark.parametrize("R,n,K,iters", [(1, 16, 1, 40), (3, 32, 4, 9), (8, 64, 8, 30), (64, 128, 16, 20), (5, 100, 1, 300)])
def test_device_merge_equals_host_merge(R, n, K, iters):
    from bayesianinference_b200 import engine
    runs = _synthetic_runs(R, n, K, iters, seed=R * 1000 + n)
    host = api._merge_samples(runs, [n] * R)
    dev, live = engine.merge_runs(runs)
    assert dev["LogLikelihood"].size == host["LogLikelihood"].size
    for k in ("Point", "LogLikelihood", "LogPriorPDF", "AcceptanceRate", "PoolSize", "RunIndex"):
        assert np.array_equal(dev[k], host[k]), k
    M = host["PoolSize"].size
    agree = host["PoolSize"] == np.arange(M, 0, -1)
    want = int(M - (np.flatnonzero(~agree)[-1] + 1)) if not agree.all() else M
    assert live == want


def test_device_merge_equals_oracle_restatement():
    """against oracle.combine_runs: an independent formulation (np.unique for the duplicates, one searchsorted per run for
    the summed pool sizes) of BS:1293-1297"""
    from types import SimpleNamespace

    from bayesianinference_b200 import engine
    from oracle import oracle as O
    runs = _synthetic_runs(7, 48, 6, 15, seed=404)
    dev, _ = engine.merge_runs(runs)
    want = O.combine_runs([SimpleNamespace(points=t["Point"], logL=t["LogLikelihood"], logPrior=t["LogPriorPDF"],
                                           acc=t["AcceptanceRate"], pool=t["PoolSize"], n=48) for t in runs])
    for k, kk in (("Point", "points"), ("LogLikelihood", "logL"), ("LogPriorPDF", "logPrior"), ("AcceptanceRate", "acc"),
                  ("PoolSize", "pool"), ("RunIndex", "run_id")):
        assert np.array_equal(dev[k], want[kk]), k


def test_device_merge_edge_cases():
    """an empty run, a run that is an exact copy of another (every sample a duplicate), a single-sample run"""
    from bayesianinference_b200 import engine
    runs = _synthetic_runs(3, 24, 2, 12, seed=31)
    empty = {k: v[:0] for k, v in runs[0].items()}
    copy = {k: v.copy() for k, v in runs[1].items()}
    one = {k: v[-1:].copy() for k, v in runs[2].items()}
    one["PoolSize"] = np.array([1], dtype=np.int64)
    for tabs, pools in (([runs[0], empty, runs[1]], [24, 24, 24]), ([runs[0], runs[1], copy], [24, 24, 24]),
                        ([one, runs[0]], [1, 24]), ([empty, runs[2], empty], [24, 24, 24])):
        host = api._merge_samples(tabs, pools)
        dev, _ = engine.merge_runs(tabs)
        for k in ("Point", "LogLikelihood", "LogPriorPDF", "AcceptanceRate", "PoolSize", "RunIndex"):
            assert np.array_equal(dev[k], host[k]), k


def test_device_merge_rejects_unsorted_runs():
    from bayesianinference_b200 import _lib, engine
    runs = _synthetic_runs(2, 16, 1, 10, seed=5, dup_frac=0, tie_frac=0)
    runs[1]["LogLikelihood"] = runs[1]["LogLikelihood"][::-1].copy()
    with pytest.raises(_lib.BinestError) as e:
        engine.merge_runs(runs)
    assert e.value.code == 3


def _as_objects(runs, n, names=("a", "b")):
    out = []
    for t in runs:
        M = t["LogLikelihood"].size
        out.append(api.inferenceObject({"Samples": dict(t), "SamplePoolSize": n, "ParameterSymbols": list(names),
                                        "TotalSamples": M, "GeneratedNestedSamples": M - n}))
    return out


@pytest.mark.parametrize("scheme,K", [("Reference", 1), ("PoolSizes", 1), ("PoolSizes", 8), ("Automatic", 4)])
def test_device_combine_equals_host_path(scheme, K):
    """binest_combine_runs (merge + X sequence + crude weights + evidenceSampling + sort by weight in one call) against
    the host path (numpy merge, then binest_crude_weights / binest_evidence_sampling, numpy sort)."""
    runs = _synthetic_runs(6, 64, K, 25, seed=77 + K)
    objs = _as_objects(runs, 64)
    dev = api.combineRuns(*objs, MergeScheme=scheme, PostProcessSamplingRuns=20, Seed=9)
    api._HOST_MERGE = True
    try:
        host = api.combineRuns(*objs, MergeScheme=scheme, PostProcessSamplingRuns=20, Seed=9)
    finally:
        api._HOST_MERGE = False
    for k in ("SamplePoolSize", "GeneratedNestedSamples", "TotalSamples", "MergeScheme"):
        assert dev[k] == host[k], k
    assert dev.Normal().get("_LiveBlock") == host.Normal().get("_LiveBlock")
    Sd, Sh = dev["Samples"], host["Samples"]
    for k in ("Point", "LogLikelihood", "LogPriorPDF", "AcceptanceRate", "PoolSize", "RunIndex", "LogX",
              "CrudeLogPosteriorWeight"):
        assert np.array_equal(Sd[k], Sh[k]), k
    for k in ("X", "CrudePosteriorWeight"):  # exp on the device vs numpy's
        assert np.allclose(Sd[k], Sh[k], rtol=4e-16, atol=0), k
    for k in ("SampledLogX", "LogPosteriorWeight"):
        for sub in ("Mean", "StandardError"):
            assert np.array_equal(Sd[k][sub], Sh[k][sub]), (k, sub)
    for k in ("CrudeLogEvidence", "LogLikelihoodMaximum", "LogEstimatedMissingEvidence", "CrudeRelativeEntropy"):
        assert dev[k] == host[k], k
    for k in ("LogEvidence", "RelativeEntropy"):
        assert dev[k] == host[k], k
    assert dev["ParameterExpectedValues"] == host["ParameterExpectedValues"]


def test_parallel_runs_device_and_host_merge_agree():
    """parallelNestedSampling end to end (C4-shaped, small): the device combine is what the call uses; the host merge of
    the same runs gives the same object."""
    c = cfg.c4_gbm(T=1024)
    obj = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), GeneratingDistribution=api.GeometricBrownianMotionProcess("mu", "sigma", 100.0),
        Parameters=[("mu", -1, 1), ("sigma", 0.01, 2)], PriorDistribution=["LocationParameter", "ScaleParameter"])
    kw = dict(ParallelRuns=6, SamplePoolSize=64, BatchSize=8, MaxIterations=100000, Seed=21, PostProcessSamplingRuns=30)
    dev = api.parallelNestedSampling(obj, **kw)
    assert "device_merge_s" in dev.Normal()["_Timing"]  # the device path ran
    api._HOST_MERGE = True
    try:
        host = api.parallelNestedSampling(obj, **kw)
    finally:
        api._HOST_MERGE = False
    assert "combine_merge_s" in host.Normal()["_Timing"]
    assert dev["TotalSamples"] == host["TotalSamples"] and dev["MergeScheme"] == host["MergeScheme"] == "PoolSizes"
    for k in ("Point", "LogLikelihood", "PoolSize", "RunIndex", "CrudeLogPosteriorWeight"):
        assert np.array_equal(dev["Samples"][k], host["Samples"][k]), k
    assert dev["LogEvidence"] == host["LogEvidence"]
    assert np.array_equal(dev["ParameterRanges"], host["ParameterRanges"])
    assert set(dev.keys()) == set(host.keys())


def test_run_group_merge_equals_merge_of_fetched_runs():
    """binest_run_merge (the Join built on the device from the engine's state) = binest_merge_runs of the runs fetched one
    by one = the host merge; and merging two partial merges equals merging all runs at once (the multi-GPU path)."""
    from bayesianinference_b200 import engine
    c = cfg.c4_gbm(T=512)
    gp = engine.Problem.from_config(c)
    o = engine.default_options(pool_size=48, batch_k=6, mc_steps=50, seed=5, n_runs=5, first_run_id=3)
    grp = engine.RunGroup(gp, o)
    grp.advance(0)
    tabs = []
    for i in range(5):
        s = grp.fetch(i, weights=False)
        tabs.append({"Point": s["points"], "LogLikelihood": s["logL"], "LogPriorPDF": s["logPrior"],
                     "AcceptanceRate": s["acc"], "PoolSize": s["pool"]})
    host = api._merge_samples(tabs, [48] * 5)
    dev, live = grp.merge()
    for k in ("Point", "LogLikelihood", "LogPriorPDF", "PoolSize"):
        assert np.array_equal(dev[k], host[k]), k
    assert np.array_equal(np.isnan(dev["AcceptanceRate"]), np.isnan(host["AcceptanceRate"]))
    assert np.array_equal(np.nan_to_num(dev["AcceptanceRate"]), np.nan_to_num(host["AcceptanceRate"]))
    assert np.array_equal(dev["RunIndex"], host["RunIndex"] + 3)
    # hierarchical: merge runs {0,1,2} and {3,4} separately, then merge the two merges
    a, _ = engine.merge_runs(tabs[:3])
    b, _ = engine.merge_runs(tabs[3:])
    b["RunIndex"] = b["RunIndex"] + 3
    two, live2 = engine.merge_runs([a, b])
    for k in ("Point", "LogLikelihood", "LogPriorPDF", "PoolSize"):
        assert np.array_equal(two[k], host[k]), k
    assert np.array_equal(two["RunIndex"], host["RunIndex"]) and live2 == live


def _two_gpu_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from bayesianinference_b200 import engine
    engine.init(device=rank)
    out = {}
    for mode in ("nccl", "host"):
        api._HOST_GATHER = mode == "host"
        res = _c4_small_parallel(5)  # 5 runs over 2 ranks: 3 + 2
        out[mode] = (res["Samples"]["Point"], res["Samples"]["LogLikelihood"], res["Samples"]["PoolSize"],
                     res["Samples"]["RunIndex"], res["LogEvidence"], sorted(res.Normal()["_Timing"]))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _c4_small_parallel(runs):
    c = cfg.c4_gbm(T=1024)
    obj = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), GeneratingDistribution=api.GeometricBrownianMotionProcess("mu", "sigma", 100.0),
        Parameters=[("mu", -1, 1), ("sigma", 0.01, 2)], PriorDistribution=["LocationParameter", "ScaleParameter"])
    return api.parallelNestedSampling(obj, ParallelRuns=runs, SamplePoolSize=64, BatchSize=8, MaxIterations=100000, Seed=33,
                                      PostProcessSamplingRuns=30)


def test_two_gpu_parallel_runs_nccl_gather_equals_single_gpu():
    """run-sharded parallelNestedSampling on 2 GPUs: per-GPU merges gathered over NCCL in device memory (and, second
    mode, through the host) and merged again — the same object as all runs on one GPU, on every rank."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 34500 + os.getpid() % 2000
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    allres = dict(q.get(timeout=600) for _ in procs)
    [p.join(120) for p in procs]
    one = _c4_small_parallel(5)
    want = (one["Samples"]["Point"], one["Samples"]["LogLikelihood"], one["Samples"]["PoolSize"], one["Samples"]["RunIndex"])
    for rank in (0, 1):
        for mode in ("nccl", "host"):
            got = allres[rank][mode]
            for a, b in zip(got[:4], want):
                assert np.array_equal(a, b), (rank, mode)
            assert got[4] == one["LogEvidence"], (rank, mode)
            assert "gather_s" in got[5] and "device_combine_s" in got[5]
