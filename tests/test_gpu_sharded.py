"""Data-sharded mode (include/binest.h "data-sharded mode") on 2 GPUs, through the C ABI:
rows split across the ranks, per-walker shard sums all-gathered over NCCL and combined in rank order.
  * logL of a theta batch == the __float128 oracle on the FULL data (1e-12 relative), identical bits on both ranks;
  * a short nested-sampling run walks the same chains on every rank and reproduces the unsharded run.
Needs >= 2 GPUs (skipped otherwise): `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bayesianinference_b200 import configs as cfg  # noqa: E402

RUN = dict(pool_size=64, batch_k=16, mc_steps=20, max_iter=64, min_iter=64, seed=5)


def _configs():
    return [cfg.c2_polyreg(N=200_001), cfg.c4_gbm(T=5000), cfg.c3_logistic(N=50_000)]


GP_RUN = dict(pool_size=24, batch_k=7, mc_steps=6, max_iter=21, min_iter=21, seed=8)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # host channel for the communicator id only
    from bayesianinference_b200 import engine
    engine.init(device=rank)
    res = {}
    for mode in ("peer", "nccl"):  # in-kernel exchange over peer-mapped memory, then the ncclAllGather fallback
        if mode == "nccl":
            os.environ["BINEST_XCHG"] = "nccl"
        comm = engine.Comm(rank, world)
        out = []
        for c in _configs():
            p = engine.Problem.from_config(c, comm=comm)
            th = p.sample_prior(40, seed=11)
            th[3, -1] = c.lo[-1] - 1.0  # outside the box -> logzero
            ll = p.loglike(th)
            run = engine.RunGroup(p, engine.default_options(**RUN))
            path = run.walk_path()
            run.advance(0)
            s = run.fetch(0)
            out.append((c.name, th, ll, s["logL"], s["points"], s["crude_logZ"], path))
            run.close()
            p.close()
        # batch-sharded GP: ragged batch (41 over 2 ranks), then a short walk
        cg = cfg.c5_gp(N=200)
        p = engine.Problem.from_config(cg, comm=comm, shard="batch")
        thg = p.sample_prior(41, seed=12)
        thg[5, 1] = -1.0
        llg = p.loglike(thg)
        run = engine.RunGroup(p, engine.default_options(**GP_RUN), p.sample_prior(24, seed=13))
        run.advance(0)
        sg = run.fetch(0)
        run.close()
        p.close()
        res[mode] = (out, (thg, llg, sg["logL"], sg["points"]), comm.stats())
        dist.barrier()
        comm.close()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_data_sharded_two_gpus_match_oracle_and_unsharded():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as tmp
    from bayesianinference_b200 import engine
    from oracle import oracle as O
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    allres = dict(q.get(timeout=900) for _ in procs)
    [p.join(120) for p in procs]
    engine.init(device=0)
    # the exchange really ran inside our kernels over peer-mapped memory (and through NCCL when asked to)
    st = allres[0]["peer"][2]
    assert st["peer_path"] is True, "CUDA IPC peer mapping failed on this box: the in-kernel exchange did not run"
    assert allres[0]["nccl"][2]["peer_path"] is False
    assert st["exchanges"] > 3 * 64 // 16 * 20 and st["bytes_pushed"] == allres[1]["peer"][2]["bytes_pushed"] > 0
    for mode in ("peer", "nccl"):
        res = {r: allres[r][mode][0] for r in (0, 1)}
        _check_data_sharded(res, engine, O)
        # the two exchange paths carry the same numbers
        for a, b in zip(allres[0]["peer"][0], allres[0][mode][0]):
            assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
        assert allres[0][mode][0][0][6] == "stepped-sharded"
        # batch-sharded GP: identical on both ranks, equal to the unsharded evaluation and walk
        thg, ll0, L0, pts0 = allres[0][mode][1]
        _, ll1, L1, pts1 = allres[1][mode][1]
        assert np.array_equal(ll0, ll1) and np.array_equal(L0, L1) and np.array_equal(pts0, pts1)
        cg = cfg.c5_gp(N=200)
        p = engine.Problem.from_config(cg)
        np.testing.assert_array_equal(p.loglike(thg), ll0)  # same kernels on the same matrices: bit-identical
        assert ll0[5] == engine.LOGZERO
        run = engine.RunGroup(p, engine.default_options(**GP_RUN), p.sample_prior(24, seed=13))
        run.advance(0)
        s = run.fetch(0)
        np.testing.assert_array_equal(s["logL"], L0)
        np.testing.assert_array_equal(s["points"], pts0)


def _check_data_sharded(res, engine, O):
    for i, c in enumerate(_configs()):
        name, th, ll0, L0, pts0, z0, _ = res[0][i]
        _, _, ll1, L1, pts1, z1, _ = res[1][i]
        # every rank holds the same numbers, bit for bit
        assert np.array_equal(ll0, ll1) and np.array_equal(L0, L1) and np.array_equal(pts0, pts1) and z0 == z1
        op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
        hi, lo = op.loglike_quad(th)
        ok = np.arange(40) != 3
        rel = np.abs((ll0 - hi) - lo)[ok] / np.abs(hi[ok])
        assert rel.max() < 1e-12, (name, rel.max())
        assert ll0[3] == engine.LOGZERO
        # the unsharded engine on one GPU: same seed -> same chains (sums differ only in association)
        p = engine.Problem.from_config(c)
        run = engine.RunGroup(p, engine.default_options(**RUN))
        run.advance(0)
        s = run.fetch(0)
        assert s["logL"].size == L0.size, name
        np.testing.assert_allclose(L0, s["logL"], rtol=1e-11)
        np.testing.assert_allclose(pts0, s["points"], rtol=1e-11)
        assert abs(z0 - s["crude_logZ"]) < 1e-9 * abs(z0)
