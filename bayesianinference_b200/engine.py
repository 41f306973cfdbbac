"""Thin object layer over the C ABI (include/binest.h): Problem = operator + resident data + prior,
RunGroup = one or more lock-step nested-sampling runs.  No numerics live here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import LOGZERO, Options, check, dptr, iptr

_initialised = False


def init(logzero: float = LOGZERO, device: int = -1):
    """binest_init: fails loudly without a B200 (no CPU fallback)."""
    global _initialised
    check(_lib.load().binest_init(logzero, device))
    _initialised = True


def _ensure_init():
    if not _initialised:
        init()


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_options(**kw) -> Options:
    o = Options()
    _lib.load().binest_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


def shard_rows(n_rows: int, rank: int, world: int, overlap: int = 0):
    """Row range [lo, hi) of `rank` in the data-sharded mode: `world` contiguous blocks whose sizes differ by at
    most one.  overlap = 1 for point series whose rows are increments (GBM): shard r > 0 starts one point early, so
    the increment across the cut belongs to the later shard and every increment is counted exactly once."""
    if not (0 <= rank < world):
        raise ValueError("0 <= rank < world")
    units = n_rows - overlap  # rows (or increments) to distribute
    if units < world:
        raise ValueError(f"cannot split {units} rows over {world} ranks")
    base, extra = divmod(units, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi + overlap


class Comm:
    """Communicator of the data-sharded mode (binest_comm_*): one per process, NCCL underneath.

    `exchange(obj_or_None)` must hand rank 0's bytes to every rank; with torch.distributed initialised the default
    uses broadcast_object_list (any backend)."""

    def __init__(self, rank: int, world: int, exchange=None):
        _ensure_init()
        L = _lib.load()
        ident = (C.c_uint8 * _lib.COMM_ID_BYTES)()
        if rank == 0:
            check(L.binest_comm_unique_id(ident))
        payload = bytes(ident) if rank == 0 else None
        if exchange is None:
            import torch.distributed as dist

            def exchange(b):
                box = [b]
                dist.broadcast_object_list(box, src=0)
                return box[0]
        payload = exchange(payload)
        ident = (C.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(payload)
        h = C.c_void_p()
        check(L.binest_comm_create(rank, world, ident, C.byref(h)))
        self.h, self.rank, self.world = h, rank, world

    def stats(self):
        """Exchange traffic so far: exchanges issued, payload bytes this rank pushed to its peers, and whether the
        exchange runs in-kernel over peer-mapped memory (True) or through ncclAllGather (False)."""
        ex, by, pp = C.c_int64(), C.c_int64(), C.c_int()
        check(_lib.load().binest_comm_stats(self.h, C.byref(ex), C.byref(by), C.byref(pp)))
        return dict(exchanges=ex.value, bytes_pushed=by.value, peer_path=bool(pp.value))

    def close(self):
        if getattr(self, "h", None):
            _lib.load().binest_comm_free(self.h)
            self.h = None


class Problem:
    """Device-resident inference problem (the data-carrying half of defineInferenceProblem, BS:167-307).

    comm: sharded modes — `inputs`/`outputs` are the FULL data on every rank.
      shard="rows"  (data-sharded) only this rank's row block (shard_rows) is uploaded and the problem is declared a
                    shard (binest_problem_shard, collective);
      shard="batch" (GP operator) the data are replicated and every theta batch is split across the ranks
                    (binest_problem_shard_batch)."""

    def __init__(self, op, inputs, outputs, iparam, kinds, lo, hi, p0=None, p1=None, comm=None, shard="rows"):
        _ensure_init()
        L = _lib.load()
        self.op = int(op)
        inputs = _f64(inputs)
        if inputs.ndim == 1:
            inputs = inputs.reshape(-1, 1)
        n = inputs.shape[0]
        outputs = None if outputs is None else _f64(outputs).reshape(n, -1)
        if shard not in ("rows", "batch"):
            raise ValueError("shard must be 'rows' or 'batch'")
        if comm is not None and shard == "rows":
            from .configs import OP_GBM
            r0, r1 = shard_rows(n, comm.rank, comm.world, overlap=1 if self.op == OP_GBM else 0)
            inputs = np.ascontiguousarray(inputs[r0:r1])
            outputs = None if outputs is None else np.ascontiguousarray(outputs[r0:r1])
            n = r1 - r0
        self.comm = comm
        self.d = len(kinds)
        ip = np.ascontiguousarray(list(iparam) + [0] * (4 - len(iparam)), dtype=np.int64)
        kinds = np.ascontiguousarray(kinds, dtype=np.int32)
        lo, hi = _f64(lo), _f64(hi)
        p0 = None if p0 is None or len(p0) == 0 else _f64(p0)
        p1 = None if p1 is None or len(p1) == 0 else _f64(p1)
        h = C.c_void_p()
        check(L.binest_problem_create(self.op, iptr(ip), dptr(inputs), n, inputs.shape[1], dptr(outputs),
                                      0 if outputs is None else outputs.shape[1], self.d,
                                      kinds.ctypes.data_as(C.POINTER(C.c_int32)), dptr(lo), dptr(hi), dptr(p0),
                                      dptr(p1), C.byref(h)))
        self.h = h
        self.n_rows = n
        if comm is not None:
            check((L.binest_problem_shard if shard == "rows" else L.binest_problem_shard_batch)(self.h, comm.h))

    @classmethod
    def from_config(cls, cfg, comm=None, shard="rows"):
        return cls(cfg.op, cfg.inputs, cfg.outputs, cfg.iparam, cfg.kinds, cfg.lo, cfg.hi, cfg.p0, cfg.p1, comm=comm,
                   shard=shard)

    def loglike(self, theta):
        """"LogLikelihoodFunction" (Listable): theta (P, d) -> (P,)."""
        theta = _f64(np.atleast_2d(theta))
        if theta.shape[1] != self.d:
            raise ValueError(f"theta must have {self.d} columns")
        out = np.empty(theta.shape[0])
        check(_lib.load().binest_loglike(self.h, dptr(theta), theta.shape[0], dptr(out)))
        return out

    def predictive_components(self, theta, inputs):
        """predictiveDistribution (BS:1437-1483): parameters of the mixture components dist[theta_m, x_q] for every
        sample and input -> (M, Q, C); C = 2 (mean, sd) for polynomial regression, K class probabilities for softmax."""
        theta = _f64(np.atleast_2d(theta))
        if theta.shape[1] != self.d:
            raise ValueError(f"theta must have {self.d} columns")
        inputs = np.asarray(inputs, dtype=np.float64)
        inputs = _f64(inputs.reshape(-1, 1) if inputs.ndim == 1 else inputs)
        w = C.c_int64()
        check(_lib.load().binest_predictive_width(self.h, C.byref(w)))
        if w.value == 0:
            raise ValueError("operator has no independent variables")
        out = np.empty((theta.shape[0], inputs.shape[0], w.value))
        check(_lib.load().binest_predictive_components(self.h, dptr(theta), theta.shape[0], dptr(inputs), inputs.shape[0],
                                                       dptr(out)))
        return out

    def gp_predict(self, theta, xstar):
        """predictFromGaussianProcessInternal (GP:395-420) for every parameter vector: theta (M, 3), xstar (Q, D)
        -> (mean, sd), each (M, Q): the NormalDistribution parameters of the predictive at every input."""
        theta = _f64(np.atleast_2d(theta))
        if theta.shape[1] != self.d:
            raise ValueError(f"theta must have {self.d} columns")
        xstar = _f64(np.asarray(xstar, dtype=np.float64))
        xstar = _f64(xstar.reshape(-1, 1) if xstar.ndim == 1 else xstar)
        M, Q = theta.shape[0], xstar.shape[0]
        mean, sd = np.empty((M, Q)), np.empty((M, Q))
        check(_lib.load().binest_gp_predict(self.h, dptr(theta), M, dptr(xstar), Q, dptr(mean), dptr(sd)))
        return mean, sd

    def logprior(self, theta):
        """"LogPriorPDFFunction"."""
        theta = _f64(np.atleast_2d(theta))
        out = np.empty(theta.shape[0])
        check(_lib.load().binest_logprior(self.h, dptr(theta), theta.shape[0], dptr(out)))
        return out

    def sample_prior(self, n, seed=1, run_id=0):
        """generateStartingPoints (BS:1055-1068)."""
        out = np.empty((n, self.d))
        check(_lib.load().binest_sample_prior(self.h, n, seed, run_id, dptr(out)))
        return out

    def stream(self) -> int:
        """cudaStream_t (as an integer) all kernels of this problem and of its runs are launched on."""
        st = C.c_void_p()
        check(_lib.load().binest_problem_stream(self.h, C.byref(st)))
        return int(st.value or 0)

    def bench_loglike(self, P, reps=20, warmup=3, flush_l2=True):
        a, b = C.c_double(), C.c_double()
        check(_lib.load().binest_bench_loglike(self.h, P, reps, warmup, 1 if flush_l2 else 0, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if getattr(self, "h", None):
            _lib.load().binest_problem_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RunGroup:
    """n_runs lock-step runs of nestedSamplingInternal (BS:859-1040) on the current device."""

    def __init__(self, problem: Problem, options: Options, start_points=None):
        self.problem = problem
        self.options = options
        sp = None
        if start_points is not None:
            sp = _f64(start_points).reshape(options.n_runs, options.pool_size, problem.d)
        h = C.c_void_p()
        check(_lib.load().binest_run_create(problem.h, C.byref(options), dptr(sp), C.byref(h)))
        self.h = h
        self.n_runs = int(options.n_runs)

    def advance(self, max_batches=0) -> bool:
        fin = C.c_int32()
        check(_lib.load().binest_run_advance(self.h, max_batches, C.byref(fin)))
        return bool(fin.value)

    def sizes(self, run=0):
        M, nd, it, ev = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(_lib.load().binest_run_sizes(self.h, run, C.byref(M), C.byref(nd), C.byref(it), C.byref(ev)))
        return dict(M=M.value, n_deleted=nd.value, iterations=it.value, evals=ev.value)

    def timing(self):
        ms, g, b = C.c_double(), C.c_int64(), C.c_int64()
        check(_lib.load().binest_run_timing(self.h, C.byref(ms), C.byref(g), C.byref(b)))
        return dict(walk_ms=ms.value, walk_graphs=g.value, batches=b.value)

    WALK_PATHS = ("stepped-graph", "cluster-resident", "grid-resident", "stepped-sharded", "stepped-gp", "device-loop")

    def walk_path(self) -> str:
        v = C.c_int()
        check(_lib.load().binest_run_path(self.h, C.byref(v)))
        return self.WALK_PATHS[v.value]

    def fetch(self, run=0, weights=True):
        """Sorted sample list of one run + (weights=True) the calculateWeightsCrude columns (BS:812-831).
        weights=False skips the per-run weight kernel: the caller is going to merge runs and re-weight anyway."""
        L = _lib.load()
        # an unfinished run is flushed by the fetch itself, so sizes are read after a first NULL fetch (flush only)
        check(L.binest_run_fetch(self.h, run, None, None, None, None, None, None, None, None))
        s = self.sizes(run)
        M, d = s["M"], self.problem.d
        pts = np.empty((M, d))
        logL, logPr, acc = (np.empty(M) for _ in range(3))
        logX, lw, summ = (np.empty(M), np.empty(M), np.empty(4)) if weights else (None, None, None)
        pool = np.empty(M, dtype=np.int64)
        check(L.binest_run_fetch(self.h, run, dptr(pts), dptr(logL), dptr(logPr), dptr(acc), iptr(pool), dptr(logX),
                                 dptr(lw), dptr(summ)))
        out = dict(points=pts, logL=logL, logPrior=logPr, acc=acc, pool=pool, n=int(self.options.pool_size), **s)
        if weights:
            out.update(logX=logX, crude_logw=lw, crude_logZ=float(summ[0]), entropy=float(summ[1]),
                       logLmax=float(summ[2]), log_missing=float(summ[3]))
        return out

    def merge(self):
        """combineRuns BS:1293-1297 of the group's runs on the device (binest_run_merge): merged table (summed PoolSize,
        RunIndex = global run ids) and the live-block length.  No per-run fetch."""
        L = _lib.load()
        Mt = C.c_int64()
        check(L.binest_run_merge_size(self.h, C.byref(Mt)))
        Mt, d = Mt.value, self.problem.d
        pts, cols = _host_out((Mt, d)), [_host_out(Mt) for _ in range(3)]
        pool, rid = _host_out(Mt, np.int64), _host_out(Mt, np.int64)
        M, live = C.c_int64(), C.c_int64()
        check(L.binest_run_merge(self.h, dptr(pts), dptr(cols[0]), dptr(cols[1]), dptr(cols[2]), iptr(pool), iptr(rid),
                                 C.byref(M), C.byref(live)))
        m = M.value
        return ({"Point": pts[:m], "LogLikelihood": cols[0][:m], "LogPriorPDF": cols[1][:m], "AcceptanceRate": cols[2][:m],
                 "PoolSize": pool[:m], "RunIndex": rid[:m]}, live.value)

    def combine(self, reference_scheme, post_runs=100, seed=1):
        """combineRuns -> evidenceSampling of the group's runs in one device call (binest_run_combine)."""
        L = _lib.load()
        Mt = C.c_int64()
        check(L.binest_run_merge_size(self.h, C.byref(Mt)))
        Mt, d = Mt.value, self.problem.d
        o_pts, tab, itab = _host_out((Mt, d)), _host_out((len(COLS), Mt)), _host_out((2, Mt), np.int64)
        z, H, pm, summ = np.empty(post_runs), np.empty(post_runs), np.empty((post_runs, d)), np.empty(4)
        M, n_live = C.c_int64(), C.c_int64()
        check(L.binest_run_combine(self.h, 0 if reference_scheme else 1, int(post_runs), int(seed), dptr(o_pts), dptr(tab),
                                   iptr(itab), dptr(z), dptr(pm), dptr(H), dptr(summ), C.byref(M), C.byref(n_live)))
        return _combined(M.value, o_pts, tab, itab, z, H, pm, summ, n_live.value, True, True)

    def merge_dev(self):
        """binest_run_merge_dev: the merge of the group's runs as a packed table IN DEVICE MEMORY (a torch tensor
        [M, d + 5] on this GPU: point, logL, logPrior, acc, pool size, run id) for an NCCL gather across ranks."""
        import torch
        L = _lib.load()
        Mt = C.c_int64()
        check(L.binest_run_merge_size(self.h, C.byref(Mt)))
        t = torch.empty((Mt.value, self.problem.d + 5), dtype=torch.float64, device="cuda")
        M, live = C.c_int64(), C.c_int64()
        check(L.binest_run_merge_dev(self.h, C.c_void_p(t.data_ptr()), C.byref(M), C.byref(live)))
        return t[:M.value], live.value

    def estimates(self, run=0):
        d = self.problem.d
        m, c = np.empty(d), np.empty((d, d))
        check(_lib.load().binest_run_estimates(self.h, run, dptr(m), dptr(c)))
        return m, c

    def close(self):
        if getattr(self, "h", None):
            _lib.load().binest_run_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Chain:
    """createMCMCChain (BS:630-703): adaptive-Metropolis chains on the log posterior of a Problem."""

    def __init__(self, problem: Problem, start, init_cov, learn_delay=20, seed=1):
        self.problem = problem
        start = _f64(np.atleast_2d(np.asarray(start, dtype=np.float64)))
        if start.shape[1] != problem.d:
            raise ValueError(f"starting points must have {problem.d} columns")
        init_cov = _f64(np.asarray(init_cov, dtype=np.float64).reshape(problem.d, problem.d))
        self.n_chains, self.d = start.shape[0], problem.d
        self.h = C.c_void_p()
        check(_lib.load().binest_chain_create(problem.h, dptr(start), self.n_chains, dptr(init_cov), int(learn_delay),
                                              int(seed), C.byref(self.h)))

    def iterate(self, n_steps, record=True):
        """n_steps of every chain -> states (n_steps, n_chains, d) (or None with record=False)."""
        out = np.empty((int(n_steps), self.n_chains, self.d)) if record else None
        check(_lib.load().binest_chain_iterate(self.h, int(n_steps), dptr(out)))
        return out

    def state(self):
        Cn, d = self.n_chains, self.d
        x, lp, mean, cov = np.empty((Cn, d)), np.empty(Cn), np.empty((Cn, d)), np.empty((Cn, d, d))
        t, acc = np.empty(Cn, dtype=np.int64), np.empty(Cn, dtype=np.int64)
        check(_lib.load().binest_chain_state(self.h, dptr(x), dptr(lp), dptr(mean), dptr(cov), iptr(t), iptr(acc)))
        return dict(x=x, logdensity=lp, mean=mean, cov=cov, t=t, accepted=acc)

    def close(self):
        if getattr(self, "h", None):
            _lib.load().binest_chain_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def evidence_sampling(points, logL, pool, n_live, post_runs=100, seed=1):
    """evidenceSampling (BS:1158-1291) on a sorted sample list; raw per-draw outputs."""
    _ensure_init()
    points, logL = _f64(points), _f64(logL)
    pool = np.ascontiguousarray(pool, dtype=np.int64)
    M, d = points.shape
    z, H = np.empty(post_runs), np.empty(post_runs)
    lwm, lws, sxm, sxs = (np.empty(M) for _ in range(4))
    pm = np.empty((post_runs, d))
    check(_lib.load().binest_evidence_sampling(M, d, dptr(points), dptr(logL), iptr(pool), n_live, post_runs, seed,
                                               dptr(z), dptr(lwm), dptr(lws), dptr(sxm), dptr(sxs), dptr(pm), dptr(H)))
    return dict(z=z, H=H, logw_mean=lwm, logw_sd=lws, slx_mean=sxm, slx_sd=sxs, pmean=pm)


def crude_weights(logL, pool, n_live):
    """calculateWeightsCrude + logSumExp + calculateEntropy on a sorted list (BS:812-831, BU:318-335)."""
    _ensure_init()
    logL = _f64(logL)
    pool = np.ascontiguousarray(pool, dtype=np.int64)
    M = logL.size
    lx, lw, s = np.empty(M), np.empty(M), np.empty(4)
    check(_lib.load().binest_crude_weights(M, dptr(logL), iptr(pool), n_live, dptr(lx), dptr(lw), dptr(s)))
    return dict(logX=lx, crude_logw=lw, crude_logZ=float(s[0]), entropy=float(s[1]), logLmax=float(s[2]),
                log_missing=float(s[3]))


def _host_out(shape, dtype=np.float64):
    """Output buffer of a device merge: page-locked when torch is there to provide it (its caching host allocator reuses
    the blocks, so repeated calls neither pin nor page-fault 30 MB of fresh memory: the D2H copy of a merged C4 table
    went from ~9 ms to ~2 ms), plain numpy otherwise.  The array owns (a reference to) its memory either way."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int64): torch.int64}[np.dtype(dtype)]
            return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()
    except Exception:  # no torch / no pinned memory: pageable buffers work the same, only slower
        pass
    return np.empty(shape, dtype=dtype)


def _join_runs(tables):
    """the runs' columns concatenated in Join order (BS:1293) for the device merge"""
    sizes = np.array([t["LogLikelihood"].size for t in tables], dtype=np.int64)
    d = max(np.asarray(t["Point"]).shape[1] for t in tables if np.asarray(t["Point"]).ndim == 2)
    pts = _f64(np.concatenate([np.asarray(t["Point"], float).reshape(t["LogLikelihood"].size, d) for t in tables]))
    col = lambda k: _f64(np.concatenate([t[k] for t in tables])) if all(k in t for t in tables) else None  # noqa: E731
    pool = np.ascontiguousarray(np.concatenate([t["PoolSize"] for t in tables]), dtype=np.int64)
    rid = (np.ascontiguousarray(np.concatenate([t["RunIndex"] for t in tables]), dtype=np.int64)
           if all("RunIndex" in t for t in tables) else None)
    return sizes, pts, col("LogLikelihood"), col("LogPriorPDF"), col("AcceptanceRate"), pool, rid


def merge_runs(tables):
    """combineRuns BS:1293-1297 on the device (binest_merge_runs): tables = the runs' sample tables, each sorted by
    {logL, point} and carrying its own "PoolSize" column.  Returns the merged table and the live-block length."""
    _ensure_init()
    sizes, pts, L, lp, acc, pool, rid = _join_runs(tables)
    Mt, d = pts.shape
    o_pts, o_L, o_lp, o_acc = np.empty((Mt, d)), np.empty(Mt), np.empty(Mt), np.empty(Mt)
    o_pool, o_rid = np.empty(Mt, dtype=np.int64), np.empty(Mt, dtype=np.int64)
    M, live = C.c_int64(), C.c_int64()
    check(_lib.load().binest_merge_runs(len(tables), iptr(sizes), d, dptr(pts), dptr(L), dptr(lp), dptr(acc), iptr(pool),
                                        iptr(rid), dptr(o_pts), dptr(o_L), dptr(o_lp), dptr(o_acc), iptr(o_pool),
                                        iptr(o_rid), C.byref(M), C.byref(live)))
    m = M.value
    out = {"Point": o_pts[:m], "PoolSize": o_pool[:m], "RunIndex": o_rid[:m], "LogLikelihood": o_L[:m]}
    if lp is not None:
        out["LogPriorPDF"] = o_lp[:m]
    if acc is not None:
        out["AcceptanceRate"] = o_acc[:m]
    return out, live.value


COLS = ("LogLikelihood", "LogPriorPDF", "AcceptanceRate", "LogX", "X", "CrudeLogPosteriorWeight", "CrudePosteriorWeight",
        "SampledLogX.Mean", "SampledLogX.StandardError", "LogPosteriorWeight.Mean", "LogPosteriorWeight.StandardError")


def combine_runs(tables, reference_scheme, n_tot, post_runs=100, seed=1):
    """combineRuns -> evidenceSampling (BS:1293-1315, 1158-1291) in one device call (binest_combine_runs): merge, X
    sequence, crude weights, Monte-Carlo evidence error, final sort by posterior weight.  Returns the finished sample
    table (columns as api.evidenceSampling builds them) and the per-draw / summary outputs."""
    _ensure_init()
    sizes, pts, L, lp, acc, pool, rid = _join_runs(tables)
    Mt, d = pts.shape
    o_pts, tab, itab = _host_out((Mt, d)), _host_out((len(COLS), Mt)), _host_out((2, Mt), np.int64)
    z, H, pm, summ = np.empty(post_runs), np.empty(post_runs), np.empty((post_runs, d)), np.empty(4)
    M, n_live = C.c_int64(), C.c_int64()
    check(_lib.load().binest_combine_runs(len(tables), iptr(sizes), d, dptr(pts), dptr(L), dptr(lp), dptr(acc), iptr(pool),
                                          iptr(rid), 0 if reference_scheme else 1, int(n_tot), int(post_runs), int(seed),
                                          dptr(o_pts), dptr(tab), iptr(itab), dptr(z), dptr(pm), dptr(H), dptr(summ),
                                          C.byref(M), C.byref(n_live)))
    return _combined(M.value, o_pts, tab, itab, z, H, pm, summ, n_live.value, lp is not None, acc is not None)


def combine_runs_dev(tables, reference_scheme, n_tot, post_runs=100, seed=1):
    """binest_combine_runs_dev: as combine_runs, the inputs being packed device tables (torch tensors [M_r, d + 5] on this
    GPU, e.g. the NCCL-gathered RunGroup.merge_dev() of every rank, in rank order)."""
    import torch
    _ensure_init()
    tables = [t for t in tables if t.shape[0] > 0]
    sizes = np.array([t.shape[0] for t in tables], dtype=np.int64)
    packed = torch.cat(tables).contiguous()
    Mt, d = int(sizes.sum()), packed.shape[1] - 5
    o_pts, tab, itab = _host_out((Mt, d)), _host_out((len(COLS), Mt)), _host_out((2, Mt), np.int64)
    z, H, pm, summ = np.empty(post_runs), np.empty(post_runs), np.empty((post_runs, d)), np.empty(4)
    M, n_live = C.c_int64(), C.c_int64()
    torch.cuda.current_stream().synchronize()  # the gathered tensor is complete before the library's streams read it
    check(_lib.load().binest_combine_runs_dev(len(tables), iptr(sizes), d, C.c_void_p(packed.data_ptr()),
                                              0 if reference_scheme else 1, int(n_tot), int(post_runs), int(seed), dptr(o_pts),
                                              dptr(tab), iptr(itab), dptr(z), dptr(pm), dptr(H), dptr(summ), C.byref(M),
                                              C.byref(n_live)))
    return _combined(M.value, o_pts, tab, itab, z, H, pm, summ, n_live.value, True, True)


def _combined(m, o_pts, tab, itab, z, H, pm, summ, n_live, has_lp, has_acc):
    S = {"Point": o_pts[:m], "PoolSize": itab[0, :m], "RunIndex": itab[1, :m]}
    for c, name in enumerate(COLS):
        if name == "LogPriorPDF" and not has_lp or name == "AcceptanceRate" and not has_acc:
            continue
        if "." in name:
            k, sub = name.split(".")
            S.setdefault(k, {})[sub] = tab[c, :m]
        else:
            S[name] = tab[c, :m]
    return dict(Samples=S, z=z, H=H, pmean=pm, crude_logZ=float(summ[0]), entropy=float(summ[1]), logLmax=float(summ[2]),
                log_missing=float(summ[3]), n_live=n_live)


def fp64_peak():
    _ensure_init()
    t, ms = C.c_double(), C.c_double()
    check(_lib.load().binest_measure_fp64_peak(C.byref(t), C.byref(ms)))
    return t.value


def launch_count():
    return int(_lib.load().binest_launch_count())
