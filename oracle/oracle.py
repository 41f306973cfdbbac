"""ctypes front end of the CPU oracle (oracle/binest_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.  PARITY UNPINNED by the
reference (pure Wolfram Language, no tests/fixtures, no Wolfram Engine here): every pin is
constructed; see tests/test_oracle_pins.py.

Reference citations use BS = BayesianStatistics.wl, BU = BayesianUtilities.wl (under
/root/reference/BayesianInference/Kernel/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbinest_oracle.so")

OP_GAUSSIAN_IID, OP_POLYREG, OP_LOGISTIC, OP_GBM, OP_GP_SE = 1, 2, 3, 4, 5
PRIOR_UNIFORM, PRIOR_SCALE, PRIOR_NORMAL_TRUNC = 1, 2, 3
LOGZERO = -1.7976931348623157e308  # -$MaxMachineNumber; BU:47 passes -MachineInfinity


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "binest_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        dp, ip, u32p = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint32)
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_uniform2.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, dp]
        L.orc_normal2.argtypes = L.orc_uniform2.argtypes
        for f in (L.orc_logsubtract, L.orc_logadd):
            f.argtypes = [C.c_double, C.c_double]
            f.restype = C.c_double
        L.orc_logsumexp.argtypes = [dp, C.c_int64]
        L.orc_logsumexp.restype = C.c_double
        L.orc_xvalues_log.argtypes = [C.c_int64, C.c_int64, dp]
        L.orc_xvalues_log_pool.argtypes = [C.c_int64, C.c_int64, ip, dp]
        L.orc_trapezoid_log.argtypes = [dp, C.c_int64, dp]
        L.orc_entropy.argtypes = [dp, dp, C.c_int64, C.c_double]
        L.orc_entropy.restype = C.c_double
        L.orc_problem_create.restype = C.c_void_p
        L.orc_problem_create.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, dp, dp,
                                         C.POINTER(C.c_int), C.c_double]
        L.orc_problem_free.argtypes = [C.c_void_p]
        L.orc_prior_create.restype = C.c_void_p
        L.orc_prior_create.argtypes = [C.c_int, C.POINTER(C.c_int), dp, dp, dp, dp, C.c_double]
        L.orc_prior_free.argtypes = [C.c_void_p]
        L.orc_loglike.argtypes = [C.c_void_p, C.c_void_p, dp]
        L.orc_loglike.restype = C.c_double
        L.orc_loglike_q.argtypes = [C.c_void_p, dp, dp, dp]
        L.orc_loglike_batch.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_int64, dp, C.c_int]
        L.orc_logprior_batch.argtypes = [C.c_void_p, dp, C.c_int64, dp]
        L.orc_predictive_components.restype = C.c_int
        L.orc_predictive_components.argtypes = [C.c_void_p, dp, C.c_int64, dp, C.c_int64, dp]
        L.orc_mcmc_chain.restype = C.c_int
        L.orc_mcmc_chain.argtypes = [C.c_void_p, C.c_void_p, dp, dp, C.c_int64, C.c_uint64, C.c_uint32, C.c_int64, dp, dp, dp,
                                     ip, ip]
        L.orc_gp_predict.restype = C.c_int
        L.orc_gp_predict.argtypes = [C.c_void_p, dp, dp, C.c_int64, C.c_int, dp, dp]
        L.orc_sample_prior.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, dp]
        L.orc_nested_sampling.restype = C.c_void_p
        L.orc_nested_sampling.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, dp]
        L.orc_run_sizes.argtypes = [C.c_void_p, ip, ip, ip, ip]
        L.orc_run_fetch.argtypes = [C.c_void_p, dp, dp, dp, dp, ip, dp, dp, dp]
        L.orc_run_free.argtypes = [C.c_void_p]
        L.orc_evidence_sampling.argtypes = [C.c_int64, C.c_int, dp, dp, ip, C.c_int64, C.c_int64,
                                            C.c_uint64, C.c_int, dp, dp, dp, dp, dp, dp, dp]
        L.orc_bench_walks.restype = C.c_int64
        L.orc_bench_walks.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_int64, C.c_double, C.c_int64,
                                      C.c_int64, C.c_uint64, C.c_int, dp]
        L.orc_max_threads.restype = C.c_int
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- RNG
def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def uniform2(seed, c0, c1, c2, tag, run_id=0):
    o = (C.c_double * 2)()
    lib().orc_uniform2(seed, c0, c1, c2, tag, run_id, o)
    return float(o[0]), float(o[1])


def normal2(seed, c0, c1, c2, tag, run_id=0):
    o = (C.c_double * 2)()
    lib().orc_normal2(seed, c0, c1, c2, tag, run_id, o)
    return float(o[0]), float(o[1])


# ---------------------------------------------------------------- log-space helpers (BU:318-356)
def logsumexp(v):
    v = _f64(v)
    return float(lib().orc_logsumexp(_dp(v), v.size))


def logadd(a, b):
    return float(lib().orc_logadd(a, b))


def logsubtract(a, b):
    return float(lib().orc_logsubtract(a, b))


def xvalues_log(n, n_deleted, pool=None):
    out = np.empty(n + n_deleted)
    if pool is None:
        lib().orc_xvalues_log(n, n_deleted, _dp(out))  # BS:785-799
    else:
        pool = np.ascontiguousarray(pool, dtype=np.int64)
        lib().orc_xvalues_log_pool(n, n_deleted, _ip(pool), _dp(out))
    return out


def trapezoid_log(logx):
    logx = _f64(logx)
    out = np.empty_like(logx)
    lib().orc_trapezoid_log(_dp(logx), logx.size, _dp(out))  # BS:756-771
    return out


def entropy(crude_logw, logL, logZ):
    a, b = _f64(crude_logw), _f64(logL)
    return float(lib().orc_entropy(_dp(a), _dp(b), a.size, logZ))  # BS:801-810


# ---------------------------------------------------------------- problem / prior
class Prior:
    """Box + product prior: BS:25-64, BS:327-427."""

    def __init__(self, kinds, lo, hi, p0=None, p1=None, logzero=LOGZERO):
        self.d = len(kinds)
        self.kinds = np.ascontiguousarray(kinds, dtype=np.int32)
        self.lo, self.hi = _f64(lo), _f64(hi)
        self.p0 = _f64(p0 if p0 is not None else np.zeros(self.d))
        self.p1 = _f64(p1 if p1 is not None else np.ones(self.d))
        self.h = lib().orc_prior_create(self.d, self.kinds.ctypes.data_as(C.POINTER(C.c_int)),
                                        _dp(self.lo), _dp(self.hi), _dp(self.p0), _dp(self.p1), logzero)

    def logpdf(self, theta):
        theta = _f64(np.atleast_2d(theta))
        out = np.empty(theta.shape[0])
        lib().orc_logprior_batch(self.h, _dp(theta), theta.shape[0], _dp(out))
        return out

    def sample(self, n, seed, run_id=0):
        out = np.empty((n, self.d))
        lib().orc_sample_prior(self.h, n, seed, run_id, _dp(out))
        return out

    def __del__(self):
        try:
            lib().orc_prior_free(self.h)
        except Exception:
            pass


class Problem:
    """Likelihood operator bound to its data: BS:429-505, BS:517-595, GP:27-199."""

    def __init__(self, op, d, inputs, outputs=None, iparam=(0, 0, 0, 0), logzero=LOGZERO):
        self.op, self.d = op, d
        self.inputs = _f64(np.atleast_2d(np.asarray(inputs).T).T if np.ndim(inputs) == 1 else inputs)
        n = self.inputs.shape[0]
        self.outputs = None if outputs is None else _f64(np.asarray(outputs).reshape(n, -1))
        ip = (C.c_int * 4)(*[int(v) for v in iparam])
        self.h = lib().orc_problem_create(op, d, n, self.inputs.shape[1],
                                          0 if self.outputs is None else self.outputs.shape[1],
                                          _dp(self.inputs), _dp(self.outputs), ip, logzero)

    def loglike(self, theta, prior: Prior | None = None, threads=1):
        theta = _f64(np.atleast_2d(theta))
        out = np.empty(theta.shape[0])
        lib().orc_loglike_batch(self.h, prior.h if prior else None, _dp(theta), theta.shape[0], _dp(out), threads)
        return out

    def predictive_components(self, theta, inputs):
        """BS:1437-1483: (M, Q, C) component parameters; C = 2 (mean, sd) or K class probabilities."""
        theta = _f64(np.atleast_2d(theta))
        inputs = _f64(np.asarray(inputs, dtype=np.float64).reshape(-1, self.inputs.shape[1]))
        M, Q = theta.shape[0], inputs.shape[0]
        Cw = {2: 2, 3: None}.get(self.op, 0)
        if Cw is None:
            Cw = (theta.shape[1] // (self.inputs.shape[1] + 1)) + 1
        if not Cw:
            raise ValueError("operator has no independent variables")
        out = np.empty((M, Q, Cw))
        got = lib().orc_predictive_components(self.h, _dp(theta), M, _dp(inputs), Q, _dp(out))
        assert got == Cw, (got, Cw)
        return out

    def gp_predict(self, theta, xstar, long_double=False):
        """predictFromGaussianProcessInternal (GP:395-420): (mean, sd), each M x Q; NaN rows where K is singular."""
        theta = _f64(np.atleast_2d(theta))
        xstar = _f64(np.asarray(xstar, dtype=np.float64).reshape(-1, self.inputs.shape[1]))
        M, Q = theta.shape[0], xstar.shape[0]
        mean, sd = np.full((M, Q), np.nan), np.full((M, Q), np.nan)
        for i in range(M):
            row = np.ascontiguousarray(theta[i])
            m, s = np.empty(Q), np.empty(Q)
            if lib().orc_gp_predict(self.h, _dp(row), _dp(xstar), Q, int(long_double), _dp(m), _dp(s)):
                mean[i], sd[i] = m, s
        return mean, sd

    def loglike_quad(self, theta):
        """__float128 value of the sum, returned as (hi, lo) double-double."""
        theta = _f64(np.atleast_2d(theta))
        hi, lo = np.empty(theta.shape[0]), np.empty(theta.shape[0])
        a, b = C.c_double(), C.c_double()
        for i in range(theta.shape[0]):
            row = np.ascontiguousarray(theta[i])
            lib().orc_loglike_q(self.h, _dp(row), C.byref(a), C.byref(b))
            hi[i], lo[i] = a.value, b.value
        return hi, lo

    def __del__(self):
        try:
            lib().orc_problem_free(self.h)
        except Exception:
            pass


class _Options(C.Structure):
    _fields_ = [("pool_size", C.c_int64), ("batch_k", C.c_int64), ("mc_steps", C.c_int64),
                ("max_iter", C.c_int64), ("min_iter", C.c_int64), ("term_frac", C.c_double),
                ("acc_min", C.c_double), ("acc_max", C.c_double), ("seed", C.c_uint64),
                ("run_id", C.c_int64), ("adapt_in_walk", C.c_int64), ("loglmax", C.c_double)]


@dataclass
class RunResult:
    points: np.ndarray
    logL: np.ndarray
    logPrior: np.ndarray
    acc: np.ndarray
    pool: np.ndarray
    logX: np.ndarray
    crude_logw: np.ndarray
    crude_logZ: float
    entropy: float
    logLmax: float
    n: int
    n_deleted: int
    iterations: int
    evals: int


def nested_sampling(problem: Problem, prior: Prior, pool_size=100, batch_k=1, mc_steps=200,
                    max_iter=10000, min_iter=100, term_frac=0.01, acc_range=(0.0, 1.0), seed=1,
                    run_id=0, adapt_in_walk=True, start_points=None, loglmax=float("nan")) -> RunResult:
    """BS:859-1040 (+ BS:707-745 walk protocol), sequential."""
    o = _Options(pool_size, batch_k, mc_steps, max_iter, min_iter, term_frac, acc_range[0], acc_range[1],
                 seed, run_id, 1 if adapt_in_walk else 0, loglmax)
    sp = None if start_points is None else _f64(start_points)
    h = lib().orc_nested_sampling(problem.h, prior.h, C.byref(o), _dp(sp))
    M, nd, it, ev = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    lib().orc_run_sizes(h, C.byref(M), C.byref(nd), C.byref(it), C.byref(ev))
    M_ = M.value
    pts = np.empty((M_, problem.d))
    logL, logPr, acc, logX, lw = (np.empty(M_) for _ in range(5))
    pool = np.empty(M_, dtype=np.int64)
    summ = np.empty(3)
    lib().orc_run_fetch(h, _dp(pts), _dp(logL), _dp(logPr), _dp(acc), _ip(pool), _dp(logX), _dp(lw), _dp(summ))
    lib().orc_run_free(h)
    return RunResult(pts, logL, logPr, acc, pool, logX, lw, float(summ[0]), float(summ[1]), float(summ[2]),
                     pool_size, nd.value, it.value, ev.value)


def evidence_sampling(points, logL, pool, n, nruns=100, seed=1, sorted_draws=False):
    """BS:1158-1291 on a sorted sample list.  Returns a dict mirroring the reference keys."""
    points, logL = _f64(points), _f64(logL)
    pool = np.ascontiguousarray(pool, dtype=np.int64)
    M, d = points.shape
    z, H = np.empty(nruns), np.empty(nruns)
    lwm, lws, sxm, sxs = (np.empty(M) for _ in range(4))
    pm = np.empty((nruns, d))
    lib().orc_evidence_sampling(M, d, _dp(points), _dp(logL), _ip(pool), n, nruns, seed, int(sorted_draws), _dp(z), _dp(lwm),
                                _dp(lws), _dp(sxm), _dp(sxs), _dp(pm), _dp(H))
    return {
        "zSamples": z,
        "LogEvidence": {"Mean": float(z.mean()), "StandardError": float(z.std(ddof=1))},  # BS:1254, 1142
        "LogPosteriorWeight": {"Mean": lwm, "StandardError": lws},
        "SampledLogX": {"Mean": sxm, "StandardError": sxs},
        "ParameterExpectedValues": {"Mean": pm.mean(0), "StandardError": pm.std(0, ddof=1)},
        "RelativeEntropy": {"Mean": float(H.mean()), "StandardError": float(H.std(ddof=1))},
        "parameterSamples": pm,
        "HSamples": H,
    }


def combine_runs(runs):
    """BS:1293-1315: join, DeleteDuplicatesBy Point (first occurrence kept), SortBy {logL, Point};
    SamplePoolSize = sum of pool sizes.  Per-sample pool sizes (batched replacement) are summed
    across runs at each sample's likelihood level; with constant pools this is the reference's
    constant sum n_r."""
    pts = np.concatenate([r.points for r in runs])
    logL = np.concatenate([r.logL for r in runs])
    logPr = np.concatenate([r.logPrior for r in runs])
    acc = np.concatenate([r.acc for r in runs])
    rid = np.concatenate([np.full(r.logL.size, i) for i, r in enumerate(runs)])
    _, first = np.unique(pts, axis=0, return_index=True)
    keep = np.sort(first)
    pts, logL, logPr, acc, rid = pts[keep], logL[keep], logPr[keep], acc[keep], rid[keep]
    order = np.lexsort(tuple(pts[:, j] for j in range(pts.shape[1] - 1, -1, -1)) + (logL,))
    pts, logL, logPr, acc, rid = pts[order], logL[order], logPr[order], acc[order], rid[order]
    n_tot = int(sum(r.n for r in runs))
    M = logL.size
    pool = np.zeros(M, dtype=np.int64)
    for i, r in enumerate(runs):
        # pool size of run i at level L: pool of its first sample with logL >= L (0 when exhausted)
        idx = np.searchsorted(r.logL, logL, side="left")
        contrib = np.where(idx < r.logL.size, r.pool[np.minimum(idx, r.logL.size - 1)], 0)
        pool += contrib
    return dict(points=pts, logL=logL, logPrior=logPr, acc=acc, run_id=rid, pool=pool, n=n_tot,
                n_deleted=M - n_tot)


_fast = None
FAST_FLAGS = None


def fast_lib():
    """The timed CPU arm (bench.py only): the same source built -O3 with vectorised sums (oracle/Makefile).  Tries
    -march=native on the machine being timed (into a private temp dir), else the prebuilt x86-64-v3 library."""
    global _fast, FAST_FLAGS
    if _fast is None:
        import shutil
        import tempfile
        path, arch = os.path.join(_HERE, "_build", "libbinest_oracle_fast.so"), "x86-64-v3"
        if shutil.which("gcc") or os.path.exists("/usr/bin/gcc"):
            try:
                tmp = os.path.join(tempfile.mkdtemp(prefix="binest_oracle_"), "libbinest_oracle_fast.so")
                subprocess.check_call(["make", "-s", "-C", _HERE, tmp, f"FAST_OUT={tmp}", "FAST_ARCH=native"],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                path, arch = tmp, "native"
            except Exception:
                pass
        if not os.path.exists(path):
            build(force=True)
        _fast = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        _fast.orc_bench_walks.restype = C.c_int64
        _fast.orc_bench_walks.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_int64, C.c_double, C.c_int64,
                                          C.c_int64, C.c_uint64, C.c_int, dp]
        FAST_FLAGS = f"gcc -O3 -march={arch} -ffast-math -fopenmp, omp simd sums"
    return _fast


def bench_walks(problem: Problem, prior: Prior, start_points, Lstar, reps_per_thread, S, seed, threads, fast=True):
    """Timed CPU leg: returns the number of likelihood evaluations PERFORMED (proposals outside the box are rejected
    without one, as nsDensity BS:602-617 does).  The handles come from the parity library; both builds share the
    struct layout (same source)."""
    sp = _f64(start_points)
    sink = np.empty(max(threads, 1))
    L = fast_lib() if fast else lib()
    return int(L.orc_bench_walks(problem.h, prior.h, _dp(sp), sp.shape[0], Lstar, reps_per_thread, S,
                                 seed, threads, _dp(sink)))


def max_threads():
    return int(lib().orc_max_threads())


def mcmc_chain(problem, prior, start, init_cov, delay=20, seed=1, chain_id=0, n_steps=100):
    """createMCMCChain + iterateMCMC (BS:630-703) restated: states (n_steps, d) and the final chain estimates."""
    d = problem.d
    start = _f64(np.asarray(start, dtype=np.float64).reshape(d))
    init_cov = _f64(np.asarray(init_cov, dtype=np.float64).reshape(d, d))
    out, mean, cov = np.empty((n_steps, d)), np.empty(d), np.empty((d, d))
    t, acc = C.c_int64(), C.c_int64()
    ok = lib().orc_mcmc_chain(problem.h, prior.h, _dp(start), _dp(init_cov), delay, seed, chain_id, n_steps, _dp(out), _dp(mean),
                              _dp(cov), C.byref(t), C.byref(acc))
    if not ok:
        raise ValueError("bad starting point or InitialCovariance")
    return dict(states=out, mean=mean, cov=cov, t=t.value, accepted=acc.value)
