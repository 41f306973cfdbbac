"""Test-only backend: the same interface as bayesianinference_b200.engine, served by the CPU oracle.
It lets the host-side logic of bayesianinference_b200.api (option handling, association assembly,
combineRuns, run sharding under torch.distributed/gloo) be exercised without a GPU.  Never imported by the
product package."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from oracle import oracle as O


def default_options(**kw):
    o = SimpleNamespace(pool_size=100, batch_k=1, mc_steps=200, max_iter=10000, min_iter=100, term_frac=0.01,
                        acc_min=0.0, acc_max=1.0, seed=1, first_run_id=0, n_runs=1, loglmax=float("nan"))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(k)
        setattr(o, k, v)
    return o


class Comm:
    """Data-sharded mode on CPU: the per-rank sums travel over torch.distributed (gloo) instead of NCCL."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def allgather(self, x):
        import torch
        import torch.distributed as dist
        out = [torch.empty(x.shape, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(out, torch.from_numpy(np.ascontiguousarray(x)))
        return np.stack([t.numpy() for t in out])


class Problem:
    def __init__(self, op, inputs, outputs, iparam, kinds, lo, hi, p0=None, p1=None, comm=None, shard="rows"):
        self.d = len(kinds)
        self.comm = comm if shard == "rows" else None
        self.comm_batch = comm if shard == "batch" else None
        if comm is not None and shard == "rows":
            from bayesianinference_b200 import configs as _cfg
            from bayesianinference_b200.engine import shard_rows
            inputs = np.asarray(inputs, float)
            inputs = inputs.reshape(-1, 1) if inputs.ndim == 1 else inputs
            r0, r1 = shard_rows(inputs.shape[0], comm.rank, comm.world, overlap=1 if op == _cfg.OP_GBM else 0)
            inputs = inputs[r0:r1]
            outputs = None if outputs is None else np.asarray(outputs, float).reshape(-1, 1)[r0:r1]
        self.prob = O.Problem(op, self.d, inputs, outputs, iparam)
        self.prior = O.Prior(kinds, lo, hi, p0 or None, p1 or None)

    def loglike(self, theta):
        if self.comm_batch is not None:
            # batch-sharded (gp.cu: gp_loglike_device_strided): contiguous slices of cnt = ceil(P / world), one per rank
            theta = np.atleast_2d(np.asarray(theta, float))
            P, W, r = theta.shape[0], self.comm_batch.world, self.comm_batch.rank
            cnt = -(-P // W)
            lo, hi = min(P, r * cnt), min(P, r * cnt + cnt)
            mine = np.zeros(cnt)
            if hi > lo:
                mine[:hi - lo] = self.prob.loglike(theta[lo:hi], self.prior)
            return self.comm_batch.allgather(mine).reshape(-1)[:P]
        v = self.prob.loglike(theta, self.prior)
        if self.comm is None:
            return v
        parts = self.comm.allgather(v)  # [world][P], summed in rank order like shard_exchange (problem.cuh)
        bad = (parts <= O.LOGZERO).any(0)
        tot = np.zeros(parts.shape[1])
        for r in range(self.comm.world):
            tot = tot + parts[r]
        return np.where(bad, O.LOGZERO, tot)

    def logprior(self, theta):
        return self.prior.logpdf(theta)

    def predictive_components(self, theta, inputs):
        return self.prob.predictive_components(theta, inputs)

    def gp_predict(self, theta, xstar):
        return self.prob.gp_predict(theta, xstar)

    def sample_prior(self, n, seed=1, run_id=0):
        return self.prior.sample(n, seed, run_id)


class RunGroup:
    def __init__(self, problem, options, start_points=None):
        self.p, self.o = problem, options
        self.start = None if start_points is None else np.asarray(start_points).reshape(options.n_runs, options.pool_size, problem.d)
        self.results = None

    def advance(self, max_batches=0):
        o = self.o
        self.results = []
        for i in range(o.n_runs):
            r = O.nested_sampling(self.p.prob, self.p.prior, pool_size=o.pool_size, batch_k=o.batch_k,
                                  mc_steps=o.mc_steps, max_iter=o.max_iter, min_iter=o.min_iter,
                                  term_frac=o.term_frac, acc_range=(o.acc_min, o.acc_max), seed=o.seed, loglmax=o.loglmax,
                                  run_id=o.first_run_id + i, adapt_in_walk=False,
                                  start_points=None if self.start is None else self.start[i])
            self.results.append(r)
        return True

    def fetch(self, run=0):
        r = self.results[run]
        return dict(points=r.points, logL=r.logL, logPrior=r.logPrior, acc=r.acc, pool=r.pool, logX=r.logX,
                    crude_logw=r.crude_logw, crude_logZ=r.crude_logZ, entropy=r.entropy, logLmax=r.logLmax,
                    M=r.logL.size, n_deleted=r.n_deleted, iterations=r.iterations, evals=r.evals, n=r.n)

    def close(self):
        pass


class Chain:
    """engine.Chain served by the oracle restatement (one chain at a time, run from scratch on every call)."""

    def __init__(self, problem, start, init_cov, learn_delay=20, seed=1):
        self.p, self.start = problem, np.atleast_2d(np.asarray(start, float))
        self.cov0, self.delay, self.seed = np.asarray(init_cov, float), learn_delay, seed
        self.n_chains, self.d, self.steps, self.last = self.start.shape[0], problem.d, 0, None

    def iterate(self, n_steps, record=True):
        tot = self.steps + int(n_steps)
        self.last = [O.mcmc_chain(self.p.prob, self.p.prior, self.start[c], self.cov0, self.delay, self.seed, c, tot)
                     for c in range(self.n_chains)]
        out = np.stack([r["states"][self.steps:] for r in self.last], 1)
        self.steps = tot
        return out if record else None

    def state(self):
        r = self.last
        return dict(x=np.stack([q["states"][-1] for q in r]), mean=np.stack([q["mean"] for q in r]),
                    cov=np.stack([q["cov"] for q in r]), t=np.array([q["t"] for q in r]),
                    accepted=np.array([q["accepted"] for q in r]))

    def close(self):
        pass


def crude_weights(logL, pool, n_live):
    logL = np.asarray(logL, float)
    M = logL.size
    lx = O.xvalues_log(n_live, M - n_live, np.asarray(pool, np.int64))
    lw = O.trapezoid_log(lx) + logL
    z = O.logsumexp(lw)
    return dict(logX=lx, crude_logw=lw, crude_logZ=z, entropy=O.entropy(lw, logL, z), logLmax=float(logL.max()),
                log_missing=float(lx[-1] + logL.max()))


def evidence_sampling(points, logL, pool, n_live, post_runs=100, seed=1):
    r = O.evidence_sampling(points, logL, pool, n_live, post_runs, seed)
    return dict(z=r["zSamples"], H=np.full(post_runs, r["RelativeEntropy"]["Mean"]) if False else _H(r, post_runs),
                logw_mean=r["LogPosteriorWeight"]["Mean"], logw_sd=r["LogPosteriorWeight"]["StandardError"],
                slx_mean=r["SampledLogX"]["Mean"], slx_sd=r["SampledLogX"]["StandardError"],
                pmean=r["parameterSamples"])


def _H(r, post_runs):
    h = r.get("HSamples")
    if h is None:
        h = np.full(post_runs, r["RelativeEntropy"]["Mean"])
    return h
