"""Synthetic inputs of the five BASELINE.json configs.

Host-side numpy only (``Generator(Philox(key=seed))``, fp64, draws in the stated order); shared by
tests/ and bench.py.  The reference's worked examples lived in a notebook that is stripped from
the checkout (/root/reference/.MISSING_LARGE_BLOBS), so these are the shapes the configs name,
not copies of reference data.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# operator / prior ids (include/binest.h)
OP_GAUSSIAN_IID, OP_POLYREG, OP_LOGISTIC, OP_GBM, OP_GP_SE = 1, 2, 3, 4, 5
PRIOR_UNIFORM, PRIOR_SCALE, PRIOR_NORMAL_TRUNC = 1, 2, 3


@dataclass
class Config:
    name: str
    op: int
    inputs: np.ndarray           # N x n_in
    outputs: np.ndarray | None   # N x n_out
    iparam: tuple                # (degree | -, n_classes | -, -, -)
    names: list
    kinds: list
    lo: list
    hi: list
    p0: list = field(default_factory=list)
    p1: list = field(default_factory=list)
    pool_size: int = 100
    parallel_runs: int = 1
    truth: dict = field(default_factory=dict)

    @property
    def d(self):
        return len(self.names)


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=seed))


def c1_gaussian(N=100, seed=101):
    """C1: 1-D Gaussian mean/sigma, n=100 live points, N=100 data."""
    x = _rng(seed).normal(1.5, 0.7, N)
    return Config("C1-gaussian", OP_GAUSSIAN_IID, x.reshape(-1, 1), None, (0, 0, 0, 0), ["mu", "sigma"],
                  [PRIOR_UNIFORM, PRIOR_SCALE], [-10.0, 0.01], [10.0, 10.0], pool_size=100,
                  truth={"logZ": -114.641064} if (N, seed) == (100, 101) else {})


def c2_polyreg(N=1_000_000, seed=102, degree=3):
    """C2: polynomial regression, 5 params, N=1e6, n=1024."""
    g = _rng(seed)
    x = g.uniform(-1.0, 1.0, N)
    y = 0.5 - 1.2 * x + 0.8 * x**2 + 0.3 * x**3 + g.normal(0.0, 0.25, N)
    names = [f"c{j}" for j in range(degree + 1)] + ["sigma"]
    return Config("C2-polyreg", OP_POLYREG, x.reshape(-1, 1), y.reshape(-1, 1), (degree, 0, 0, 0), names,
                  [PRIOR_UNIFORM] * (degree + 1) + [PRIOR_SCALE], [-5.0] * (degree + 1) + [0.01],
                  [5.0] * (degree + 1) + [5.0], pool_size=1024,
                  truth={"logZ": -31513.459131, "logLmax": -31470.229840} if (N, seed, degree) == (1_000_000, 102, 3) else {})


def c3_logistic(N=1_000_000, seed=103, F=4, K=3):
    """C3: softmax classification, reference class K (z_K = 0), (K-1)(F+1) parameters, n=2048."""
    g = _rng(seed)
    x = g.standard_normal((N, F))
    W = g.standard_normal((K, F))
    b = g.standard_normal(K)
    W[K - 1] = 0.0
    b[K - 1] = 0.0
    z = x @ W.T + b
    z -= z.max(1, keepdims=True)
    p = np.exp(z)
    p /= p.sum(1, keepdims=True)
    u = g.uniform(size=N)
    y = (u[:, None] > np.cumsum(p, 1)).sum(1).clip(0, K - 1)
    d = (K - 1) * (F + 1)
    names = [f"w{k}{f}" for k in range(K - 1) for f in list(range(F)) + ["b"]]
    theta_true = np.concatenate([np.concatenate([W[k], [b[k]]]) for k in range(K - 1)])
    return Config("C3-logistic", OP_LOGISTIC, x, y.astype(np.float64).reshape(-1, 1), (0, K, 0, 0), names,
                  [PRIOR_NORMAL_TRUNC] * d, [-10.0] * d, [10.0] * d, [0.0] * d, [5.0] * d, pool_size=2048,
                  truth={"theta": theta_true})


def c4_gbm(T=16384, seed=104, mu=0.08, sigma=0.25, x0=100.0, dt=1.0 / 252.0):
    """C4: GeometricBrownianMotionProcess path, 64 runs x 512 live points."""
    z = _rng(seed).standard_normal(T)
    r = (mu - sigma**2 / 2) * dt + sigma * np.sqrt(dt) * z
    xs = x0 * np.exp(np.concatenate([[0.0], np.cumsum(r)]))
    ts = dt * np.arange(T + 1)
    return Config("C4-gbm", OP_GBM, ts.reshape(-1, 1), xs.reshape(-1, 1), (0, 0, 0, 0), ["mu", "sigma"],
                  [PRIOR_UNIFORM, PRIOR_SCALE], [-1.0, 0.01], [1.0, 2.0], pool_size=512, parallel_runs=64,
                  truth={"logZ": -72306.535014, "logL(0.08,0.25)": -72297.49504288615} if (T, seed) == (16384, 104) else {})


def c5_gp(N=4096, seed=105):
    """C5: GP regression, SE kernel + nugget, theta = (sigma_f, ell, sigma_n), 256 theta-sets."""
    g = _rng(seed)
    x = np.sort(g.uniform(0.0, 10.0, N))
    y = np.sin(x) + 0.5 * np.cos(2.3 * x) + g.normal(0.0, 0.1, N)
    return Config("C5-gp", OP_GP_SE, x.reshape(-1, 1), y.reshape(-1, 1), (0, 0, 1, 0), ["sigma_f", "ell", "sigma_n"],
                  [PRIOR_SCALE] * 3, [0.05, 0.05, 0.02], [5.0, 5.0, 1.0], pool_size=256)


ALL = {"C1": c1_gaussian, "C2": c2_polyreg, "C3": c3_logistic, "C4": c4_gbm, "C5": c5_gp}
