"""Generates tests/golden/gp_vectors.npz: GP marginal likelihoods (GP:181-199), GP predictive means / standard deviations
(GP:395-420) and predictive-component tables (BS:1437-1483) of small problems, computed with mpmath at 40 digits from
the textbook formulas — independent of oracle/binest_oracle.c and of the CUDA code.
Run from the repo root:  python tests/golden/make_golden_gp.py
"""
import os
import sys

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from bayesianinference_b200 import configs as cfg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 40


def gp_case(N=48, Q=7, n_theta=4, seed=31):
    c = cfg.c5_gp(N=N)
    rng = np.random.default_rng(seed)
    x = [mp.mpf(float(v)) for v in c.inputs[:, 0]]
    y = mp.matrix([mp.mpf(float(v)) for v in c.outputs[:, 0]])
    lo, hi = np.array(c.lo), np.array(c.hi)
    theta = np.exp(np.log(lo) + (np.log(hi) - np.log(lo)) * rng.uniform(0.3, 0.8, (n_theta, 3)))
    theta[0] = [1.0, 0.8, 0.3]
    xs = np.concatenate([rng.uniform(-0.5, 10.5, Q - 1), [c.inputs[5, 0]]])
    logL, mean, sd = [], [], []
    for sf, ell, sn in theta:
        sf, ell, sn = mp.mpf(float(sf)), mp.mpf(float(ell)), mp.mpf(float(sn))
        k = lambda a, b: sf**2 * mp.e ** (-(a - b) ** 2 / (2 * ell**2))
        K = mp.matrix(N, N)
        for i in range(N):
            for j in range(N):
                K[i, j] = k(x[i], x[j]) + (sn**2 if i == j else 0)
        L = mp.cholesky(K)
        z = mp.lu_solve(L, y)            # L z = y
        logdet = 2 * sum(mp.log(L[i, i]) for i in range(N))
        logL.append(float(-(N * mp.log(2 * mp.pi) + logdet + (z.T * z)[0]) / 2))
        alpha = mp.lu_solve(K, y)
        m_row, s_row = [], []
        for q in xs:
            q = mp.mpf(float(q))
            ks = mp.matrix([k(xi, q) for xi in x])
            m_row.append(float((ks.T * alpha)[0]))
            v = sf**2 + sn**2 - (ks.T * mp.lu_solve(K, ks))[0]
            s_row.append(float(mp.sqrt(v)))
        mean.append(m_row)
        sd.append(s_row)
    return {"gp_N": np.array(N), "gp_theta": theta, "gp_xstar": xs, "gp_logL": np.array(logL),
            "gp_mean": np.array(mean), "gp_sd": np.array(sd)}


def predictive_cases(seed=32):
    rng = np.random.default_rng(seed)
    out = {}
    th = np.column_stack([rng.uniform(-5, 5, (5, 4)), rng.uniform(0.05, 4.0, 5)])
    xs = rng.uniform(-1, 1, 6)
    out["poly_theta"], out["poly_x"] = th, xs
    out["poly_mean"] = np.array([[float(sum(mp.mpf(float(t[j])) * mp.mpf(float(x)) ** j for j in range(4))) for x in xs] for t in th])
    W = rng.normal(0, 2, (4, 10))
    X = rng.normal(size=(5, 4))
    probs = np.empty((4, 5, 3))
    for m in range(4):
        for q in range(5):
            z = [sum(mp.mpf(float(W[m, k * 5 + f])) * mp.mpf(float(X[q, f])) for f in range(4)) + mp.mpf(float(W[m, k * 5 + 4]))
                 for k in range(2)] + [mp.mpf(0)]
            den = sum(mp.e ** v for v in z)
            probs[m, q] = [float(mp.e ** v / den) for v in z]
    out["soft_theta"], out["soft_x"], out["soft_probs"] = W, X, probs
    return out


if __name__ == "__main__":
    d = gp_case()
    d.update(predictive_cases())
    np.savez(os.path.join(HERE, "gp_vectors.npz"), **d)
    print({k: v.shape for k, v in d.items()})
