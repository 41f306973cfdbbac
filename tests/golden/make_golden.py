"""Generates tests/golden/pins.json and tests/golden/loglike_vectors.npz.

The reference is Wolfram Language and cannot run in this image, so the golden numbers are
 (a) independent numerical evaluations of quantities the reference defines in closed form —
     the C1 / C4 evidences by 2-D quadrature (the operation directPosteriorDistribution performs with
     NIntegrate, BayesianStatistics.wl:114-126), and
 (b) per-parameter log-likelihood vectors of small synthetic problems computed with mpmath at 40 digits
     from the per-datum formulas of SURVEY.md §8a (independent of oracle/binest_oracle.c).
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import mpmath as mp
import numpy as np
from scipy import integrate

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from bayesianinference_b200 import configs as cfg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 40


def c1_logz():
    c = cfg.c1_gaussian()
    x = c.inputs[:, 0]
    N, xb, s2 = x.size, x.mean(), ((x - x.mean()) ** 2).sum()
    logLmax = -106.2516

    def f(sig, mu):
        ll = -N * np.log(sig) - 0.5 * N * np.log(2 * np.pi) - (s2 + N * (xb - mu) ** 2) / (2 * sig * sig)
        prior = 1 / 20 * 1 / (sig * np.log(10 / 0.01))
        return np.exp(ll - logLmax) * prior

    val, err = integrate.dblquad(f, 0.5, 3.0, lambda m: 0.3, lambda m: 2.0, epsabs=1e-14, epsrel=1e-12)
    return float(np.log(val) + logLmax)


def c4_logz():
    c = cfg.c4_gbm()
    t, x = c.inputs[:, 0], c.outputs[:, 0]
    dt, r = np.diff(t), np.diff(np.log(x))
    T = dt.size
    A, B, Cc = (r * r / dt).sum(), r.sum(), dt.sum()
    cst = -np.log(x[1:]).sum() - 0.5 * np.log(dt).sum() - T * 0.5 * np.log(2 * np.pi)
    ref = -72297.339738

    def f(sg, mu):
        m = mu - sg * sg / 2
        ll = cst - T * np.log(sg) - (A - 2 * m * B + m * m * Cc) / (2 * sg * sg)
        return np.exp(ll - ref) / 2.0 / (sg * np.log(2 / 0.01))

    val, _ = integrate.dblquad(f, -1, 1, lambda m: 0.23, lambda m: 0.27, epsabs=1e-14, epsrel=1e-12)
    return float(np.log(val) + ref)


def mp_loglike(c, th):
    th = [mp.mpf(float(v)) for v in th]
    X = c.inputs
    Y = None if c.outputs is None else c.outputs[:, 0]
    s = mp.mpf(0)
    if c.op == cfg.OP_GAUSSIAN_IID:
        mu, sg = th
        for v in X[:, 0]:
            s += -(mp.mpf(float(v)) - mu) ** 2 / (2 * sg ** 2) - mp.log(sg) - mp.log(2 * mp.pi) / 2
    elif c.op == cfg.OP_POLYREG:
        deg = c.iparam[0]
        sg = th[deg + 1]
        for xv, yv in zip(X[:, 0], Y):
            xv = mp.mpf(float(xv))
            m = sum(th[j] * xv ** j for j in range(deg + 1))
            s += -(mp.mpf(float(yv)) - m) ** 2 / (2 * sg ** 2) - mp.log(sg) - mp.log(2 * mp.pi) / 2
    elif c.op == cfg.OP_LOGISTIC:
        F, K = X.shape[1], c.iparam[1]
        for row, yv in zip(X, Y):
            z = [th[k * (F + 1) + F] + sum(th[k * (F + 1) + f] * mp.mpf(float(row[f])) for f in range(F)) for k in range(K - 1)]
            z.append(mp.mpf(0))
            s += z[int(yv)] - mp.log(sum(mp.e ** v for v in z))
    elif c.op == cfg.OP_GBM:
        mu, sg = th
        for i in range(1, X.shape[0]):
            dt = mp.mpf(float(X[i, 0])) - mp.mpf(float(X[i - 1, 0]))
            r = mp.log(mp.mpf(float(Y[i])) / mp.mpf(float(Y[i - 1])))
            s += -mp.log(mp.mpf(float(Y[i]))) - mp.log(2 * mp.pi * sg ** 2 * dt) / 2 \
                 - (r - (mu - sg ** 2 / 2) * dt) ** 2 / (2 * sg ** 2 * dt)
    return s


def vectors():
    out = {}
    rng = np.random.default_rng(2026)
    cases = {"C1": cfg.c1_gaussian(), "C2": cfg.c2_polyreg(N=300), "C3": cfg.c3_logistic(N=200), "C4": cfg.c4_gbm(T=150)}
    for name, c in cases.items():
        lo, hi = np.array(c.lo), np.array(c.hi)
        th = lo + (hi - lo) * rng.uniform(0.05, 0.95, (6, c.d))
        vals = [mp_loglike(c, row) for row in th]
        out[name + "_theta"] = th
        out[name + "_logL"] = np.array([float(v) for v in vals])
    return out


if __name__ == "__main__":
    pins = {"c1_logZ_quadrature": c1_logz(), "c4_logZ_quadrature": c4_logz()}
    json.dump(pins, open(os.path.join(HERE, "pins.json"), "w"), indent=1)
    np.savez(os.path.join(HERE, "loglike_vectors.npz"), **vectors())
    print(pins)
