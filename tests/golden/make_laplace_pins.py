"""Generates tests/golden/laplace_pins.json: full-size (N = 1e6) evidences of C2 and C3 by the Laplace formula the
reference itself encodes (laplacePosteriorFit, LaplaceApproximation.wl:22-30:
logZ ~= logL(th^) + log prior(th^) + (d/2) log 2pi - 1/2 log det H, H = -Hessian of the log posterior at the mode).

SURVEY.md §8c pin K5: at N = 1e6 the posterior is Gaussian to O(1/N); for C2 the formula is within 1e-5 of the exact
value -31513.459131 (coefficients integrated analytically), i.e. far inside the 3-sigma band (sigma ~ 0.2) the
full-size nested-sampling runs of tests/test_gpu_logz_full.py are graded against.

Independent of oracle/: plain numpy Newton iterations on the per-datum formulas of SURVEY.md §8a.
Run from the repo root:  python tests/golden/make_laplace_pins.py
"""
import json
import os
import sys

import numpy as np
from scipy.special import ndtr

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from bayesianinference_b200 import configs as cfg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LOG2PI = float(np.log(2 * np.pi))


def c2_laplace():
    c = cfg.c2_polyreg()
    x, y = c.inputs[:, 0], c.outputs[:, 0]
    N = x.size
    V = np.vander(x, 4, increasing=True)
    G, b = V.T @ V, V.T @ y

    def parts(th):
        cf, s = th[:4], th[4]
        r = y - V @ cf
        rss = float(r @ r)
        f = -rss / (2 * s * s) - N * np.log(s) - 0.5 * N * LOG2PI - np.log(s)  # + scale prior 1/s (unnormalised)
        g = np.concatenate([(b - G @ cf) / (s * s), [rss / s**3 - (N + 1) / s]])
        H = np.zeros((5, 5))
        H[:4, :4] = -G / (s * s)
        H[:4, 4] = H[4, :4] = -2 * (b - G @ cf) / s**3
        H[4, 4] = -3 * rss / s**4 + (N + 1) / (s * s)
        return f, g, H

    th = np.concatenate([np.linalg.solve(G, b), [0.25]])
    for _ in range(50):
        f, g, H = parts(th)
        step = np.linalg.solve(H, g)
        th = th - step
        if np.abs(step).max() < 1e-14:
            break
    f, g, H = parts(th)
    logprior_norm = -4 * np.log(10.0) - np.log(np.log(5.0 / 0.01))
    logz = f + logprior_norm + 2.5 * LOG2PI - 0.5 * np.linalg.slogdet(-H)[1]
    return {"logZ_laplace": float(logz), "mode": th.tolist(), "logL_mode": float(f + np.log(th[4]))}


def c3_laplace():
    c = cfg.c3_logistic()
    X, lab = c.inputs, c.outputs[:, 0].astype(np.int64)
    N, F = X.shape
    K = 3
    d = (K - 1) * (F + 1)
    A = np.concatenate([X, np.ones((N, 1))], axis=1)          # parameter order per class: w_0..w_{F-1}, b
    Y = np.zeros((N, K - 1))
    for k in range(K - 1):
        Y[:, k] = lab == k
    sd = 5.0

    def parts(th):
        W = th.reshape(K - 1, F + 1)
        z = A @ W.T                                            # N x (K-1); reference class K has z = 0
        m = np.maximum(z.max(1), 0.0)
        e = np.exp(z - m[:, None])
        den = e.sum(1) + np.exp(-m)
        p = e / den[:, None]
        ll = float((z * Y).sum() - (m + np.log(den)).sum())
        f = ll - 0.5 * float(th @ th) / sd**2
        g = ((Y - p).T @ A).reshape(-1) - th / sd**2
        H = np.zeros((d, d))
        for a in range(K - 1):
            for b_ in range(K - 1):
                w = p[:, a] * ((a == b_) - p[:, b_])
                H[a * (F + 1):(a + 1) * (F + 1), b_ * (F + 1):(b_ + 1) * (F + 1)] = -(A * w[:, None]).T @ A
        H -= np.eye(d) / sd**2
        return f, g, H, ll

    th = np.zeros(d)
    for _ in range(60):
        f, g, H, ll = parts(th)
        step = np.linalg.solve(H, g)
        th = th - step
        if np.abs(step).max() < 1e-13:
            break
    f, g, H, ll = parts(th)
    # N(0, 5^2) truncated to (-10, 10) per dimension (BS:51-59): density phi(t/5)/5 / (Phi(2) - Phi(-2))
    lognorm = d * (-np.log(sd) - 0.5 * LOG2PI - np.log(ndtr(2.0) - ndtr(-2.0)))
    logz = f + lognorm + 0.5 * d * LOG2PI - 0.5 * np.linalg.slogdet(-H)[1]
    # information H = E_post[logL] - logZ ~= logL_mode - d/2 - logZ for a Gaussian posterior
    return {"logZ_laplace": float(logz), "mode": th.tolist(), "logL_mode": float(ll),
            "information_nats": float(ll - 0.5 * d - logz)}


def c5_small_quadrature(N=256, npts=(28, 40)):
    """C5-shaped GP problem at N = 256 (SURVEY §8c pin K7): logZ by 3-D Gauss-Legendre quadrature in u = log(theta)
    (the scale prior 1/(theta log(hi/lo)) is uniform in u), on a +-8 sigma box of the Laplace-whitened coordinates
    around the mode.  Two grid sizes must agree (recorded)."""
    from scipy.linalg import cho_factor, cho_solve
    from scipy.optimize import minimize
    c = cfg.c5_gp(N=N)
    x, y = c.inputs[:, 0], c.outputs[:, 0]
    D2 = (x[:, None] - x[None, :]) ** 2
    lo, hi = np.array(c.lo), np.array(c.hi)

    def ll(u):
        sf, ell, sn = np.exp(u)
        Kmat = sf * sf * np.exp(-D2 / (2 * ell * ell))
        Kmat[np.diag_indices(N)] += sn * sn
        try:
            cf = cho_factor(Kmat, lower=True)
        except np.linalg.LinAlgError:
            return -1e300
        a = cho_solve(cf, y)
        return -0.5 * (N * LOG2PI + 2 * np.log(np.diag(cf[0])).sum() + y @ a)

    r = minimize(lambda u: -ll(u), np.log([1.0, 0.8, 0.1]), method="Nelder-Mead",
                 options={"xatol": 1e-7, "fatol": 1e-9, "maxiter": 4000})
    u0, f0 = r.x, -r.fun
    h = 1e-3
    H = np.zeros((3, 3))
    E = np.eye(3) * h
    for i in range(3):
        for j in range(3):
            H[i, j] = -(ll(u0 + E[i] + E[j]) - ll(u0 + E[i] - E[j]) - ll(u0 - E[i] + E[j]) + ll(u0 - E[i] - E[j])) / (4 * h * h)
    L = np.linalg.cholesky(np.linalg.inv(0.5 * (H + H.T)))   # u = u0 + L t, t ~ N(0, I) under the Laplace fit
    lognorm = -np.log(np.log(hi / lo)).sum()
    vals = {}
    for n in npts:
        t, w = np.polynomial.legendre.leggauss(n)
        t, w = 8.0 * t, 8.0 * w
        acc = 0.0
        for a_, wa in zip(t, w):
            for b_, wb in zip(t, w):
                for c_, wc in zip(t, w):
                    u = u0 + L @ np.array([a_, b_, c_])
                    if np.any(u <= np.log(lo)) or np.any(u >= np.log(hi)):
                        continue
                    acc += wa * wb * wc * np.exp(ll(u) - f0)
        vals[n] = float(np.log(acc) + f0 + np.log(abs(np.linalg.det(L))) + lognorm)
    laplace = float(f0 + 1.5 * LOG2PI + np.log(abs(np.linalg.det(L))) + lognorm)
    ns = sorted(vals)
    return {"N": N, "logZ_quadrature": vals[ns[-1]], "logZ_quadrature_coarse": vals[ns[0]], "logZ_laplace_logparams": laplace,
            "mode": np.exp(u0).tolist(), "logL_mode": float(f0)}


if __name__ == "__main__":
    out = {"C2": c2_laplace(), "C3": c3_laplace(), "C5_small": c5_small_quadrature(),
           "note": "Laplace evidences (LA:22-30) of the full-size C2/C3 configs; C2 exact value is -31513.459131"}
    with open(os.path.join(HERE, "laplace_pins.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))
