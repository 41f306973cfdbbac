"""CPU emulation of the pivoted polynomial operator (csrc/operators.cuh OpPolyReg, abi_problem.cu pivots) in numpy
fp64, compared with the __float128 oracle — run before spending GPU time on the adversarial parity cases.
Emulates: pivot choice, pivoted rows, dd Taylor shift (via fractions, exact then rounded), Horner FMA chain is
approximated by plain fp64 ops (no FMA in numpy: slightly pessimistic), Sum t^2 by chunked fp64 partial sums."""
import math
import sys
from fractions import Fraction

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle as O  # noqa: E402
from bayesianinference_b200 import configs as cfg  # noqa: E402


def pivots(x, y, deg):
    n = x.size
    mean = x.sum() / n
    var = max((x * x).sum() / n - mean * mean, 0.0)
    xbar = float(mean) if abs(mean) > 0.5 * math.sqrt(var) else 0.0
    u = x - xbar
    sd = math.sqrt((u * u).sum() / n)
    V = np.vander(u / sd, deg + 1, increasing=True)
    a = np.linalg.lstsq(V, y, rcond=None)[0]
    return xbar, float(a[0])


def emulate(x, y, deg, th):
    xbar, piv = pivots(x, y, deg)
    xp, yp = x - xbar, y - piv
    m = [math.fsum(yp)] + [math.fsum(xp**k) for k in range(1, deg + 1)]
    out = []
    for t in th:
        c = [Fraction(float(v)) for v in t[:deg + 1]]
        sg = float(t[deg + 1])
        # exact Taylor shift, then round (the device does it in double-double)
        xb = Fraction(xbar)
        sh = []
        for k in range(deg + 1):
            sh.append(sum(c[j] * math.comb(j, k) * xb ** (j - k) for j in range(k, deg + 1)))
        ct = [float(v) for v in sh]
        delta = float(sh[0] - Fraction(piv))
        tt = np.full_like(xp, ct[deg]) if deg >= 1 else np.zeros_like(xp)
        for j in range(deg - 1, 0, -1):
            tt = tt * xp + ct[j]
        tt = tt * xp - yp
        acc = 0.0
        for ch in np.array_split(tt * tt, 148 * 8):  # per-warp partials, then a sequential combine
            acc += float(np.sum(ch))
        st = -m[0] + sum(ct[j] * m[j] for j in range(1, deg + 1))
        sse = acc + delta * (2.0 * st + x.size * delta)
        out.append(x.size * (-math.log(sg) - 0.9189385332046727) - sse / (2 * sg * sg))
    return np.array(out), xbar, piv


def case(name, x, y, deg, lo, hi, N, seed=1):
    names = [f"c{j}" for j in range(deg + 1)] + ["sigma"]
    c = cfg.Config(name, cfg.OP_POLYREG, x.reshape(-1, 1), y.reshape(-1, 1), (deg, 0, 0, 0), names,
                   [1] * (deg + 1) + [2], lo, hi)
    op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi)
    th = pr.sample(24, seed)
    # plus walkers near the least-squares fit (the posterior bulk, where the old form cancelled)
    V = np.vander(x, deg + 1, increasing=True)
    coef = np.linalg.lstsq(V, y, rcond=None)[0]
    s = math.sqrt(((y - V @ coef) ** 2).mean())
    near = np.tile(np.concatenate([coef, [s]]), (8, 1)) * (1 + 1e-6 * np.random.default_rng(3).standard_normal((8, deg + 2)))
    th = np.vstack([th, near])
    hi_, lo_ = op.loglike_quad(th)
    got, xbar, piv = emulate(x, y, deg, th)
    ok = hi_ > 0.5 * O.LOGZERO
    rel = np.abs((got[ok] - hi_[ok]) - lo_[ok]) / np.abs(hi_[ok])
    ref64 = op.loglike(th, pr)
    rel64 = np.abs((ref64[ok] - hi_[ok]) - lo_[ok]) / np.abs(hi_[ok])
    print(f"{name:28s} xbar={xbar:10.4f} piv={piv:12.5f}  max rel emu {rel.max():.2e} (near-fit {rel[-8:].max():.2e})  "
          f"seq fp64 oracle {rel64.max():.2e}")


if __name__ == "__main__":
    g = np.random.default_rng(0)
    N = 200_000
    x = g.uniform(-1, 1, N)
    base = 0.5 - 1.2 * x + 0.8 * x**2 + 0.3 * x**3
    case("C2-like", x, base + g.normal(0, 0.25, N), 3, [-5] * 4 + [0.01], [5] * 4 + [5], N)
    case("offset 50, sigma 0.05", x, base + 50 + g.normal(0, 0.05, N), 3, [-100] * 4 + [0.001], [100] * 4 + [5], N)
    case("offset 1000, sigma 0.01", x, base + 1000 + g.normal(0, 0.01, N), 3, [-2000] * 4 + [0.001], [2000] * 4 + [5], N)
    x2 = g.uniform(100, 101, N)
    case("x in (100,101) deg1", x2, 2 + 0.5 * (x2 - 100) + g.normal(0, 0.05, N), 1, [-100, -5, 0.001], [100, 5, 5], N)
    case("x in (100,101) deg3", x2, 2 + 0.5 * (x2 - 100.5) - 0.7 * (x2 - 100.5) ** 2 + g.normal(0, 0.05, N), 3,
         [-5] * 4 + [0.001], [5] * 4 + [5], N)
    case("y = 1000 x^2", x, 1000 * x * x + g.normal(0, 0.01, N), 2, [-2000] * 3 + [0.001], [2000] * 3 + [5], N)
