"""CPU tests of the oracle against constructed known-answer pins.

The reference ships no tests, fixtures or golden vectors (and is Wolfram Language, which cannot run here),
so parity is UNPINNED by the reference; every pin below is tied to a formula the reference encodes or to
a published known-answer vector.  Golden numbers under tests/golden/ were produced by
tests/golden/make_golden.py (committed)."""
import json
import os

import mpmath as mp
import numpy as np
import pytest

from bayesianinference_b200 import configs as cfg
from oracle import oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pins.json")))


def test_k8_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert O.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert O.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_and_normal_streams():
    u = np.array([O.uniform2(7, i, 0, 0, 4)[0] for i in range(4000)])
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.02
    z = np.array([v for i in range(4000) for v in O.normal2(7, i, 1, 2, 3)])
    assert abs(z.mean()) < 0.04 and abs(z.std() - 1) < 0.03


def test_k2_logspace_helpers_vs_mpmath():
    mp.mp.dps = 50
    rng = np.random.default_rng(1)
    for _ in range(200):
        a, b = rng.uniform(-800, 5, 2)
        hi, lo = max(a, b), min(a, b)
        assert abs(O.logadd(a, b) - float(mp.log(mp.e ** mp.mpf(a) + mp.e ** mp.mpf(b)))) <= 4e-16 * max(1, abs(hi))
        if hi - lo > 1e-3:
            ref = float(mp.log(mp.e ** mp.mpf(hi) - mp.e ** mp.mpf(lo)))
            assert abs(O.logsubtract(hi, lo) - ref) <= 1e-12 * max(1, abs(ref))
    v = rng.normal(-31500, 30, 5000)
    ref = float(mp.log(mp.fsum(mp.e ** mp.mpf(float(x)) for x in v)))
    assert abs(O.logsumexp(v) - ref) < 1e-11 * abs(ref)
    assert O.logsumexp(np.array([-np.inf, -3.0, -np.inf])) == -3.0  # Select[NumericQ] BU:333


def test_k3_x_sequence_pins():
    lx = O.xvalues_log(100, 900)
    x = np.exp(lx)
    np.testing.assert_allclose(x[[0, 899, 900, 999]],
                               [np.exp(-0.01), np.exp(-9.0), 100 / 101 * np.exp(-9.0), 1 / 101 * np.exp(-9.0)], rtol=1e-14)
    pool = np.concatenate([np.full(900, 100), np.arange(100, 0, -1)])
    np.testing.assert_allclose(O.xvalues_log(100, 900, pool), lx, rtol=1e-13)


@pytest.mark.parametrize("n,nd", [(100, 900), (2, 0), (5, 1), (1024, 40000)])
def test_k1_trapezoid_weights_sum_to_one(n, nd):
    w = np.exp(O.trapezoid_log(O.xvalues_log(n, nd)))
    assert abs(w.sum() - 1.0) < 1e-12
    # linear form BS:747-755 == log form BS:756-771
    x = np.exp(O.xvalues_log(n, nd))
    lin = 0.5 * (np.concatenate([[2 - x[0]], x[:-1]]) - np.concatenate([x[1:], [-x[-1]]]))
    np.testing.assert_allclose(w, lin, rtol=1e-9, atol=1e-300)


def test_data_pins_survey_appendix_a():
    c1 = cfg.c1_gaussian()
    x = c1.inputs[:, 0]
    assert abs(x.mean() - 1.555826847319) < 1e-11 and abs(x.std() - 0.700176812403) < 1e-11
    c2 = cfg.c2_polyreg()
    X = np.vander(c2.inputs[:, 0], 4, increasing=True)
    coef, *_ = np.linalg.lstsq(X, c2.outputs[:, 0], rcond=None)
    np.testing.assert_allclose(coef, [0.50012547, -1.19948747, 0.79976443, 0.29847954], atol=5e-9)
    rss = ((c2.outputs[:, 0] - X @ coef) ** 2).sum()
    assert abs(rss - 62353.42935214733) < 1e-6
    s = np.sqrt(rss / x.size * x.size / c2.inputs.shape[0])
    p2 = O.Problem(c2.op, c2.d, c2.inputs, c2.outputs, c2.iparam)
    th = np.concatenate([coef, [np.sqrt(rss / c2.inputs.shape[0])]])
    assert abs(p2.loglike(th)[0] - (-31470.229840)) < 1e-6  # logL_max pin
    hi, lo = p2.loglike_quad(th)
    assert abs(p2.loglike(th)[0] - hi[0] - lo[0]) < 1e-10 * abs(hi[0])


def test_k6_gbm_sufficient_sums():
    """C4: the per-datum sum equals the closed form through the four sufficient sums."""
    c = cfg.c4_gbm()
    p = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    t, x = c.inputs[:, 0], c.outputs[:, 0]
    dt, r = np.diff(t), np.diff(np.log(x))
    T = dt.size
    for mu, sg in [(0.08, 0.25), (-0.3, 0.9), (0.5, 0.05)]:
        m = mu - sg * sg / 2
        closed = (-np.log(x[1:]).sum() - 0.5 * np.log(dt).sum() - T * (0.5 * np.log(2 * np.pi) + np.log(sg))
                  - ((r * r / dt).sum() - 2 * m * r.sum() + m * m * dt.sum()) / (2 * sg * sg))
        assert abs(p.loglike([mu, sg])[0] - closed) < 2e-11 * abs(closed)
    assert abs(p.loglike([0.08, 0.25])[0] - (-72297.49504288615)) < 1e-7


def test_logistic_matches_numpy_and_constraints():
    c = cfg.c3_logistic(N=5000)
    p = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi, c.p0, c.p1)
    th = pr.sample(6, 3)
    y = c.outputs[:, 0].astype(int)
    for row, got in zip(th, p.loglike(th, pr)):
        W = row.reshape(2, 5)
        z = np.concatenate([c.inputs @ W[:, :4].T + W[:, 4], np.zeros((len(y), 1))], 1)
        ref = (z[np.arange(len(y)), y] - np.log(np.exp(z - z.max(1, keepdims=True)).sum(1)) - z.max(1)).sum()
        assert abs(got - ref) < 1e-10 * abs(ref)
    bad = th[0].copy()
    bad[0] = 11.0
    assert p.loglike(bad, pr)[0] == O.LOGZERO  # outside the box -> logzero (BS:580-583)
    c1 = cfg.c1_gaussian()
    p1 = O.Problem(c1.op, c1.d, c1.inputs, None, c1.iparam)
    assert p1.loglike([0.0, -1.0])[0] == O.LOGZERO  # sigma <= 0 (BS:439)


def test_k7_gp_lu_vs_cholesky_vs_scipy():
    import scipy.linalg as sl
    c = cfg.c5_gp(N=200)
    p = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    x, y = c.inputs[:, 0], c.outputs[:, 0]
    for th in ([1.0, 0.8, 0.3], [0.5, 2.0, 0.1], [2.5, 0.3, 0.9]):
        K = th[0] ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * th[1] ** 2)) + th[2] ** 2 * np.eye(x.size)
        cf = sl.cho_factor(K, lower=True)
        ref = -0.5 * (x.size * np.log(2 * np.pi) + 2 * np.log(np.diag(cf[0])).sum() + y @ sl.cho_solve(cf, y))
        got = p.loglike(th)[0]  # LU path, GP:130-141
        hi, lo = p.loglike_quad(th)  # long-double Cholesky
        assert abs(got - ref) < 1e-9 * abs(ref) and abs(hi[0] - ref) < 1e-9 * abs(ref)


def test_prior_logpdf_and_sampling():
    pr = O.Prior([O.PRIOR_UNIFORM, O.PRIOR_SCALE, O.PRIOR_NORMAL_TRUNC], [-10, 0.01, -10], [10, 10, 10], [0, 0, 0], [1, 1, 5])
    from scipy import stats
    th = np.array([[1.0, 0.5, 2.0]])
    ref = -np.log(20) + (-np.log(0.5) - np.log(np.log(1000))) + stats.truncnorm.logpdf(2.0, -2, 2, 0, 5)
    assert abs(pr.logpdf(th)[0] - ref) < 1e-13
    assert pr.logpdf([[1.0, 0.5, 10.0]])[0] == O.LOGZERO  # open box (BS:327-336)
    s = pr.sample(4000, 5)
    assert (s > [-10, 0.01, -10]).all() and (s < [10, 10, 10]).all()
    assert abs(np.log(s[:, 1]).mean() - 0.5 * (np.log(0.01) + np.log(10))) < 0.15  # log-uniform
    assert abs(s[:, 2].std() - stats.truncnorm.std(-2, 2, 0, 5)) < 0.2


def test_k4_c1_evidence_pin_and_ns_statistics():
    """K4: quadrature value of the C1 evidence (what directPosteriorDistribution computes, BS:114-126) and
    K10-lite: pulls of the sequential oracle (reference scheme, K = 1, in-walk adaptation) over seeds."""
    assert abs(GOLD["c1_logZ_quadrature"] - (-114.641064)) < 2e-6
    c = cfg.c1_gaussian()
    p = O.Problem(c.op, c.d, c.inputs, None, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi)
    pulls, sds = [], []
    for seed in range(1, 9):
        r = O.nested_sampling(p, pr, pool_size=100, batch_k=1, seed=seed, adapt_in_walk=True)
        ev = O.evidence_sampling(r.points, r.logL, r.pool, 100, 100, seed, sorted_draws=True)
        pulls.append((ev["LogEvidence"]["Mean"] - GOLD["c1_logZ_quadrature"]) / ev["LogEvidence"]["StandardError"])
        sds.append(ev["LogEvidence"]["StandardError"])
        assert 700 < r.iterations < 1000
    assert abs(np.mean(pulls)) < 1.0 and 0.4 < np.std(pulls) < 1.8, pulls
    assert 0.2 < np.mean(sds) < 0.35  # sqrt(H/n) = 0.27 (Skilling 2006)


def test_batched_replacement_is_consistent():
    """K = 8 worst replaced per iteration with pool sizes n - j: statistically the same evidence."""
    c = cfg.c1_gaussian()
    p = O.Problem(c.op, c.d, c.inputs, None, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi)
    pulls = []
    for seed in range(1, 7):
        r = O.nested_sampling(p, pr, pool_size=100, batch_k=8, seed=seed, adapt_in_walk=False)
        assert set(np.unique(r.pool[: r.n_deleted])) <= set(range(93, 101))
        ev = O.evidence_sampling(r.points, r.logL, r.pool, 100, 100, seed)
        pulls.append((ev["LogEvidence"]["Mean"] - GOLD["c1_logZ_quadrature"]) / ev["LogEvidence"]["StandardError"])
    assert max(abs(x) for x in pulls) < 3.5 and abs(np.mean(pulls)) < 1.3, pulls


def test_evidence_sampling_sorted_vs_renyi_same_law():
    rng = np.random.default_rng(2)
    n, nd = 50, 400
    logL = np.sort(rng.normal(-50, 5, n + nd))
    pts = rng.normal(size=(n + nd, 2))
    pool = np.concatenate([np.full(nd, n), np.arange(n, 0, -1)])
    a = O.evidence_sampling(pts, logL, pool, n, 400, 3, sorted_draws=True)
    b = O.evidence_sampling(pts, logL, pool, n, 400, 4, sorted_draws=False)
    assert abs(a["LogEvidence"]["Mean"] - b["LogEvidence"]["Mean"]) < 4 * a["LogEvidence"]["StandardError"] / np.sqrt(200)
    assert abs(a["LogEvidence"]["StandardError"] / b["LogEvidence"]["StandardError"] - 1) < 0.25


def test_k9_combine_runs_invariants():
    c = cfg.c1_gaussian()
    p = O.Problem(c.op, c.d, c.inputs, None, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi)
    runs = [O.nested_sampling(p, pr, pool_size=40, batch_k=1, mc_steps=30, max_iter=150, seed=3, run_id=i) for i in range(3)]
    m = O.combine_runs(runs)
    assert m["n"] == 120  # BS:1307
    total = sum(r.logL.size for r in runs)
    dup = total - np.unique(np.concatenate([r.points for r in runs]), axis=0).shape[0]
    assert m["logL"].size == total - dup and m["n_deleted"] == m["logL"].size - 120  # BS:1294-1309
    assert np.all(np.diff(m["logL"]) >= 0)
    assert m["pool"][0] == 120 and m["pool"].max() == 120


def test_gp_predict_oracle_matches_closed_form():
    """predictFromGaussianProcessInternal (GP:395-420) restated with LU (fp64) and long-double Cholesky against the
    textbook formulas evaluated with scipy."""
    from scipy.linalg import cho_factor, cho_solve
    from bayesianinference_b200 import configs as cfg
    c = cfg.c5_gp(N=150)
    op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    th = np.array([[1.0, 0.8, 0.3], [0.7, 1.5, 0.05], [2.0, 0.3, 0.1]])
    xs = np.linspace(-0.5, 10.5, 17)
    m, s = op.gp_predict(th, xs)
    ml, sl = op.gp_predict(th, xs, long_double=True)
    x, y = c.inputs[:, 0], c.outputs[:, 0]
    for i, (sf, ell, sn) in enumerate(th):
        K = sf**2 * np.exp(-(x[:, None] - x[None]) ** 2 / (2 * ell * ell)) + sn**2 * np.eye(x.size)
        ks = sf**2 * np.exp(-(x[:, None] - xs[None]) ** 2 / (2 * ell * ell))
        cf = cho_factor(K)
        mm = ks.T @ cho_solve(cf, y)
        vv = sf**2 + sn**2 - np.sum(ks * cho_solve(cf, ks), 0)
        np.testing.assert_allclose(m[i], mm, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(ml[i], mm, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(s[i], np.sqrt(vv), rtol=1e-9)
        np.testing.assert_allclose(sl[i], np.sqrt(vv), rtol=1e-9)
    bad = op.gp_predict(np.array([[1.0, 0.8, 0.3], [0.0, 1.0, 0.0]]), xs[:3])  # K = 0: singular -> the reference Throws
    assert np.isnan(bad[0][1]).all()


def test_predictive_components_oracle_matches_numpy():
    """BS:1437-1483 restated: component parameters of the predictive mixture (polynomial regression, softmax)."""
    from bayesianinference_b200 import configs as cfg
    rng = np.random.default_rng(3)
    c = cfg.c2_polyreg(N=50)
    op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    th = O.Prior(c.kinds, c.lo, c.hi).sample(7, 2)
    xs = rng.uniform(-1, 1, 5)
    out = op.predictive_components(th, xs)
    np.testing.assert_allclose(out[:, :, 0], np.polynomial.polynomial.polyval(xs, th[:, :4].T), rtol=1e-14)
    np.testing.assert_array_equal(out[:, :, 1], np.repeat(th[:, 4:5], 5, 1))
    c = cfg.c3_logistic(N=50)
    op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    th = rng.normal(size=(6, 10))
    X = rng.normal(size=(4, 4))
    out = op.predictive_components(th, X)
    W = th.reshape(6, 2, 5)
    z = np.concatenate([np.einsum("mkf,qf->mqk", W[:, :, :4], X) + W[:, None, :, 4], np.zeros((6, 4, 1))], -1)
    pr = np.exp(z - z.max(-1, keepdims=True))
    pr /= pr.sum(-1, keepdims=True)
    np.testing.assert_allclose(out, pr, rtol=1e-13)


def test_mcmc_chain_oracle_samples_the_known_posterior():
    """createMCMCChain / iterateMCMC restated (BS:630-703): on C1 (flat prior on mu, 1/sigma prior on sigma) the
    marginal posterior of mu is Student t_{N-1}(xbar, s^2/N): mean xbar, variance s^2/N (N-1)/(N-3)."""
    from bayesianinference_b200 import configs as cfg
    c = cfg.c1_gaussian()
    op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi)
    x = c.inputs[:, 0]
    N = x.size
    r = O.mcmc_chain(op, pr, [1.0, 1.0], np.eye(2) * 0.01, delay=20, seed=5, chain_id=0, n_steps=40000)
    st = r["states"][4000:]
    assert r["t"] == 40001 and 0.15 < r["accepted"] / 40000 < 0.6
    sd_mu = np.sqrt(x.var(ddof=1) / N * (N - 1) / (N - 3))
    assert abs(st[:, 0].mean() - x.mean()) < 0.1 * sd_mu      # ~ 3 sigma of a chain with ESS ~ 1000
    assert abs(st[:, 0].std() / sd_mu - 1.0) < 0.1
    # E[sigma^2 | data] = S / (N - 3) for the 1/sigma prior (scaled inverse chi^2 with N - 1 dof)
    assert abs((st[:, 1] ** 2).mean() / (((x - x.mean()) ** 2).sum() / (N - 3)) - 1.0) < 0.05
    # chain estimates = running moments of ALL visited states (start included)
    allst = np.vstack([[1.0, 1.0], r["states"]])
    np.testing.assert_allclose(r["mean"], allst.mean(0), rtol=1e-10)
    np.testing.assert_allclose(r["cov"], np.cov(allst.T), rtol=1e-8)
    # same seed -> same chain; other chain id -> another stream
    r2 = O.mcmc_chain(op, pr, [1.0, 1.0], np.eye(2) * 0.01, delay=20, seed=5, chain_id=0, n_steps=50)
    np.testing.assert_array_equal(r2["states"], r["states"][:50])
    r3 = O.mcmc_chain(op, pr, [1.0, 1.0], np.eye(2) * 0.01, delay=20, seed=5, chain_id=1, n_steps=50)
    assert not np.array_equal(r3["states"], r2["states"])
    with pytest.raises(ValueError):
        O.mcmc_chain(op, pr, [1.0, -1.0], np.eye(2), n_steps=5)  # outside the box
