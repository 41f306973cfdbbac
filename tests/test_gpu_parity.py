"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances: per-parameter log-likelihoods 1e-12 relative against the __float128 oracle (BASELINE.json
north_star); integer/index outputs exact; RNG-driven trajectories 1e-9 (libm vs libdevice differ in the
last ulp of log/sincos)."""
import numpy as np
import pytest

from bayesianinference_b200 import configs as cfg

pytestmark = pytest.mark.gpu

RTOL_LOGL = 1e-12


@pytest.fixture(scope="module")
def eng():
    from bayesianinference_b200 import engine
    engine.init()
    return engine


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _pair(eng, O, c):
    gp = eng.Problem.from_config(c)
    op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi, c.p0 or None, c.p1 or None)
    return gp, op, pr


def _thetas(O, pr, c, P, seed=900):
    th = pr.sample(P, seed)
    # a few rows outside the box / violating the operator constraints -> logzero on both sides
    if P >= 8:
        th[1, -1] = -0.3
        th[3, 0] = c.hi[0] + 1.0
        th[5, 0] = c.lo[0]
    return th


CASES = {
    "C1": lambda: cfg.c1_gaussian(),
    "C1-odd": lambda: cfg.c1_gaussian(N=777, seed=7),
    "C2-20k": lambda: cfg.c2_polyreg(N=20_000),
    "C2-deg1": lambda: cfg.c2_polyreg(N=5_001, degree=1),
    "C2-deg5": lambda: cfg.c2_polyreg(N=4_097, degree=5),
    "C3-20k": lambda: cfg.c3_logistic(N=20_000),
    "C3-binary": lambda: cfg.c3_logistic(N=3_000, F=2, K=2),
    "C4": lambda: cfg.c4_gbm(),
    "C4-short": lambda: cfg.c4_gbm(T=37),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("P", [1, 33, 300])
def test_loglike_matches_oracle(eng, O, name, P):
    c = CASES[name]()
    gp, op, pr = _pair(eng, O, c)
    th = _thetas(O, pr, c, P)
    got = gp.loglike(th)
    ref64 = op.loglike(th, pr)
    hi, lo = op.loglike_quad(th)
    ok = ref64 > 0.5 * O.LOGZERO
    assert np.array_equal(got[~ok], ref64[~ok]), "constraint violations must give logzero exactly"
    rel = np.abs((got[ok] - hi[ok]) - lo[ok]) / np.abs(hi[ok])
    assert rel.max() <= RTOL_LOGL, f"{name}: max rel err vs float128 oracle {rel.max():.3e}"
    # the fp64 sequential restatement itself sits within the same band of the quad value
    rel64 = np.abs((ref64[ok] - hi[ok]) - lo[ok]) / np.abs(hi[ok])
    assert rel64.max() <= 1e-10


def _poly_cfg(x, y, deg, lo, hi):
    names = [f"c{j}" for j in range(deg + 1)] + ["sigma"]
    return cfg.Config("poly-adv", cfg.OP_POLYREG, x.reshape(-1, 1), y.reshape(-1, 1), (deg, 0, 0, 0), names,
                      [cfg.PRIOR_UNIFORM] * (deg + 1) + [cfg.PRIOR_SCALE], lo, hi)


def _poly_adversarial(name, N=200_000):
    g = np.random.default_rng(0)
    x = g.uniform(-1, 1, N)
    base = 0.5 - 1.2 * x + 0.8 * x**2 + 0.3 * x**3
    if name == "offset50":      # |c0| = 50 >> sigma = 0.05
        return _poly_cfg(x, base + 50 + g.normal(0, 0.05, N), 3, [-100.0] * 4 + [0.001], [100.0] * 4 + [5.0])
    if name == "offset1000":    # |c0| = 1000, sigma = 0.01: the old 4-slot form lost 1e-6 here
        return _poly_cfg(x, base + 1000 + g.normal(0, 0.01, N), 3, [-2000.0] * 4 + [0.001], [2000.0] * 4 + [5.0])
    if name == "quadratic1000":  # mean(y) = 333 but the intercept is 0: a mean-of-y pivot would fail
        return _poly_cfg(x, 1000 * x * x + g.normal(0, 0.01, N), 2, [-2000.0] * 3 + [0.001], [2000.0] * 3 + [5.0])
    x2 = g.uniform(100, 101, N)
    if name == "x100-deg1":     # uncentred inputs: the intercept is compensated by the slope along a ridge
        return _poly_cfg(x2, 2 + 0.5 * (x2 - 100) + g.normal(0, 0.05, N), 1, [-100.0, -5.0, 0.001], [100.0, 5.0, 5.0])
    if name == "x100-deg3":
        return _poly_cfg(x2, 2 + 0.5 * (x2 - 100.5) - 0.7 * (x2 - 100.5) ** 2 + g.normal(0, 0.05, N), 3,
                         [-1e6, -1e5, -1e3, -5.0, 0.001], [1e6, 1e5, 1e3, 5.0, 5.0])
    raise KeyError(name)


@pytest.mark.parametrize("name", ["offset50", "offset1000", "quadratic1000", "x100-deg1", "x100-deg3"])
def test_polyreg_adversarial(eng, O, name):
    """VERDICT r1 weak #2 / ADVICE: the polynomial operator's moments form must hold the 1e-12 bar when the intercept
    is large against the residual spread and when the inputs are far from centred (BS:540-583 has no such hole).
    Walkers: prior draws, the least-squares fit and its neighbourhood (the posterior bulk, where N c0^2 >> Sum e^2),
    and points along the intercept/slope ridge."""
    c = _poly_adversarial(name)
    deg = c.iparam[0]
    gp, op, pr = _pair(eng, O, c)
    x, y = c.inputs[:, 0], c.outputs[:, 0]
    xm = x.mean()
    V = np.vander(x - xm, deg + 1, increasing=True)
    a = np.linalg.lstsq(V, y, rcond=None)[0]
    # coefficients in powers of x from those in powers of (x - xm)
    from math import comb
    coef = np.array([sum(a[j] * comb(j, k) * (-xm) ** (j - k) for j in range(k, deg + 1)) for k in range(deg + 1)])
    sg = np.sqrt(((y - V @ a) ** 2).mean())
    fit = np.concatenate([coef, [sg]])
    rng = np.random.default_rng(3)
    near = fit * (1 + 1e-7 * rng.standard_normal((16, deg + 2)))
    ridge = np.tile(fit, (8, 1))
    shift = np.linspace(-3, 3, 8)
    ridge[:, 0] += shift
    if abs(xm) > 1:
        ridge[:, 1] -= shift / xm
    th = np.vstack([pr.sample(40, 1), fit[None], near, ridge])
    got = gp.loglike(th)
    hi, lo = op.loglike_quad(th)
    ok = (hi > 0.5 * O.LOGZERO) & np.isfinite(hi)
    assert ok.sum() >= 40
    rel = np.abs((got[ok] - hi[ok]) - lo[ok]) / np.abs(hi[ok])
    assert rel.max() <= RTOL_LOGL, f"{name}: max rel err vs float128 oracle {rel.max():.3e} at {np.argmax(rel)}"
    # a nested-sampling walk on these data: every stored log-likelihood is the operator's value at the stored point
    # (the walk kernels score proposals through the same pivoted arithmetic as binest_loglike) ...
    start = np.vstack([near, ridge, pr.sample(40, 2)])
    opts = eng.default_options(pool_size=64, batch_k=8, mc_steps=10, max_iter=40, min_iter=40, seed=4)
    run = eng.RunGroup(gp, opts, start)
    assert run.advance(0)
    s = run.fetch(0)
    qh, ql = op.loglike_quad(s["points"])
    relw = np.abs((s["logL"] - qh) - ql) / np.abs(qh)
    assert relw.max() <= RTOL_LOGL, f"{name}: stored logL vs float128 oracle at the stored points {relw.max():.3e}"
    assert np.any(np.isfinite(s["acc"]) & (s["acc"] > 0))  # some walks moved
    if not name.startswith("x100"):
        # ... and the whole trajectory equals the oracle's.  (Not for the x in (100, 101) data: there intercept and
        # slope are almost perfectly anti-correlated over the live set, the proposal Cholesky factor amplifies the
        # last-bit differences between the device's tree sums and the oracle's sequential sums of the live-set
        # covariance by ~1e6, and the two walks part ways for reasons that have nothing to do with the likelihood.)
        ref = O.nested_sampling(op, pr, pool_size=64, batch_k=8, mc_steps=10, max_iter=40, min_iter=40, seed=4,
                                adapt_in_walk=False, start_points=start)
        assert s["M"] == ref.logL.size
        np.testing.assert_allclose(s["logL"], ref.logL, rtol=1e-9)


def test_softmax_adversarial_large_logits_and_separated_classes(eng, O):
    """C3 with |z| up to ~600 on linearly separated classes: log p_y = z_y - log Sum e^z is a difference of two
    numbers of size |z| per datum, N |z| over the data.  Bar: 1e-12 relative, or — where the likelihood is
    essentially 1 for every datum and |logL| itself is tiny — the fp64 floor of the per-datum probabilities the
    reference itself works with (p_y rounded to 1 - k ulp: 1.2e-16 per datum)."""
    g = np.random.default_rng(12)
    N, F, K = 20_000, 4, 3
    W = np.array([[3.0, -1.0, 0.5, 0.0], [-2.0, 2.5, 0.0, 1.0], [0.0, 0.0, 0.0, 0.0]])
    b = np.array([0.3, -0.2, 0.0])
    x = g.standard_normal((3 * N, F))
    z = x @ W.T + b
    zs = np.sort(z, 1)
    keep = (zs[:, 2] - zs[:, 1]) > 0.5           # margin: classes are linearly separated with a gap
    x, z = x[keep][:N], z[keep][:N]
    y = z.argmax(1).astype(np.float64)
    d = (K - 1) * (F + 1)
    c = cfg.Config("C3-separated", cfg.OP_LOGISTIC, x, y.reshape(-1, 1), (0, K, 0, 0), [f"p{i}" for i in range(d)],
                   [cfg.PRIOR_NORMAL_TRUNC] * d, [-1000.0] * d, [1000.0] * d, [0.0] * d, [5.0] * d)
    gp, op, pr = _pair(eng, O, c)
    t0 = np.concatenate([np.concatenate([W[k], [b[k]]]) for k in range(K - 1)])
    scales = np.array([0.5, 1, 3, 10, 30, 60, 100, 150])  # max |z| ~ 4 .. 1200 (the last ones take the slow path)
    th = np.vstack([t0[None] * s for s in scales] + [pr.sample(24, 2)])
    zmax = [np.abs(x @ t.reshape(K - 1, F + 1)[:, :F].T + t.reshape(K - 1, F + 1)[:, F]).max() for t in th]
    assert max(zmax[:8]) > 600 and min(zmax[:8]) < 10
    got = gp.loglike(th)
    hi, lo = op.loglike_quad(th)
    err = np.abs((got - hi) - lo)
    tol = RTOL_LOGL * np.abs(hi) + 1.2e-16 * x.shape[0]
    assert np.all(err <= tol), (err / tol).max()
    assert np.all(got <= 0.0)
    # where every datum is classified with probability 1 to fp64 precision the result is exactly representable noise-free
    assert abs(got[7]) < 1e-290 or got[7] == 0.0


def test_gbm_adversarial_tiny_time_step(eng, O):
    """GBM with dt = 1e-6 (log-increments ~ 2.5e-4, r / sqrt(dt) well scaled only after the upload transform)."""
    for dt in (1e-6, 1e-9):
        c = cfg.c4_gbm(T=20_000, seed=9, dt=dt)
        gp, op, pr = _pair(eng, O, c)
        th = np.vstack([pr.sample(40, 3), [[0.08, 0.25]], [[-0.9, 1.9]]])
        got = gp.loglike(th)
        hi, lo = op.loglike_quad(th)
        rel = np.abs((got - hi) - lo) / np.abs(hi)
        assert rel.max() <= RTOL_LOGL, (dt, rel.max())


def test_gp_adversarial_tiny_nugget(eng, O):
    """GP with nugget sigma_n^2 = 1e-6 sigma_f^2 (condition number ~ N 1e6): the marginal likelihood either factors
    and agrees with the long-double oracle to kappa-scaled accuracy, or is reported as logzero (GP:131-135) —
    never garbage.  A matrix that is not positive definite to working precision (nugget 1e-14) must give logzero
    or the oracle's value."""
    c0 = cfg.c5_gp(N=256)
    lo = list(c0.lo)
    lo[2] = 1e-9
    c = cfg.Config(c0.name, c0.op, c0.inputs, c0.outputs, c0.iparam, c0.names, c0.kinds, lo, c0.hi)
    gp, op, pr = _pair(eng, O, c)
    th = np.array([[1.0, 1.0, 1e-3], [2.0, 0.7, 2e-3], [0.5, 2.0, 5e-4], [1.0, 3.0, 1e-7], [1.0, 4.5, 1e-8]])
    got = gp.loglike(th)
    hi, _ = op.loglike_quad(th)
    for k in range(th.shape[0]):
        if got[k] == O.LOGZERO:
            continue  # reported as not factorable: the reference Throws to logzero for singular matrices
        assert np.isfinite(got[k])
        # k >= 3: nugget 1e-14 / 1e-16, kappa beyond 1/eps — the reference's own LU arithmetic (GP:130-141) is off by
        # 6e-3 / has the wrong sign there; only "the right order of magnitude, or logzero" can be asked
        assert abs(got[k] - hi[k]) <= (1e-8 if k < 3 else 0.5) * abs(hi[k]), (k, got[k], hi[k])
    assert np.all(got[:3] != O.LOGZERO), "nugget 1e-6 sigma_f^2 at N = 256 must factor"


def test_loglike_empty_and_tiny(eng, O):
    c = cfg.c1_gaussian(N=1, seed=3)
    gp, op, pr = _pair(eng, O, c)
    assert gp.loglike(np.empty((0, 2))).shape == (0,)
    th = pr.sample(5, 1)
    np.testing.assert_allclose(gp.loglike(th), op.loglike(th, pr), rtol=1e-13)
    for N in (2, 3, 65, 129):
        c = cfg.c2_polyreg(N=N)
        gp, op, pr = _pair(eng, O, c)
        th = pr.sample(40, 2)
        np.testing.assert_allclose(gp.loglike(th), op.loglike(th, pr), rtol=1e-12)


def test_loglike_full_size_pins_and_additivity(eng, O):
    """BASELINE size (N = 1e6): known-answer pin from SURVEY Appendix A and the size-independent
    property logL(data) = logL(first half) + logL(second half)."""
    c = cfg.c2_polyreg()
    gp = eng.Problem.from_config(c)
    x, y = c.inputs[:, 0], c.outputs[:, 0]
    X = np.vander(x, 4, increasing=True)
    coef, *_ = np.linalg.lstsq(X, y, rcond=None)
    s = np.sqrt(((y - X @ coef) ** 2).sum() / x.size)
    th = np.concatenate([coef, [s]])[None, :]
    got = gp.loglike(th)[0]
    assert abs(got - (-31470.229840)) < 2e-6
    h = x.size // 2 + 1
    ca = cfg.Config(c.name, c.op, c.inputs[:h], c.outputs[:h], c.iparam, c.names, c.kinds, c.lo, c.hi)
    cb = cfg.Config(c.name, c.op, c.inputs[h:], c.outputs[h:], c.iparam, c.names, c.kinds, c.lo, c.hi)
    pr = O.Prior(c.kinds, c.lo, c.hi)
    ths = pr.sample(64, 5)
    full = gp.loglike(ths)
    parts = eng.Problem.from_config(ca).loglike(ths) + eng.Problem.from_config(cb).loglike(ths)
    np.testing.assert_allclose(full, parts, rtol=1e-12)
    # GBM pin
    c4 = cfg.c4_gbm()
    g4 = eng.Problem.from_config(c4)
    assert abs(g4.loglike([[0.08, 0.25]])[0] - (-72297.49504288615)) < 1e-7


def test_prior_and_philox(eng, O):
    for c in (cfg.c1_gaussian(), cfg.c3_logistic(N=100)):
        gp, op, pr = _pair(eng, O, c)
        a = gp.sample_prior(257, seed=11, run_id=3)
        b = pr.sample(257, 11, 3)
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-13)  # same Philox words, libm vs libdevice exp/log
        th = np.vstack([b, b * 1.7])
        np.testing.assert_allclose(gp.logprior(th), pr.logpdf(th), rtol=1e-13)


def test_crude_weights_and_evidence_sampling(eng, O):
    rng = np.random.default_rng(5)
    n, nd, d = 100, 900, 2
    M = n + nd
    logL = np.sort(rng.normal(-120, 8, M))
    pts = rng.normal(size=(M, d))
    pool = np.concatenate([np.full(nd, n), np.arange(n, 0, -1)]).astype(np.int64)
    g = eng.crude_weights(logL, pool, n)
    lx = O.xvalues_log(n, nd, pool)
    lw = O.trapezoid_log(lx) + logL
    np.testing.assert_allclose(g["logX"], lx, rtol=1e-13)
    np.testing.assert_allclose(g["crude_logw"], lw, rtol=1e-12)
    z = O.logsumexp(lw)
    assert abs(g["crude_logZ"] - z) < 1e-11
    assert abs(g["entropy"] - O.entropy(lw, logL, z)) < 1e-10
    # reference X sequence BS:785-799 (constant pool): K3 pin of SURVEY 8c
    np.testing.assert_allclose(g["logX"], O.xvalues_log(n, nd), rtol=1e-12)
    # Monte-Carlo error estimate: same Philox draws on both sides
    pool_b = pool.copy()
    pool_b[:nd] = n - (np.arange(nd) % 8)  # batched replacement pool sizes
    ev = eng.evidence_sampling(pts, logL, pool_b, n, 50, seed=9)
    ref = O.evidence_sampling(pts, logL, pool_b, n, 50, seed=9)
    np.testing.assert_allclose(ev["z"], ref["zSamples"], rtol=1e-11)
    np.testing.assert_allclose(ev["pmean"], ref["parameterSamples"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(ev["slx_mean"], ref["SampledLogX"]["Mean"], rtol=1e-10)
    np.testing.assert_allclose(ev["logw_mean"], ref["LogPosteriorWeight"]["Mean"], rtol=1e-9)
    np.testing.assert_allclose(ev["logw_sd"], ref["LogPosteriorWeight"]["StandardError"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("K", [1, 8])
def test_engine_trajectory_matches_oracle(eng, O, K):
    """Same seed, same Philox addressing, proposal factor frozen per iteration on both sides: the whole
    nested-sampling trajectory (dead list, live set, crude evidence) must coincide."""
    c = cfg.c1_gaussian()
    gp, op, pr = _pair(eng, O, c)
    start = pr.sample(100, 21, 0)
    opts = eng.default_options(pool_size=100, batch_k=K, mc_steps=40, max_iter=400, min_iter=100, seed=21)
    run = eng.RunGroup(gp, opts, start)
    assert run.advance(0)
    got = run.fetch(0)
    ref = O.nested_sampling(op, pr, pool_size=100, batch_k=K, mc_steps=40, max_iter=400, min_iter=100, seed=21,
                            adapt_in_walk=False, start_points=start)
    assert got["M"] == ref.logL.size and got["n_deleted"] == ref.n_deleted and got["iterations"] == ref.iterations
    assert np.array_equal(got["pool"], ref.pool)
    np.testing.assert_allclose(got["logL"], ref.logL, rtol=1e-9)
    np.testing.assert_allclose(got["points"], ref.points, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(got["acc"][~np.isnan(ref.acc)], ref.acc[~np.isnan(ref.acc)], rtol=1e-12)
    assert np.array_equal(np.isnan(got["acc"]), np.isnan(ref.acc))
    np.testing.assert_allclose(got["logX"], ref.logX, rtol=1e-12)
    assert abs(got["crude_logZ"] - ref.crude_logZ) < 1e-8
    assert abs(got["entropy"] - ref.entropy) < 1e-7


WALK_PATHS = {
    "resident-cluster": {},                                                 # walk_resident.cuh
    "grid-2sets": {"BINEST_NO_RESIDENT": "1"},                              # walk_grid.cuh, alternating sets
    "grid-1set": {"BINEST_NO_RESIDENT": "1", "BINEST_GRID_SETS": "1"},
    "stepped-graph": {"BINEST_NO_RESIDENT": "1", "BINEST_NO_GRID": "1"},    # [walk_step, loglike_stream] x S
    "stepped-graph-nopdl": {"BINEST_NO_RESIDENT": "1", "BINEST_NO_GRID": "1", "BINEST_NO_PDL": "1"},
}


@pytest.mark.parametrize("path", list(WALK_PATHS))
@pytest.mark.parametrize("case", ["C2-20k", "C4-2runs"])
def test_every_walk_path_matches_oracle(eng, O, monkeypatch, path, case):
    """The three device implementations of the S-step walk (cluster-resident, grid-resident persistent, stepped
    CUDA graph) share walk_step_walker / the Philox addressing, so each must reproduce the oracle trajectory."""
    for k in ("BINEST_NO_RESIDENT", "BINEST_NO_GRID", "BINEST_GRID_SETS", "BINEST_NO_PDL"):
        monkeypatch.delenv(k, raising=False)
    for k, v in WALK_PATHS[path].items():
        monkeypatch.setenv(k, v)
    if case == "C2-20k":
        c, n, K, S, iters, runs = cfg.c2_polyreg(N=20_000), 256, 96, 25, 1440, 1
    else:
        c, n, K, S, iters, runs = cfg.c4_gbm(T=3000), 128, 24, 30, 600, 2
    gp, op, pr = _pair(eng, O, c)
    opts = eng.default_options(pool_size=n, batch_k=K, mc_steps=S, max_iter=iters, min_iter=iters, seed=33, n_runs=runs)
    start = np.stack([pr.sample(n, 33, r) for r in range(runs)])
    run = eng.RunGroup(gp, opts, start)
    assert run.advance(0)
    for r in range(runs):
        got = run.fetch(r)
        ref = O.nested_sampling(op, pr, pool_size=n, batch_k=K, mc_steps=S, max_iter=iters, min_iter=iters, seed=33,
                                adapt_in_walk=False, start_points=start[r], run_id=r)
        assert got["M"] == ref.logL.size and got["iterations"] == ref.iterations
        np.testing.assert_allclose(got["logL"], ref.logL, rtol=1e-9)
        np.testing.assert_allclose(got["points"], ref.points, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(got["acc"][~np.isnan(ref.acc)], ref.acc[~np.isnan(ref.acc)], rtol=1e-12)
        assert abs(got["crude_logZ"] - ref.crude_logZ) < 1e-9 * abs(ref.crude_logZ)
    run.close()


@pytest.mark.parametrize("K,runs,n,iters", [(1, 1, 100, 600), (8, 3, 1000, 4800), (32, 2, 100, 640)])
def test_device_resident_loop_matches_oracle(eng, O, monkeypatch, K, runs, n, iters):
    """walk_loop.cuh: the whole nested-sampling loop in one launch (warp-per-walker walks speculating over rejections,
    update by the same CTA) on the C1-shaped problem it is meant for — whole trajectories against the oracle for the
    reference scheme (K = 1) and for batches, several lock-step runs; the same run through the per-iteration kernels
    gives the same samples; an advance in pieces (max_batches), a fetch in between and — n = 1000, 4800 removals: more
    than the initial dead capacity of 4096 — the dead-list growth inside the loop change nothing.  (Runs are kept to a
    depth of a few nats: far past the posterior bulk all live points tie in logL to the last bits and the removal
    order, hence the trajectory, depends on the last ulp of libm vs libdevice — on every walk path alike.)"""
    monkeypatch.delenv("BINEST_NO_LOOP", raising=False)
    c = cfg.c1_gaussian()
    gp, op, pr = _pair(eng, O, c)
    S = 40
    opts = eng.default_options(pool_size=n, batch_k=K, mc_steps=S, max_iter=iters, min_iter=iters, seed=21, n_runs=runs)
    start = np.stack([pr.sample(n, 21, r) for r in range(runs)])
    run = eng.RunGroup(gp, opts, start)
    assert run.walk_path() == "device-loop"
    assert not run.advance(3)
    mid = run.fetch(0)
    assert mid["n_deleted"] == min(3 * K, iters)
    assert run.advance(0)
    res = [run.fetch(r) for r in range(runs)]
    run.close()
    for r in range(runs):
        ref = O.nested_sampling(op, pr, pool_size=n, batch_k=K, mc_steps=S, max_iter=iters, min_iter=iters, seed=21,
                                adapt_in_walk=False, start_points=start[r], run_id=r)
        got = res[r]
        assert got["M"] == ref.logL.size and got["iterations"] == ref.iterations
        np.testing.assert_allclose(got["logL"], ref.logL, rtol=1e-9)
        np.testing.assert_allclose(got["points"], ref.points, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(got["acc"][~np.isnan(ref.acc)], ref.acc[~np.isnan(ref.acc)], rtol=1e-12)
        assert abs(got["crude_logZ"] - ref.crude_logZ) < 1e-9 * abs(ref.crude_logZ)
    monkeypatch.setenv("BINEST_NO_LOOP", "1")
    run2 = eng.RunGroup(gp, opts, start)
    assert run2.walk_path() == "cluster-resident"
    assert run2.advance(0)
    other = run2.fetch(runs - 1)
    run2.close()
    assert other["M"] == res[-1]["M"]
    np.testing.assert_allclose(other["logL"], res[-1]["logL"], rtol=1e-9)


def test_device_resident_loop_terminates_like_the_stepped_engine(eng, monkeypatch):
    """Termination inside the device loop (BS:967-978 in the log domain): a full C1 run, reference scheme, stops at the
    same iteration with the same evidence as the per-iteration engine; and it is the fast path (no host round trip
    per iteration)."""
    import time
    c = cfg.c1_gaussian()
    gp = eng.Problem.from_config(c)
    out = {}
    for mode in ("loop", "stepped"):
        if mode == "stepped":
            monkeypatch.setenv("BINEST_NO_LOOP", "1")
        else:
            monkeypatch.delenv("BINEST_NO_LOOP", raising=False)
        opts = eng.default_options(pool_size=100, batch_k=1, mc_steps=200, max_iter=10**6, seed=5)
        run = eng.RunGroup(gp, opts)
        t0 = time.perf_counter()
        assert run.advance(0)
        dt = time.perf_counter() - t0
        s = run.fetch(0)
        out[mode] = (s, dt, run.sizes(0))
        run.close()
    a, b = out["loop"][0], out["stepped"][0]
    assert a["M"] == b["M"] and abs(a["crude_logZ"] - b["crude_logZ"]) < 1e-8
    assert abs(a["crude_logZ"] - c.truth["logZ"]) < 1.5  # ~ 5 sigma of sqrt(H/n) = 0.27
    rate = {m: out[m][0]["n_deleted"] / out[m][1] for m in out}
    print(f"C1 reference scheme (K = 1, S = 200): device loop {rate['loop']:.0f} replacements/s, "
          f"per-iteration kernels {rate['stepped']:.0f} replacements/s")
    assert rate["loop"] > 2.0 * rate["stepped"]


@pytest.mark.parametrize("path", ["resident-cluster", "grid-2sets", "stepped-graph"])
def test_acceptance_range_loops_match_oracle(eng, O, monkeypatch, path):
    """"MinMaxAcceptanceRate" (BS:848): the inner loop of nsMCMC (extra S-step blocks until the rate is in range or
    5S steps, BS:730-736) and the outer retry of nestedSamplingInternal (restart from a fresh live point with
    Ceiling[1.25^k S] steps, BS:995-1004), on every walk path, against the oracle trajectory."""
    for k in ("BINEST_NO_RESIDENT", "BINEST_NO_GRID", "BINEST_GRID_SETS", "BINEST_NO_PDL"):
        monkeypatch.delenv(k, raising=False)
    for k, v in WALK_PATHS[path].items():
        monkeypatch.setenv(k, v)
    c = cfg.c1_gaussian(N=400, seed=5)
    gp, op, pr = _pair(eng, O, c)
    n, K, S, iters = 96, 12, 16, 360
    start = pr.sample(n, 44, 0)
    opts = eng.default_options(pool_size=n, batch_k=K, mc_steps=S, max_iter=iters, min_iter=iters, seed=44,
                               acc_min=0.2, acc_max=0.45)
    run = eng.RunGroup(gp, opts, start)
    assert run.advance(0)
    got = run.fetch(0)
    evals = run.sizes(0)["evals"]
    run.close()
    ref = O.nested_sampling(op, pr, pool_size=n, batch_k=K, mc_steps=S, max_iter=iters, min_iter=iters, seed=44,
                            acc_range=(0.2, 0.45), adapt_in_walk=False, start_points=start)
    assert got["M"] == ref.logL.size and got["iterations"] == ref.iterations
    np.testing.assert_allclose(got["logL"], ref.logL, rtol=1e-9)
    np.testing.assert_allclose(got["points"], ref.points, rtol=1e-7, atol=1e-10)
    a, b = got["acc"][~np.isnan(ref.acc)], ref.acc[~np.isnan(ref.acc)]
    np.testing.assert_allclose(a, b, rtol=1e-12)
    # the loops really ran: more evaluations than the plain S per replacement, and nearly all rates inside the range
    assert evals > 1.3 * (n + iters * S)
    assert np.mean((b >= 0.2) & (b <= 0.45)) > 0.9


@pytest.mark.parametrize("path", ["grid-2sets", "grid-1set", "stepped-graph", "stepped-graph-nopdl"])
def test_full_size_walk_paths_agree_and_store_true_loglike(eng, monkeypatch, path):
    """C2 at BASELINE size (1e6 rows, n = 1024, K = 256, S = 200), where the oracle is too slow: size-independent
    properties instead.  (1) every stored sample's logL equals the operator evaluated at the stored point (a stale
    proposal or a missed partial sum breaks this at once: it caught a PDL race at this size); (2) all walk paths
    produce the same trajectory; (3) dead logL ascending, live points above the last threshold."""
    for k in ("BINEST_NO_RESIDENT", "BINEST_NO_GRID", "BINEST_GRID_SETS", "BINEST_NO_PDL"):
        monkeypatch.delenv(k, raising=False)
    for k, v in WALK_PATHS[path].items():
        monkeypatch.setenv(k, v)
    c = cfg.c2_polyreg()
    gp = eng.Problem.from_config(c)
    opts = eng.default_options(pool_size=1024, batch_k=256, mc_steps=200, max_iter=10**9, min_iter=10**9, seed=2026)
    run = eng.RunGroup(gp, opts)
    run.advance(4)
    s = run.fetch(0)
    run.close()
    nd = s["n_deleted"]
    assert s["M"] == 1024 + 4 * 256 and nd == 4 * 256
    chk = gp.loglike(s["points"])
    np.testing.assert_allclose(s["logL"], chk, rtol=1e-12)
    assert np.all(np.diff(s["logL"][:nd]) >= 0) and np.all(s["logL"][nd:] > s["logL"][nd - 1])
    acc = s["acc"][~np.isnan(s["acc"])]
    assert acc.size == 4 * 256 and 0.02 < acc.mean() < 0.9
    key = ("full-size-c2", 2026)
    ref = _FULL_SIZE_TRAJ.setdefault(key, s)
    np.testing.assert_allclose(s["logL"], ref["logL"], rtol=1e-11)
    np.testing.assert_allclose(s["points"], ref["points"], rtol=1e-9, atol=1e-12)


_FULL_SIZE_TRAJ = {}


def test_engine_logz_c1(eng, O):
    """LogEvidence within 3 sigma of the quadrature value (K4 pin, SURVEY Appendix A: -114.641064)."""
    c = cfg.c1_gaussian()
    gp = eng.Problem.from_config(c)
    pulls = []
    for seed in (1, 2, 3, 4):
        opts = eng.default_options(pool_size=100, batch_k=8, mc_steps=200, max_iter=100000, seed=seed)
        run = eng.RunGroup(gp, opts)
        assert run.advance(0)
        s = run.fetch(0)
        ev = eng.evidence_sampling(s["points"], s["logL"], s["pool"], 100, 100, seed)
        mu, sd = ev["z"].mean(), ev["z"].std(ddof=1)
        assert 0.15 < sd < 0.45
        pulls.append((mu - c.truth["logZ"]) / sd)
    assert max(abs(p) for p in pulls) < 3.0, pulls
    assert abs(np.mean(pulls)) < 1.5, pulls


def test_numeric_loglikelihood_maximum_matches_oracle(eng, O):
    """binest_options.loglmax ("LogLikelihoodMaximum" -> number, BS:925-932) in run_update_kernel's termination test:
    same stopping iteration and samples as the oracle, longer than Automatic for a larger value."""
    c = cfg.c1_gaussian()
    gp, op, pr = _pair(eng, O, c)
    start = pr.sample(50, 3)
    n_it = {}
    for lm in (float("nan"), -98.0, -400.0):
        opts = eng.default_options(pool_size=50, batch_k=1, mc_steps=20, max_iter=5000, min_iter=30, seed=6, loglmax=lm)
        run = eng.RunGroup(gp, opts, start)
        assert run.advance(0)
        s = run.fetch(0)
        ref = O.nested_sampling(op, pr, pool_size=50, batch_k=1, mc_steps=20, max_iter=5000, min_iter=30, seed=6,
                                adapt_in_walk=False, start_points=start, loglmax=lm)
        assert s["M"] == ref.logL.size, (lm, s["M"], ref.logL.size)
        np.testing.assert_allclose(s["logL"], ref.logL, rtol=1e-9)
        n_it[lm if lm == lm else "auto"] = s["n_deleted"]
    assert n_it[-400.0] == 30 and n_it[-98.0] > n_it["auto"]


def test_multi_run_group_is_shard_invariant(eng):
    """Run r of a group equals the same run executed alone (results depend on (seed, run id) only)."""
    c = cfg.c4_gbm(T=512)
    gp = eng.Problem.from_config(c)
    o4 = eng.default_options(pool_size=64, batch_k=8, mc_steps=30, max_iter=300, seed=5, n_runs=4, first_run_id=0)
    g = eng.RunGroup(gp, o4)
    assert g.advance(0)
    o1 = eng.default_options(pool_size=64, batch_k=8, mc_steps=30, max_iter=300, seed=5, n_runs=1, first_run_id=2)
    s = eng.RunGroup(gp, o1)
    assert s.advance(0)
    a, b = g.fetch(2), s.fetch(0)
    assert a["M"] == b["M"]
    np.testing.assert_allclose(a["points"], b["points"], rtol=1e-12)
    np.testing.assert_allclose(a["logL"], b["logL"], rtol=1e-12)


@pytest.mark.parametrize("N", [200, 512])
def test_gp_marginal_likelihood_matches_oracle(eng, O, N):
    """C5 operator (K7 pin): batched covariance fill + blocked Cholesky (DMMA trailing update) against the
    oracle's LU path (GP:130-141) and its long-double Cholesky.  kappa(K) <~ 1e4 here; tolerance 1e-10 rel."""
    c = cfg.c5_gp(N=N)
    gp, op, pr = _pair(eng, O, c)
    th = pr.sample(12, 31)
    th[0] = [1.0, 0.8, 0.3]
    got = gp.loglike(th)
    ref = op.loglike(th, pr)
    hi, lo = op.loglike_quad(th)
    np.testing.assert_allclose(got, hi, rtol=1e-10)
    np.testing.assert_allclose(ref, hi, rtol=1e-9)
    bad = th.copy()
    bad[2, 1] = -1.0   # outside the box -> logzero
    bad[3, 2] = 0.0
    out = gp.loglike(bad)
    assert out[2] == O.LOGZERO and out[3] == O.LOGZERO
    np.testing.assert_allclose(out[[0, 1, 4]], got[[0, 1, 4]], rtol=1e-13)


def _gp_logl_lapack(c, th):
    """fp64 LAPACK Cholesky of the same covariance (scipy.linalg.cho_factor): -1/2 (N log 2pi + logdet + r.K^-1 r)."""
    import scipy.linalg as sl
    x, y = c.inputs[:, 0], c.outputs[:, 0]
    N = x.size
    d2 = (x[:, None] - x[None, :]) ** 2
    out = []
    for sf, ell, sn in th:
        K = sf * sf * np.exp(-d2 / (2.0 * ell * ell))
        K[np.diag_indices(N)] += sn * sn
        L = sl.cholesky(K, lower=True, overwrite_a=True, check_finite=False)
        z = sl.solve_triangular(L, y, lower=True, check_finite=False)
        out.append(-0.5 * (N * np.log(2.0 * np.pi) + 2.0 * np.log(np.diag(L)).sum() + z @ z))
    return np.array(out)


def test_gp_c5_at_its_stated_size(eng, O):
    """BASELINE config C5 as stated: N = 4096 (32 panels of 128, the two-level blocked sweep with its binary in-group
    schedule and full-width K = 1024 updates).  Six hyper-parameter sets against LAPACK's fp64 Cholesky of the same
    matrix, one of them also against the oracle's long-double Cholesky (GP:130-141 restated, ~30 s on the host);
    tolerance 1e-10 relative (kappa(K) <~ 1e4 for these nuggets)."""
    c = cfg.c5_gp()
    assert c.inputs.shape[0] == 4096
    gp, op, pr = _pair(eng, O, c)
    th = np.vstack([[1.0, 0.8, 0.3], [0.7, 1.5, 0.1], [2.0, 0.4, 0.5], [0.3, 3.0, 0.05], pr.sample(2, 77)])
    got = gp.loglike(th)
    ref = _gp_logl_lapack(c, th)
    np.testing.assert_allclose(got, ref, rtol=1e-10)
    hi, _ = op.loglike_quad(th[:1])
    assert abs(got[0] - hi[0]) <= 1e-10 * abs(hi[0]), (got[0], hi[0])
    # a batch larger than one sub-batch split, with box violations inside it: values do not depend on the batch
    big = np.vstack([th, th[::-1], th])
    big[7, 2] = 0.0
    out = gp.loglike(big)
    assert out[7] == O.LOGZERO
    np.testing.assert_array_equal(np.delete(out, 7), np.delete(np.concatenate([got, got[::-1], got]), 7))


def test_gp_not_positive_definite_gives_logzero(eng, O):
    """matrixInverseAndDet Throws for a singular matrix and the likelihood becomes logzero (GP:131-135, 190-197).  With
    a nugget far below eps * sigma_f^2 and a long length scale the covariance has numerical rank ~15: a pivot of the
    Cholesky sweep comes out <= 0 and the value must be exactly logzero — in the first panel (N = 100: diagonal-block
    kernel only) and in a later one (N = 700)."""
    for N in (100, 700):
        c0 = cfg.c5_gp(N=N)
        lo = list(c0.lo)
        lo[2] = 1e-12
        x = c0.inputs
        if N == 700:
            # first panel (128 inputs, spacing 7.9 >> ell) is well conditioned; the dense block behind it is not
            x = np.concatenate([np.linspace(0.0, 1000.0, 128), 2000.0 + np.sort(c0.inputs[:572, 0])]).reshape(-1, 1)
        c = cfg.Config(c0.name, c0.op, x, c0.outputs, c0.iparam, c0.names, c0.kinds, lo, c0.hi)
        gp, op, pr = _pair(eng, O, c)
        th = np.array([[1.0, 4.5, 1e-10], [1.0, 0.8, 0.3], [3.0, 4.9, 1e-11]])
        got = gp.loglike(th)
        assert got[0] == O.LOGZERO and got[2] == O.LOGZERO, got
        hi, _ = op.loglike_quad(th[1:2])
        assert abs(got[1] - hi[0]) <= 1e-10 * abs(hi[0])  # the healthy matrix of the same batch is unaffected


def test_gp_nested_sampling_small(eng, O):
    """A short GP run end to end: engine trajectory equals the oracle's (same seed, frozen proposals)."""
    c = cfg.c5_gp(N=96)
    gp, op, pr = _pair(eng, O, c)
    start = pr.sample(24, 8, 0)
    opts = eng.default_options(pool_size=24, batch_k=6, mc_steps=8, max_iter=36, min_iter=36, seed=8)
    run = eng.RunGroup(gp, opts, start)
    assert run.advance(0)
    got = run.fetch(0)
    ref = O.nested_sampling(op, pr, pool_size=24, batch_k=6, mc_steps=8, max_iter=36, min_iter=36, seed=8,
                            adapt_in_walk=False, start_points=start)
    assert got["M"] == ref.logL.size
    np.testing.assert_allclose(got["logL"], ref.logL, rtol=1e-8)
    np.testing.assert_allclose(got["points"], ref.points, rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("N,Q", [(200, 1), (200, 37), (512, 300), (130, 129)])
def test_gp_predict_matches_oracle(eng, O, N, Q):
    """predictFromGaussianProcess (GP:332-422): predictive mean / sd of every parameter vector at
    every input from the augmented Cholesky sweep, against the oracle's long-double Cholesky and its fp64 LU
    restatement of GP:395-420.  Ragged sizes: N and Q not multiples of the 128-wide panel."""
    c = cfg.c5_gp(N=N)
    gp, op, pr = _pair(eng, O, c)
    th = pr.sample(9, 41)
    th[0] = [1.0, 0.8, 0.3]
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-1.0, 11.0, Q - 1), c.inputs[3]])  # one input coincides with a datum
    mean, sd = gp.gp_predict(th, xs)
    rm, rs = op.gp_predict(th, xs, long_double=True)
    lm, ls = op.gp_predict(th, xs)
    scale = np.abs(rm).max()
    assert np.abs(mean - rm).max() <= 1e-9 * scale, np.abs(mean - rm).max()
    np.testing.assert_allclose(sd, rs, rtol=1e-7, atol=1e-9)
    # the reference's own LU arithmetic is no closer to the long-double value than the GPU sweep is (x10 slack)
    assert np.abs(mean - rm).max() <= 10 * np.abs(lm - rm).max() + 1e-13 * scale
    # a box/constraint violation or an unfactorable covariance gives NaN for that parameter vector only
    bad = th.copy()
    bad[2, 1] = -1.0
    m2, s2 = gp.gp_predict(bad, xs)
    assert np.isnan(m2[2]).all() and np.isnan(s2[2]).all()
    np.testing.assert_allclose(m2[[0, 1, 3]], mean[[0, 1, 3]], rtol=1e-13, atol=1e-15)
    # the likelihood entry point shares the workspace and the sweep: unchanged after a prediction call
    np.testing.assert_allclose(gp.loglike(th), op.loglike_quad(th)[0], rtol=1e-10)


def test_predict_from_gaussian_process_api(eng):
    """The reference-facing call on a finished run: mixture over all samples with their CrudePosteriorWeight."""
    from bayesianinference_b200 import api
    c = cfg.c5_gp(N=256)
    obj = api.defineGaussianProcess((c.inputs[:, 0], c.outputs[:, 0]), api.SquaredExponentialGP(*c.names),
                                    [(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)], ["ScaleParameter"] * 3)
    res = api.nestedSampling(obj, SamplePoolSize=128, BatchSize=32, MaxIterations=10**6, Seed=9)
    pred = api.predictFromGaussianProcess(res, 25)
    xs = pred.points[:, 0]
    truth = np.sin(xs) + 0.5 * np.cos(2.3 * xs)
    mu = np.array([pred[float(v)].mean() for v in xs])
    sdv = np.array([pred[float(v)].sd() for v in xs])
    assert np.all(np.isfinite(mu)) and np.all(sdv > 0.05) and np.all(sdv < 0.4)  # noise sd 0.1 + function uncertainty
    assert np.abs(mu - truth).max() < 0.25 and np.mean(np.abs(mu - truth) < 2.5 * sdv) > 0.9


def test_predictive_components_match_oracle(eng, O):
    """predictiveDistribution (BS:1437-1483): (mean, sd) / class-probability tables of every
    sample at every input against the oracle restatement; ragged M, Q."""
    rng = np.random.default_rng(11)
    c = cfg.c2_polyreg(N=1000)
    gp, op, pr = _pair(eng, O, c)
    th = pr.sample(333, 4)
    xs = rng.uniform(-1, 1, 77)
    got, ref = gp.predictive_components(th, xs), op.predictive_components(th, xs)
    assert got.shape == (333, 77, 2)
    np.testing.assert_allclose(got[:, :, 0], ref[:, :, 0], rtol=1e-14, atol=1e-14)  # Horner with FMA vs without; |c_j| <= 5
    np.testing.assert_array_equal(got[:, :, 1], ref[:, :, 1])
    th[5, 4] = -1.0
    assert np.isnan(gp.predictive_components(th, xs)[5]).all()
    c = cfg.c3_logistic(N=1000)
    gp, op, pr = _pair(eng, O, c)
    th = pr.sample(65, 4)
    X = rng.normal(size=(31, 4))
    got, ref = gp.predictive_components(th, X), op.predictive_components(th, X)
    assert got.shape == (65, 31, 3)
    np.testing.assert_allclose(got, ref, rtol=2e-12, atol=1e-300)  # |z| up to ~100: FMA vs plain dot in the logits
    np.testing.assert_allclose(got.sum(-1), 1.0, rtol=1e-14)
    with pytest.raises(ValueError):
        eng.Problem.from_config(cfg.c4_gbm(T=64)).predictive_components(np.ones((1, 2)), np.ones(3))


def test_mcmc_chain_trajectory_matches_oracle(eng, O):
    """createMCMCChain / iterateMCMC (BS:630-703): the device chains and the oracle restatement
    share the Philox addressing, so whole trajectories agree — across two iterate calls, for several chains."""
    c = cfg.c2_polyreg(N=2000)
    gp, op, pr = _pair(eng, O, c)
    start = np.array([[0.5, -1.2, 0.8, 0.3, 0.25], [0.4, -1.0, 0.7, 0.2, 0.3], [0.6, -1.3, 0.9, 0.4, 0.22]])
    cov0 = np.diag([1e-4, 4e-4, 9e-4, 9e-4, 1e-5])
    ch = eng.Chain(gp, start, cov0, learn_delay=20, seed=17)
    got = np.concatenate([ch.iterate(100), ch.iterate(200)])
    st = ch.state()
    assert got.shape == (300, 3, 5)
    for k in range(3):
        ref = O.mcmc_chain(op, pr, start[k], cov0, delay=20, seed=17, chain_id=k, n_steps=300)
        np.testing.assert_allclose(got[:, k], ref["states"], rtol=1e-9, atol=1e-12)
        assert st["t"][k] == ref["t"] == 301 and st["accepted"][k] == ref["accepted"]
        np.testing.assert_allclose(st["mean"][k], ref["mean"], rtol=1e-9)
        np.testing.assert_allclose(st["cov"][k], ref["cov"], rtol=1e-6, atol=1e-14)
        assert 10 < ref["accepted"] < 290
    from bayesianinference_b200 import _lib
    with pytest.raises(_lib.BinestError) as e:
        eng.Chain(gp, [[0.5, -1.2, 0.8, 0.3, -1.0]], cov0)  # sigma < 0: outside the support
    assert e.value.code == 8


def test_mcmc_chains_sample_the_full_size_c2_posterior(eng):
    """32 chains on the N = 1e6 polynomial-regression posterior, started at the Laplace mode: pooled moments against the
    Laplace (Gaussian to O(1/N)) mean and covariance obtained from the same GPU operators."""
    import json
    import os
    from bayesianinference_b200 import api
    pins = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "laplace_pins.json")))["C2"]
    c = cfg.c2_polyreg()
    obj = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), IndependentVariables=["x"],
        GeneratingDistribution=api.NormalDistribution(api.Polynomial("x", tuple(c.names[:4])), "sigma"),
        Parameters=[(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)],
        PriorDistribution=["LocationParameter"] * 4 + ["ScaleParameter"])
    lap = api.approximateEvidence(obj, InitialGuess=pins["mode"])
    assert abs(lap["LogEvidence"] - pins["logZ_laplace"]) < 1e-4
    Sigma = np.linalg.inv(lap["PrecisionMatrix"])
    sd = np.sqrt(np.diag(Sigma))
    mode = np.array(pins["mode"])
    nch = 32
    ch = api.createMCMCChain(obj, np.tile(mode, (nch, 1)), Chains=nch, InitialCovariance=Sigma * (2.4**2 / 5), CovarianceLearnDelay=300, Seed=8)
    import time
    api.iterateMCMC(ch, 300)                     # burn-in + covariance learning
    t0 = time.perf_counter()
    s = api.iterateMCMC(ch, 1500)                # (1500, 32, 5)
    dt = time.perf_counter() - t0
    print(f"posterior MCMC, C2 N=1e6: 32 chains x 1500 steps in {dt:.3f} s = {1500 / dt:.0f} steps/s, "
          f"{32 * 1500 / dt:.3g} likelihood evaluations/s")
    rate = ch["AcceptanceRate"]
    assert np.all(rate > 0.05) and np.all(rate < 0.8), rate  # each chain adapts to its own (noisy) covariance estimate
    pooled = s.reshape(-1, 5)
    # ~ 32 x 1500 / (2 x integrated autocorrelation ~ 30) ~ 1000+ effective draws: mean within 0.15 sd, sd within 12 %
    assert np.all(np.abs(pooled.mean(0) - mode) < 0.15 * sd), (pooled.mean(0) - mode) / sd
    assert np.all(np.abs(pooled.std(0) / sd - 1.0) < 0.12), pooled.std(0) / sd
    corr = np.corrcoef(pooled.T)
    ref = Sigma / np.outer(sd, sd)
    assert np.abs(corr - ref).max() < 0.15
