"""Statistical calibration of the engine (SURVEY §8c pin K10): over >= 20 independent seeds the pull
(LogEvidence - truth) / reported sigma must have |mean| < 0.5 and a standard deviation in [0.6, 1.6] — for the
reference scheme (one replacement per iteration, BS:980-1018) and for the batched schemes production runs use
(K worst points replaced per iteration with per-sample pool sizes, DESIGN §2 divergence 2), whose X statistics are
NOT the reference's and have to earn their keep here.  Truth values: 2-D quadrature of prior x likelihood, the
operation directPosteriorDistribution performs (BS:114-126): C1 -114.641064, C4 -72306.535014.

The expected spread of the mean pull over S seeds is 1/sqrt(S) (0.2 at S = 24); the band on the standard deviation is
wide because the reported sigma (the spread of the X-sequence resampling, BS:1254) does not contain the walk's own
contribution."""
import numpy as np
import pytest

from bayesianinference_b200 import api
from bayesianinference_b200 import configs as cfg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from bayesianinference_b200 import engine
    engine.init()
    return engine


def _pulls(eng, gp, truth, n, K, runs, seed):
    opts = eng.default_options(pool_size=n, batch_k=K, mc_steps=200, max_iter=10**6, seed=seed, n_runs=runs)
    run = eng.RunGroup(gp, opts)
    assert run.advance(0)
    out = []
    for r in range(runs):
        s = run.fetch(r)
        ev = eng.evidence_sampling(s["points"], s["logL"], s["pool"], n, 100, seed + r)
        out.append((ev["z"].mean() - truth) / ev["z"].std(ddof=1))
    run.close()
    return np.array(out)


@pytest.mark.parametrize("K", [1, 8, 32])
def test_c1_pull_distribution(eng, K):
    c = cfg.c1_gaussian()
    gp = eng.Problem.from_config(c)
    pulls = _pulls(eng, gp, c.truth["logZ"], 100, K, 24, 1000 + K)
    print(f"C1 K={K}: pull mean {pulls.mean():+.3f}, sd {pulls.std(ddof=1):.3f}, min {pulls.min():+.2f}, max {pulls.max():+.2f}")
    assert abs(pulls.mean()) < 0.5, pulls
    assert 0.6 < pulls.std(ddof=1) < 1.6, pulls
    assert np.abs(pulls).max() < 4.0


def test_c4_merged_runs_pull_distribution():
    """C4 as stated (64 runs x 512 live points, K = 64), 20 seeds, each merged with combineRuns: the merged evidence
    and ITS reported error (sigma ~ 0.016) against the quadrature value."""
    c = cfg.c4_gbm()
    obj = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), GeneratingDistribution=api.GeometricBrownianMotionProcess("mu", "sigma", 100.0),
        Parameters=[("mu", -1, 1), ("sigma", 0.01, 2)], PriorDistribution=["LocationParameter", "ScaleParameter"])
    pulls, sig = [], []
    for seed in range(20):
        res = api.parallelNestedSampling(obj, ParallelRuns=64, SamplePoolSize=512, BatchSize=64, MaxIterations=10**6,
                                         Seed=3000 + seed)
        z = res["LogEvidence"]
        pulls.append((z["Mean"] - c.truth["logZ"]) / z["StandardError"])
        sig.append(z["StandardError"])
    pulls = np.array(pulls)
    print(f"C4 64x512 K=64: pull mean {pulls.mean():+.3f}, sd {pulls.std(ddof=1):.3f}, sigma {np.mean(sig):.4f}, pulls {np.round(pulls, 2)}")
    assert abs(pulls.mean()) < 0.6, pulls
    assert 0.5 < pulls.std(ddof=1) < 1.8, pulls
