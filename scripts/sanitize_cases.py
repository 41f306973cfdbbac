"""Tiny invocations of every hot kernel, for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python scripts/sanitize_cases.py [case ...]
Cases: loop (ns_loop_kernel), resident (walk_resident_kernel: DSMEM exchange, clusters), grid (walk_grid_kernel: grid
barriers), stepped (graph of walk_step + loglike_stream with PDL), gp (fill / potf2 / trsm / syrk / finish, predict),
evidence (crude weights, evidence sampling), mcmc (posterior sampler), merge (combineRuns on the device: host-array and
run-group variants).  Sizes are small: the tools slow kernels 10-100x."""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from bayesianinference_b200 import engine, configs as cfg

cases = sys.argv[1:] or ["loop", "resident", "grid", "stepped", "gp", "evidence", "mcmc", "merge"]
engine.init()


def run(c, K, n, S, iters, runs=1, env=None):
    for k in ("BINEST_NO_LOOP", "BINEST_NO_RESIDENT", "BINEST_NO_GRID", "BINEST_NO_PDL"):
        os.environ.pop(k, None)
    os.environ.update(env or {})
    gp = engine.Problem.from_config(c)
    o = engine.default_options(pool_size=n, batch_k=K, mc_steps=S, max_iter=iters, min_iter=iters, seed=3, n_runs=runs)
    r = engine.RunGroup(gp, o)
    path = r.walk_path()
    r.advance(0)
    s = r.fetch(0)
    print(f"  {c.name}: path {path}, M = {s['M']}, crude logZ = {s['crude_logZ']:.6f}")
    r.close()
    gp.close()
    return path


for case in cases:
    print("case", case)
    if case == "loop":
        assert run(cfg.c1_gaussian(), 4, 32, 12, 16, runs=2) == "device-loop"
        assert run(cfg.c2_polyreg(N=300), 1, 32, 12, 8) == "device-loop"
    elif case == "resident":
        assert run(cfg.c4_gbm(T=3000), 24, 64, 6, 48, runs=2, env={"BINEST_NO_LOOP": "1"}) == "cluster-resident"
        assert run(cfg.c4_gbm(T=16384), 64, 128, 4, 64, runs=4) == "cluster-resident"  # data over a cluster: DSMEM
        assert run(cfg.c3_logistic(N=2000), 16, 64, 4, 32, env={"BINEST_NO_LOOP": "1"}) == "cluster-resident"
    elif case == "grid":
        assert run(cfg.c2_polyreg(N=20000), 96, 256, 4, 192, env={"BINEST_NO_RESIDENT": "1"}) == "grid-resident"
    elif case == "stepped":
        e = {"BINEST_NO_RESIDENT": "1", "BINEST_NO_GRID": "1"}
        assert run(cfg.c2_polyreg(N=20000), 96, 256, 4, 192, env=e) == "stepped-graph"
        assert run(cfg.c3_logistic(N=5000), 32, 64, 3, 64, env=e) == "stepped-graph"
    elif case == "gp":
        c = cfg.c5_gp(N=300)
        gp = engine.Problem.from_config(c)
        th = gp.sample_prior(9, seed=5)
        print("  gp loglike", gp.loglike(th)[:3])
        m, s = gp.gp_predict(th[:3], np.linspace(0, 10, 130))
        print("  gp predict", m[0, :2], s[0, :2])
        run(c, 6, 24, 3, 12)
    elif case == "evidence":
        rng = np.random.default_rng(1)
        M, n = 700, 64
        logL = np.sort(rng.normal(-100, 5, M)); pts = rng.normal(size=(M, 3))
        pool = np.concatenate([np.full(M - n, n), np.arange(n, 0, -1)]).astype(np.int64)
        print("  crude", engine.crude_weights(logL, pool, n)["crude_logZ"])
        print("  evidence", engine.evidence_sampling(pts, logL, pool, n, 20, 3)["z"][:3])
    elif case == "mcmc":
        c = cfg.c2_polyreg(N=2000)
        gp = engine.Problem.from_config(c)
        ch = engine.Chain(gp, [[0.5, -1.2, 0.8, 0.3, 0.25]] * 4, np.diag([1e-4] * 5), learn_delay=5, seed=2)
        print("  mcmc", ch.iterate(20)[-1, 0])
    elif case == "merge":
        c = cfg.c4_gbm(T=512)
        gp = engine.Problem.from_config(c)
        o = engine.default_options(pool_size=48, batch_k=6, mc_steps=20, seed=5, n_runs=5, first_run_id=3)
        grp = engine.RunGroup(gp, o)
        grp.advance(0)
        tabs = []
        for i in range(5):
            f = grp.fetch(i, weights=False)
            tabs.append({"Point": f["points"], "LogLikelihood": f["logL"], "LogPriorPDF": f["logPrior"], "AcceptanceRate": f["acc"],
                         "PoolSize": f["pool"]})
        m, live = grp.merge()
        print("  run merge", m["LogLikelihood"].size, live)
        res = grp.combine(False, 20, 3)
        print("  run combine", res["z"][:2], res["n_live"])
        m2, _ = engine.merge_runs(tabs)
        res2 = engine.combine_runs(tabs, True, 5 * 48, 20, 3)
        print("  host-array merge / combine", m2["LogLikelihood"].size, res2["z"][:2])
print("done")
