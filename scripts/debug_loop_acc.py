"""Debug: device-loop path vs oracle vs resident path on C1, K = 1: where do the acceptance rates differ?"""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from bayesianinference_b200 import engine, configs as cfg
from oracle import oracle as O
engine.init()
c = cfg.c1_gaussian()
gp = engine.Problem.from_config(c)
op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
pr = O.Prior(c.kinds, c.lo, c.hi)
for iters, pieces in ((600, False), (600, True), (4800, False)):
    n, S, K = 100, 40, 1
    start = pr.sample(n, 21, 0)[None]
    opts = engine.default_options(pool_size=n, batch_k=K, mc_steps=S, max_iter=iters, min_iter=iters, seed=21, n_runs=1)
    ref = O.nested_sampling(op, pr, pool_size=n, batch_k=K, mc_steps=S, max_iter=iters, min_iter=iters, seed=21,
                            adapt_in_walk=False, start_points=start[0], run_id=0)
    for mode in ("loop", "resident"):
        if mode == "resident":
            os.environ["BINEST_NO_LOOP"] = "1"
        else:
            os.environ.pop("BINEST_NO_LOOP", None)
        run = engine.RunGroup(gp, opts, start)
        if pieces:
            run.advance(3); run.fetch(0)
        run.advance(0)
        g = run.fetch(0)
        m = ~np.isnan(ref.acc)
        bad = np.flatnonzero(np.abs(g["acc"][m] - ref.acc[m]) > 1e-12)
        pbad = np.flatnonzero(np.abs(g["points"] - ref.points).max(1) > 1e-7 * np.abs(ref.points).max(1))
        print(f"iters {iters} pieces {pieces} {mode:9s} path {run.walk_path()}: M {g['M']} vs {ref.logL.size}; acc mismatches {bad.size} first {bad[:5]} "
              f"got {g['acc'][m][bad[:5]]} ref {ref.acc[m][bad[:5]]}; point mismatches {pbad.size} first {pbad[:3]}")
        run.close()
