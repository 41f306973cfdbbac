#!/usr/bin/env python
"""Complete nested-sampling runs of the BASELINE.json configs at their stated sizes, through the reference-facing API:
LogEvidence against the closed-form / quadrature / Laplace pins, and whole-run replacement throughput (wall clock of
the API call, so host merge and post-processing are included).

  python scripts/full_runs.py [C1 C2 C3 C4 C5small ...] [--out gpurun_out/full_runs.md]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from bayesianinference_b200 import api, engine  # noqa: E402
from bayesianinference_b200 import configs as cfg  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
PINS = json.load(open(os.path.join(G, "laplace_pins.json")))
GOLD = json.load(open(os.path.join(G, "pins.json")))


def _params(c):
    return [(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)]


def c1():
    c = cfg.c1_gaussian()
    obj = api.defineInferenceProblem(Data=c.inputs[:, 0], GeneratingDistribution=api.NormalDistribution("mu", "sigma"),
                                     Parameters=_params(c), PriorDistribution=["LocationParameter", "ScaleParameter"])
    return obj, dict(SamplePoolSize=100, BatchSize=1), None, GOLD["c1_logZ_quadrature"], "2-D quadrature"


def c2():
    c = cfg.c2_polyreg()
    obj = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), IndependentVariables=["x"],
        GeneratingDistribution=api.NormalDistribution(api.Polynomial("x", tuple(c.names[:4])), "sigma"),
        Parameters=_params(c), PriorDistribution=["LocationParameter"] * 4 + ["ScaleParameter"])
    return obj, dict(SamplePoolSize=1024, BatchSize=256), None, c.truth["logZ"], "exact (analytic + 1-D quadrature)"


def c3():
    c = cfg.c3_logistic()
    obj = api.defineInferenceProblem(
        Data=(c.inputs, c.outputs[:, 0]), GeneratingDistribution=api.CategoricalSoftmax(tuple(c.names), 3),
        Parameters=_params(c), PriorDistribution=[api.NormalDistribution(0.0, 5.0)] * c.d)
    return obj, dict(SamplePoolSize=2048, BatchSize=512), None, PINS["C3"]["logZ_laplace"], "Laplace, N = 1e6"


def c4():
    c = cfg.c4_gbm()
    obj = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), GeneratingDistribution=api.GeometricBrownianMotionProcess("mu", "sigma", 100.0),
        Parameters=_params(c), PriorDistribution=["LocationParameter", "ScaleParameter"])
    return obj, dict(SamplePoolSize=512, BatchSize=64), 64, GOLD["c4_logZ_quadrature"], "2-D quadrature"


def c5small():
    p = PINS["C5_small"]
    c = cfg.c5_gp(N=p["N"])
    obj = api.defineGaussianProcess((c.inputs[:, 0], c.outputs[:, 0]), api.SquaredExponentialGP(*c.names), _params(c),
                                    ["ScaleParameter"] * 3)
    return obj, dict(SamplePoolSize=256, BatchSize=64), None, p["logZ_quadrature"], f"3-D quadrature, N = {p['N']}"


CASES = {"C1": c1, "C2": c2, "C3": c3, "C4": c4, "C5small": c5small}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=list(CASES))
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "full_runs.md"))
    ap.add_argument("--seed", type=int, default=2026)
    a = ap.parse_args()
    engine.init(device=0)
    rows = ["| config | options | samples M | replacements | wall s | replacements/s | evals/s | LogEvidence (mean ± sd) | truth | pull | H (nats) |",
            "|---|---|---|---|---|---|---|---|---|---|---|"]
    for name in a.cases:
        obj, opts, runs, truth, how = CASES[name]()
        t0 = time.perf_counter()
        if runs:
            res = api.parallelNestedSampling(obj, ParallelRuns=runs, MaxIterations=10**6, Seed=a.seed, **opts)
        else:
            res = api.nestedSampling(obj, MaxIterations=10**6, Seed=a.seed, **opts)
        dt = time.perf_counter() - t0
        z = res["LogEvidence"]
        gen = int(res["GeneratedNestedSamples"])
        pull = (z["Mean"] - truth) / z["StandardError"]
        o = ", ".join(f"{k}={v}" for k, v in opts.items()) + (f", ParallelRuns={runs}" if runs else "")
        rows.append(f"| {name} | {o} | {res['TotalSamples']} | {gen} | {dt:.2f} | {gen / dt:.0f} | {200 * gen / dt:.3g} | "
                    f"{z['Mean']:.4f} ± {z['StandardError']:.4f} | {truth:.4f} ({how}) | {pull:+.2f} | "
                    f"{res['RelativeEntropy']['Mean']:.2f} |")
        print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        fh.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
