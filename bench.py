#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json: the live-point replacement step of nestedSampling.

A "step" is one nested-sampling iteration of one run group: the K worst live points are deleted, K walkers do
S = 200 constrained-prior Metropolis steps each (every proposal scores a log-likelihood that is a reduction
over all N data rows), the new points are inserted and the logX/logZ evidence state is advanced.
metric = log-likelihood evaluations per second (K*S evals per step); replacements/s = value / S.

N = 1 workload: config C2 of BASELINE.json (polynomial regression, 5 parameters, N = 1e6 rows, 1024 live points).
N > 1: one process per GPU (torchrun), run-sharded exactly like parallelNestedSampling (BS:1349-1357): every
rank advances its own independent run (run id = rank), no data-path collective; "scaling": "weak".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bayesianinference_b200 import configs as cfg  # noqa: E402

MC_STEPS = 200  # "MonteCarloSteps" default, BS:844
# profiles/r01g_ncu_walk_grid_c2.md: dram__bytes_read.sum + dram__bytes_write.sum of one walk_grid_kernel launch
NCU_TRAFFIC_C2_GRID = 16.165e6 + 0.088e6
WORKLOADS = {
    # name: (config factory, batch_k, flop per datum-eval, algorithmic bytes per datum-eval)   SURVEY §8d
    "C1": (cfg.c1_gaussian, 32, 3, 8),
    "C2": (cfg.c2_polyreg, 256, 9, 16),
    "C3": (cfg.c3_logistic, 512, 81, 36),
    "C4": (cfg.c4_gbm, 64, 4, 16),
}


def _workload_config(c, K, n_runs, world):
    N = c.inputs.shape[0]
    return {"workload": f"{c.name}: N={N} rows, d={c.d}, {c.pool_size} live points, {n_runs} run(s)/GPU x K={K} replaced per "
                        f"iteration, {MC_STEPS} walk steps each",
            "l2": "256 MiB written between timed iterations (inside the timed region); the data set (<= 48 MB) is "
                  "L2-resident across the 200 walk steps of an iteration by design",
            "parallelism": f"run-sharded x{world}, no collective"}


def _dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def _pinned(a):
    """Copy a numpy array into page-locked host memory (the e2e leg copies from pinned memory)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


def _cpu_leg(c, threads, reps_per_thread, seed=77):
    """The reference scheme (one walker, S sequential evaluations per replacement, BS:729) on host cores —
    oracle port (the reference is Wolfram Language; no Wolfram Engine offline)."""
    from oracle import oracle as O
    op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi, c.p0 or None, c.p1 or None)
    start = pr.sample(64, seed)
    t0 = time.perf_counter()
    evals = O.bench_walks(op, pr, start, O.LOGZERO, reps_per_thread, MC_STEPS, seed, threads)
    dt = time.perf_counter() - t0
    return evals, dt


def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return
    from oracle import oracle as O
    factory, K, flop, byts = WORKLOADS[args.config]
    c = factory()
    threads = len(os.sched_getaffinity(0))  # all host cores, even when torchrun pins OMP_NUM_THREADS=1
    for _ in range(max(args.warmup, 0) and 1):  # one short warm-up pass is enough for a CPU loop
        _cpu_leg(c, threads, 1)
    tot_e, tot_t = 0, 0.0
    reps = 4 if c.inputs.shape[0] >= 100_000 else 400  # ~1 s of CPU work per step
    for _ in range(args.steps):
        e, t = _cpu_leg(c, threads, reps)
        tot_e += e
        tot_t += t
    v = tot_e / tot_t
    sample = f"{threads} threads x {reps} replacements x {MC_STEPS} evals per step on the full {c.inputs.shape[0]}-row data"
    print(json.dumps({
        "impl": "reference", "metric": "loglikelihood evals/s", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _workload_config(c, K, args.runs_per_gpu, max(args.gpus, 1)),
        "replacements_per_s": v / MC_STEPS,
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of BS:859-1040 on host cores; Wolfram reference unavailable offline (BASELINE.md §3)",
    }))


def run_ours(args):
    import torch
    rank, local_rank, world = _dist_env()
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from bayesianinference_b200 import engine
    engine.init(device=local_rank)

    factory, K, flop, byts = WORKLOADS[args.config]
    c = factory()
    n, d, N = c.pool_size, c.d, c.inputs.shape[0]
    rows = N - 1 if c.op == cfg.OP_GBM else N
    n_runs = args.runs_per_gpu
    gp = engine.Problem.from_config(c)
    opts = engine.default_options(pool_size=n, batch_k=K, mc_steps=MC_STEPS, max_iter=10**9, min_iter=10**9,
                                  seed=2026, first_run_id=rank * n_runs, n_runs=n_runs)
    run = engine.RunGroup(gp, opts)  # starting points drawn from the prior on the device
    path = run.walk_path()
    # everything of the library runs on the problem's own stream: time with CUDA events recorded on THAT stream, and
    # put the L2 flush on it too so that it really sits between two iterations
    lib_stream = torch.cuda.ExternalStream(gp.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    torch.cuda.synchronize()

    def step():
        with torch.cuda.stream(lib_stream):
            flush.fill_(1)  # L2 flush between timed iterations
        run.advance(1)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if use_dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_before = run.timing()
    launches0 = engine.launch_count()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(lib_stream)
    for _ in range(args.steps):
        step()
    ev1.record(lib_stream)
    torch.cuda.synchronize()
    dt_wall = time.perf_counter() - t0
    dt = ev0.elapsed_time(ev1) * 1e-3  # device time on the launching stream (host gaps between iterations included)
    if use_dist:
        dist.barrier()
    launches = engine.launch_count() - launches0
    t_after = run.timing()
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt_max = float(tmax.item())
    walkers = n_runs * K
    evals_per_step = walkers * MC_STEPS
    value = world * evals_per_step * args.steps / dt_max

    # ---- e2e through the C ABI with HOST (pinned) buffers: define the problem (H2D of the data), create the run
    # from host start points, one replacement step, fetch the sample list back (D2H) — every step.
    e2e = None
    roof = None
    cpu = None
    if True:
        inp, _k1 = _pinned(c.inputs)
        out, _k2 = _pinned(c.outputs) if c.outputs is not None else (None, None)
        sp, _k3 = _pinned(gp.sample_prior(n * n_runs, seed=5, run_id=rank).reshape(n_runs, n, d))
        e_opts = engine.default_options(pool_size=n, batch_k=K, mc_steps=MC_STEPS, max_iter=10**9, min_iter=10**9,
                                        seed=11, first_run_id=rank * n_runs, n_runs=n_runs)

        def e2e_step():
            t = [time.perf_counter()]
            p2 = engine.Problem(c.op, inp, out, c.iparam, c.kinds, c.lo, c.hi, c.p0, c.p1)
            t.append(time.perf_counter())
            r2 = engine.RunGroup(p2, e_opts, sp)
            t.append(time.perf_counter())
            r2.advance(1)
            t.append(time.perf_counter())
            res = r2.fetch(0)
            t.append(time.perf_counter())
            nbytes = sum(v.nbytes for v in res.values() if isinstance(v, np.ndarray))
            r2.close()
            p2.close()
            t.append(time.perf_counter())
            if os.environ.get("BINEST_E2E_DEBUG"):
                print("e2e phases ms:", [round(1e3 * (b - a), 2) for a, b in zip(t, t[1:])], file=sys.stderr)
            return nbytes

        for _ in range(min(args.warmup, 3)):
            d2h = e2e_step()
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            d2h = e2e_step()
        torch.cuda.synchronize()
        edt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if use_dist:
            dist.all_reduce(edt, op=dist.ReduceOp.MAX)
        h2d = inp.nbytes + (out.nbytes if out is not None else 0) + sp.nbytes
        e2e = {"value": world * evals_per_step * e_steps / float(edt.item()), "unit": "evals/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e_steps,
               "what": "binest_problem_create (data H2D from pinned host) + binest_run_create (host start points) + "
                       "binest_run_advance(1) + binest_run_fetch (D2H), per step"}

    if rank == 0:
        # ---- roofline of the dominant kernel (loglike_stream_kernel): algorithmic flop per launch / launch duration.
        # Duration from CUDA events recorded on the library's stream around every walk graph of the timed region
        # (S launches of the kernel, interleaved with the small walk_step kernel — so it is an upper bound).
        graphs = t_after["walk_graphs"] - t_before["walk_graphs"]
        walk_ms = t_after["walk_ms"] - t_before["walk_ms"]
        peak_tf = engine.fp64_peak()
        iso_kernel_ms, iso_total_ms = gp.bench_loglike(walkers, 20, 3, True)
        # dominant kernel.  grid-resident path: ONE launch of walk_grid_kernel scores walkers x S proposals (the whole
        # walk); stepped path: one launch of loglike_stream_kernel scores `walkers` proposals (one walk step).
        steps_per_launch = MC_STEPS if path in ("grid-resident", "cluster-resident") else 1
        kernel = {"grid-resident": "walk_grid_kernel", "cluster-resident": "walk_resident_kernel"}.get(path, "loglike_stream_kernel")
        ms_launch = walk_ms / max(graphs, 1) / MC_STEPS * steps_per_launch
        alg_flop = float(flop) * rows * walkers * steps_per_launch
        alg_bytes = float(byts) * rows  # every data row is read from HBM/L2 once per launch and shared by all walkers
        achieved_tf = alg_flop / (ms_launch * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum of that kernel, one `ncu --set full` capture (profiles/)
        traffic = {("C2", "walk_grid_kernel"): NCU_TRAFFIC_C2_GRID}.get((args.config, kernel))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        roof = {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic,
                "kernel": kernel, "walk_path": path, "walk_steps_per_launch": steps_per_launch,
                "ms_per_launch": ms_launch,
                "stream_kernel_ms_isolated": iso_kernel_ms,
                "stream_kernel_frac_isolated": float(flop) * rows * walkers / (iso_kernel_ms * 1e-3) / 1e12 / peak_tf,
                "alg_flop_per_launch": alg_flop, "alg_bytes_per_launch": alg_bytes,
                "hbm_time_bound_ms": alg_bytes / (hbm_peak * 1e9) * 1e3,
                "peak_source": "fp64 DFMA peak measured live by binest_measure_fp64_peak (not in MEASURED_PEAKS.json); "
                               f"hbm {hbm_peak} GB/s " + ("of measured" if peaks else "of fallback"),
                "why": "P walkers share every data tile, intensity = flop*P/bytes >> fp64 ridge (~6 flop/B): fp64-FMA bound"}
        # ---- CPU baseline beside it: oracle port, bounded sample
        from oracle import oracle as O
        threads = len(os.sched_getaffinity(0))  # all host cores, even when torchrun pins OMP_NUM_THREADS=1
        e1, t1 = _cpu_leg(c, threads, 1 if N >= 100_000 else 200)  # calibration pass, then ~12 s of CPU work
        e, t = _cpu_leg(c, threads, max(1, int(round(12.0 / max(t1, 1e-3)))) * (1 if N >= 100_000 else 200))
        cpu = {"value": e / t, "unit": "evals/s", "cores": threads, "kind": "port",
               "sample": f"{e} evals ({threads} threads x {e // max(threads, 1) // MC_STEPS} replacements x {MC_STEPS} steps) "
                         f"on the full {N}-row data, {t:.1f} s"}
        line = {
            "metric": "loglikelihood evals/s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _workload_config(c, K, n_runs, world),
            "replacements_per_s": value / MC_STEPS,
            "device_walk_ms_per_step": walk_ms / max(graphs, 1),
            "timing": "CUDA events on the library's stream around the K timed steps, max over ranks",
            "wall_ms_per_step": 1e3 * dt_wall / args.steps,
            "gpu_launches": int(launches),
            "clocks": clocks, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=list(WORKLOADS))
    ap.add_argument("--runs-per-gpu", type=int, default=1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
