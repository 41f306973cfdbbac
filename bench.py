#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json: the live-point replacement step of nestedSampling.

Primary line (what the driver reads).  A "step" is one nested-sampling iteration of one run group: the K worst live
points are deleted, K walkers do S = 200 constrained-prior Metropolis steps each (every proposal scores a
log-likelihood that is a reduction over all N data rows), the new points are inserted and the logX/logZ evidence state
is advanced.  metric = log-likelihood evaluations per second (K*S evals per step); replacements/s = value / S.
  N = 1 workload: config C2 of BASELINE.json (polynomial regression, 5 parameters, N = 1e6 rows, 1024 live points).
  N > 1: one process per GPU (torchrun), run-sharded exactly like parallelNestedSampling (BS:1349-1357): every rank
  advances its own independent run (run id = rank), no data-path collective; "scaling": "weak".

Secondary objects in the same JSON line ("extras", each measured at the same N GPUs, each a STRONG-scaling point —
total work fixed, so the driver's 1/2/4/8 sweep yields their curves; skipped with --no-extras):
  "c1"            config C1 as the reference runs it: 100 live points, K = 1, 200 steps, to termination (latency-bound;
                  the whole loop on the device, walk_loop.cuh).  Unit: live-point replacements/s.
  "c5"            config C5: GP marginal likelihood, N = 4096, ONE batched evaluation of 256 hyper-parameter sets
                  (fill + blocked Cholesky on FP64 DMMA); N > 1: batch-sharded, 256/N matrices per GPU, the finished
                  values exchanged in-kernel over peer-mapped memory.  Roofline: fp64 tensor pipe.
  "c4_strong"     config C4 as BASELINE states it: 64 parallelNestedSampling runs x 512 live points through
                  api.parallelNestedSampling, the WHOLE call timed (device loop + merge of the runs + evidenceSampling,
                  both on the device); N > 1: 64/N runs per GPU.  Unit: live-point replacements/s.
  "c4_strong_T262144"  the same job on 16x the data (262 144 increments): a walk step is fp64 work instead of latency.
  "data_sharded"  the data-sharded mode: C2-shaped data with --rows rows (default 6.4e7) split N ways, K = 256
                  walkers x 200 steps per iteration, per-step exchange of 8 P bytes to every peer.
A watchdog bounds the extras: if they do not finish in time the line is printed without the missing ones.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C1..C5]
                  [--mode run-sharded|data-sharded] [--rows R] [--no-extras]
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bayesianinference_b200 import configs as cfg  # noqa: E402

MC_STEPS = 200  # "MonteCarloSteps" default, BS:844
# profiles/r02_ncu_walk_grid_c2.md: dram__bytes_read.sum + dram__bytes_write.sum of one walk_grid_kernel launch
NCU_TRAFFIC_C2_GRID = 16.169e6 + 0.076e6
GP_FLOP_PER_THETA = 2.31e10   # fill 2.1e8 + Cholesky N^3/3 = 2.29e10 + solve/logdet, N = 4096
GP_BYTES_PER_THETA = 134.2e6  # lower triangle written once + read once (minimum traffic)
WORKLOADS = {
    # name: (config factory, batch_k, flop per datum-eval, algorithmic bytes per datum-eval)   SURVEY §8d
    "C1": (cfg.c1_gaussian, 32, 3, 8),
    "C2": (cfg.c2_polyreg, 256, 9, 16),
    "C3": (cfg.c3_logistic, 512, 81, 36),
    "C4": (cfg.c4_gbm, 64, 4, 16),
    "C5": (cfg.c5_gp, 256, None, None),
}
EXTRAS_DEADLINE_S = 420.0


def _workload_config(c, K, n_runs, world):
    N = c.inputs.shape[0]
    return {"workload": f"{c.name}: N={N} rows, d={c.d}, {c.pool_size} live points, {n_runs} run(s)/GPU x K={K} replaced per "
                        f"iteration, {MC_STEPS} walk steps each",
            "l2": "256 MiB written between timed iterations (inside the timed region); the data set (<= 48 MB) is "
                  "L2-resident across the 200 walk steps of an iteration by design",
            "parallelism": f"run-sharded x{world}, no collective"}


def _dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _wolfram_probe():
    """the reference's own Wolfram path can only be timed where a Wolfram Engine exists."""
    path = shutil.which("wolframscript") or shutil.which("WolframKernel") or shutil.which("wolfram")
    return {"wolframscript": path,
            "status": "unavailable: no Wolfram Engine on this machine (BASELINE.md §3)" if path is None else
                      "found, but the reference paclet is not on this machine (bench.py may not read /root/reference at run time)"}


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


# ------------------------------------------------------------------------------------------------- CPU arm
def _cpu_leg(c, threads, reps_per_thread, seed=77):
    """The reference scheme (one walker per chain, S sequential density evaluations per replacement, BS:729; a proposal
    outside the box is rejected without a likelihood evaluation, BS:602-617) on host cores — the oracle port built
    -O3 -march=native with vectorised sums (oracle/Makefile FAST_*).  The reference itself is Wolfram Language; no
    Wolfram Engine offline.  Returns (evaluations performed, replacements, seconds)."""
    from oracle import oracle as O
    O.fast_lib()  # build / load outside the timed region
    op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi, c.p0 or None, c.p1 or None)
    start = pr.sample(64, seed)
    t0 = time.perf_counter()
    evals = O.bench_walks(op, pr, start, O.LOGZERO, reps_per_thread, MC_STEPS, seed, threads, fast=True)
    dt = time.perf_counter() - t0
    return evals, reps_per_thread * threads, dt


def _cpu_gp_leg(c, threads, n_theta):
    """C5 on host cores: the reference's LU path (GP:130-141) restated in the oracle, one theta per thread."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    op = O.Problem(c.op, c.d, c.inputs, c.outputs, c.iparam)
    pr = O.Prior(c.kinds, c.lo, c.hi)
    th = pr.sample(n_theta, 905)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL
        list(ex.map(lambda t: op.loglike(t[None, :], pr), th))
    return n_theta, time.perf_counter() - t0


def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return
    from oracle import oracle as O
    factory, K, flop, byts = WORKLOADS[args.config]
    c = factory()
    threads = len(os.sched_getaffinity(0))  # all host cores, even when torchrun pins OMP_NUM_THREADS=1
    if args.config == "C5":
        n, t = _cpu_gp_leg(c, threads, max(threads, 8) if args.steps > 1 else threads)
        v = n / t
        line = {"impl": "reference", "metric": "loglikelihood evals/s", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C5-gp: N={c.inputs.shape[0]}, batch of 256 hyper-parameter sets"},
                "cpu_baseline": {"value": v, "unit": "evals/s", "cores": threads, "kind": "port",
                                 "sample": f"{n} covariance matrices of order {c.inputs.shape[0]} (fill + LU + solve), one per thread"},
                "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wolfram": _wolfram_probe()}
        print(json.dumps(line))
        return
    for _ in range(max(args.warmup, 0) and 1):  # one short warm-up pass is enough for a CPU loop
        _cpu_leg(c, threads, 1)
    tot_e, tot_r, tot_t = 0, 0, 0.0
    reps = 4 if c.inputs.shape[0] >= 100_000 else 400  # ~1 s of CPU work per step
    for _ in range(args.steps):
        e, r, t = _cpu_leg(c, threads, reps)
        tot_e += e
        tot_r += r
        tot_t += t
    v = tot_e / tot_t
    sample = (f"{threads} threads x {reps} replacements x {MC_STEPS} walk steps per step on the full {c.inputs.shape[0]}-row data; "
              f"{tot_e} likelihood evaluations performed (proposals outside the box are rejected without one, BS:602-617); {O.FAST_FLAGS}")
    print(json.dumps({
        "impl": "reference", "metric": "loglikelihood evals/s", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _workload_config(c, K, args.runs_per_gpu, max(args.gpus, 1)),
        "replacements_per_s": tot_r / tot_t,
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wolfram": _wolfram_probe(),
        "note": "oracle port of BS:859-1040 on host cores; Wolfram reference unavailable offline (BASELINE.md §3)",
    }))


# ------------------------------------------------------------------------------------------------- helpers (GPU)
class Ctx:
    """torch / torch.distributed / engine state of one rank."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank, self.local_rank, self.world = _dist_env()
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        from bayesianinference_b200 import engine
        engine.init(device=self.local_rank)
        self.engine = engine
        self._comm = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def comm(self):
        if self._comm is None and self.world > 1:
            self._comm = self.engine.Comm(self.rank, self.world)
        return self._comm


def measure_c5(ctx, steps, warmup, batch=256):
    """One batched GP evaluation per step: 256 theta-sets at N = 4096 (BASELINE config C5).  Device-timed value with the
    data and theta resident (binest_bench_loglike: CUDA events on the library's stream), e2e through binest_loglike
    with host theta in / host logL out.  N > 1: batch-sharded (strong scaling: the 256 sets are split)."""
    eng = ctx.engine
    c = cfg.c5_gp()
    N = c.inputs.shape[0]
    comm = ctx.comm()
    gp = eng.Problem.from_config(c, comm=comm, shard="batch") if comm is not None else eng.Problem.from_config(c)
    th = gp.sample_prior(batch, seed=905)
    gp.loglike(th[: max(2 * ctx.world, 8)])  # first touch (kernel attributes, streams)
    ctx.barrier()
    ms_k, ms_tot = gp.bench_loglike(batch, reps=max(steps, 1), warmup=max(min(warmup, 2), 1), flush_l2=False)
    ms = ctx.max_over_ranks(ms_tot)
    ctx.barrier()
    t0 = time.perf_counter()
    e_steps = max(2, min(steps, 3))
    for _ in range(e_steps):
        out = gp.loglike(th)
    ctx.torch.cuda.synchronize()
    e_dt = ctx.max_over_ranks((time.perf_counter() - t0) / e_steps)
    stats = comm.stats() if comm is not None else None
    peak = eng.fp64_peak()
    per_rank = -(-batch // ctx.world)
    ach = GP_FLOP_PER_THETA * per_rank / (ms * 1e-3) / 1e12
    res = {
        "metric": "loglikelihood evals/s", "unit": "evals/s", "value": batch / (ms * 1e-3), "ms_per_step": ms,
        "scaling": "strong", "n_gpus": ctx.world,
        "config": {"workload": f"C5-gp: N={N} points, squared-exponential kernel + nugget, one batched evaluation of {batch} "
                               f"hyper-parameter sets ({per_rank} matrices of {N}^2 fp64 per GPU, {per_rank * 134.2e6 / 1e9:.1f} GB)",
                   "parallelism": "single GPU" if ctx.world == 1 else f"batch-sharded x{ctx.world}: theta batch split, values "
                                  "exchanged in-kernel over peer-mapped memory", "l2": "inputs larger than L2 (134 MB per matrix)"},
        "e2e": {"value": batch / e_dt, "unit": "evals/s", "h2d_bytes_per_step": int(th.nbytes), "d2h_bytes_per_step": int(out.nbytes),
                "what": "binest_loglike: host theta in, fill + factor + solve, host logL out"},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                     "kernel": "gp_syrk_kernel (DMMA.8x8x4) inside the fill + blocked-Cholesky pipeline",
                     "alg_flop_per_launch": GP_FLOP_PER_THETA * per_rank, "alg_bytes_per_launch": GP_BYTES_PER_THETA * per_rank,
                     "peak_source": "fp64 peak measured live (register-resident DFMA loop; DMMA shares the pipe); achieved = "
                                    "2.31e10 flop x matrices per GPU / pipeline time (all kernels of the evaluation)"},
        "exchange": stats, "finite": int(np.isfinite(out).sum()),
    }
    gp.close()
    return res


def measure_c1(ctx, reps=5):
    """Config C1 exactly as the reference runs it (BS:837-851 defaults): 100 live points, ONE replacement per iteration,
    200 walk steps, to termination — a latency-bound chain of ~850 dependent iterations.  The whole loop stays on the
    device (walk_loop.cuh); wall clock of binest_run_advance(0).  Every rank runs its own seeds (replicas)."""
    eng = ctx.engine
    c = cfg.c1_gaussian()
    gp = eng.Problem.from_config(c)
    best = None
    for rep in range(reps + 1):
        o = eng.default_options(pool_size=100, batch_k=1, mc_steps=MC_STEPS, max_iter=10**6, seed=500 + rep + 100 * ctx.rank)
        run = eng.RunGroup(gp, o)
        path = run.walk_path()
        t0 = time.perf_counter()
        run.advance(0)
        dt = time.perf_counter() - t0
        sz = run.sizes(0)
        s = run.fetch(0)
        run.close()
        if rep > 0 and (best is None or sz["n_deleted"] / dt > best[0]):
            best = (sz["n_deleted"] / dt, dt, sz["n_deleted"], s["crude_logZ"])
    gp.close()
    return {"metric": "live-point replacements/s", "unit": "replacements/s", "value": best[0], "s_per_run": best[1],
            "replacements": int(best[2]), "crude_logZ": best[3], "logZ_quadrature": c.truth["logZ"], "walk_path": path,
            "scaling": "replicas", "n_gpus": ctx.world,
            "config": {"workload": "C1-gaussian: N=100 rows, 100 live points, K=1 (the reference scheme), 200 walk steps, "
                                   "run to termination", "parallelism": "one sequential chain per GPU"}}


def measure_c4_strong(ctx, reps=2, T=16384):
    """Config C4 as BASELINE states it, through the reference-facing call: parallelNestedSampling with 64 runs x 512 live
    points (BS:1317-1371), whole call timed on every rank (max over ranks): device loops of this rank's runs, the merge of
    the runs (combineRuns BS:1293-1315) and evidenceSampling (BS:1158-1291) — on the device (csrc/merge.cu): one GPU merges
    its run group in place; several GPUs merge their own runs, gather the per-GPU merges and merge those.
    T = 262144 (BASELINE.md §4 lists it beside 16384) is the same job on 16x the data: there a walk step is fp64 work, not
    latency, and the job scales."""
    from bayesianinference_b200 import api
    c = cfg.c4_gbm(T=T)
    obj = api.defineInferenceProblem(
        Data=(c.inputs[:, 0], c.outputs[:, 0]), GeneratingDistribution=api.GeometricBrownianMotionProcess("mu", "sigma", 100.0),
        Parameters=[(nm, lo, hi) for nm, lo, hi in zip(c.names, c.lo, c.hi)],
        PriorDistribution=["LocationParameter", "ScaleParameter"])
    best, info, allreps = None, None, []
    for rep in range(reps + 1):  # first pass warms the library up (kernel attributes, pools)
        ctx.barrier()
        t0 = time.perf_counter()
        res = api.parallelNestedSampling(obj, ParallelRuns=64, SamplePoolSize=512, BatchSize=64, MaxIterations=10**6,
                                         Seed=2026, PostProcessSamplingRuns=100)  # same job every repetition and GPU count
        ctx.torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        allreps.append({"s": dt, "phases": res.get("_Timing")})
        if rep > 0 and (best is None or dt < best):
            best = dt
            info = (res["GeneratedNestedSamples"], res["LogEvidence"], res.get("_Timing"))
    gen, logz, timing = info
    truth = c.truth.get("logZ")
    pull = (logz["Mean"] - truth) / max(logz["StandardError"], 1e-12) if truth is not None else None
    iters = gen / (64 * 64)
    walk_flop = 4.0 * T * 64 * 64 * MC_STEPS * iters  # 4 flop per increment and proposal
    return {"metric": "live-point replacements/s", "unit": "replacements/s", "value": gen / best, "s_per_call": best,
            "scaling": "strong", "n_gpus": ctx.world, "replacements": int(gen),
            "config": {"workload": f"C4-gbm: 64 parallelNestedSampling runs x 512 live points, T={T} increments, K=64 "
                                   "replaced per iteration per run, 200 walk steps", "parallelism": f"run-sharded x{ctx.world}: "
                                   f"{-(-64 // ctx.world)} runs per GPU, runs merged on the device (combineRuns + "
                                   "evidenceSampling in csrc/merge.cu)"},
            "phases_s_rank0": timing, "all_calls": allreps, "log_evidence": logz, "pull_vs_quadrature": pull,
            "device_loop_tflops_aggregate": walk_flop / max(timing.get("device_loop_s", float("nan")), 1e-9) / 1e12 if timing else None}


def measure_data_sharded(ctx, rows, iters=2, K=256):
    """Data-sharded mode: C2-shaped data, `rows` rows split over the ranks; one iteration = K walkers x
    200 steps, every step one exchange of the K per-rank sums (8 K bytes to each peer)."""
    eng = ctx.engine
    c = cfg.c2_polyreg(N=int(rows))
    comm = ctx.comm()
    p = eng.Problem.from_config(c, comm=comm) if comm is not None else eng.Problem.from_config(c)
    o = eng.default_options(pool_size=1024, batch_k=K, mc_steps=MC_STEPS, max_iter=10**9, min_iter=10**9, seed=3)
    run = eng.RunGroup(p, o)
    path = run.walk_path()
    run.advance(1)
    ctx.barrier()
    s0 = comm.stats() if comm is not None else None
    t0 = time.perf_counter()
    run.advance(iters)
    ctx.torch.cuda.synchronize()
    dt = ctx.max_over_ranks((time.perf_counter() - t0) / iters)
    s1 = comm.stats() if comm is not None else None
    tm = run.timing()
    peak = eng.fp64_peak()
    evals = K * MC_STEPS / dt
    flops = evals * 9.0 * rows
    ex = None
    if s1 is not None:
        n_ex = (s1["exchanges"] - s0["exchanges"]) / iters
        ex = {"peer_path": s1["peer_path"], "exchanges_per_iteration": n_ex,
              "nvlink_payload_bytes_per_step_per_rank": (s1["bytes_pushed"] - s0["bytes_pushed"]) / max(s1["exchanges"] - s0["exchanges"], 1),
              "expected_8P_times_peers": 8 * K * (ctx.world - 1)}
    run.close()
    p.close()
    return {"metric": "loglikelihood evals/s", "unit": "evals/s", "value": evals, "ms_per_iteration": 1e3 * dt,
            "scaling": "strong", "n_gpus": ctx.world, "walk_path": path, "device_walk_ms_per_iteration": tm["walk_ms"] / max(tm["walk_graphs"], 1),
            "config": {"workload": f"C2-shaped polynomial regression, {int(rows)} rows ({rows * 16 / 1e9:.2f} GB) split over {ctx.world} "
                                   f"GPU(s), K={K} walkers x {MC_STEPS} steps per iteration", "l2": "inputs larger than L2"},
            "aggregate_tflops": flops / 1e12, "frac_of_fp64_peak_aggregate": flops / 1e12 / (peak * ctx.world), "exchange": ex}


# ------------------------------------------------------------------------------------------------- primary (C1..C4)
def run_primary(ctx, args):
    torch, engine, dist = ctx.torch, ctx.engine, ctx.dist
    rank, local_rank, world = ctx.rank, ctx.local_rank, ctx.world
    use_dist = world > 1
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
    ctx.barrier()
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
    dt_max = ctx.max_over_ranks(dt)
    walkers = n_runs * K
    evals_per_step = walkers * MC_STEPS
    value = world * evals_per_step * args.steps / dt_max

    # ---- e2e through the C ABI with HOST (pinned) buffers: define the problem (H2D of the data), create the run
    # from host start points, one replacement step, fetch the sample list back (D2H) — every step.
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
    ctx.barrier()
    e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        d2h = e2e_step()
    torch.cuda.synchronize()
    edt = ctx.max_over_ranks(time.perf_counter() - t0)
    h2d = inp.nbytes + (out.nbytes if out is not None else 0) + sp.nbytes
    e2e = {"value": world * evals_per_step * e_steps / edt, "unit": "evals/s",
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e_steps,
           "what": "binest_problem_create (data H2D from pinned host) + binest_run_create (host start points) + "
                   "binest_run_advance(1) + binest_run_fetch (D2H), per step"}
    if rank != 0:
        return None
    # ---- roofline of the dominant kernel: algorithmic flop per launch / launch duration.  Duration from CUDA events
    # recorded on the library's stream around every walk launch of the timed region.
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
    peaks = _peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roof = {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf, "traffic": traffic,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture "
                              "profiles/r02_ncu_walk_grid_c2.md (not re-measured in this run)" if traffic else None,
            "kernel": kernel, "walk_path": path, "walk_steps_per_launch": steps_per_launch,
            "ms_per_launch": ms_launch,
            "pipe_slots_note": "9 algorithmic flop per datum (SURVEY §8d) execute as 4 DFMA = 8 flop: hardware busy fraction = frac * 8/9",
            "frac_executed_flop": achieved_tf / peak_tf * 8.0 / 9.0 if args.config == "C2" else None,
            "stream_kernel_ms_isolated": iso_kernel_ms,
            "stream_kernel_frac_isolated": float(flop) * rows * walkers / (iso_kernel_ms * 1e-3) / 1e12 / peak_tf,
            "alg_flop_per_launch": alg_flop, "alg_bytes_per_launch": alg_bytes,
            "hbm_time_bound_ms": alg_bytes / (hbm_peak * 1e9) * 1e3,
            "peak_source": "fp64 DFMA peak measured live by binest_measure_fp64_peak (not in MEASURED_PEAKS.json); "
                           f"hbm {hbm_peak} GB/s " + ("of measured" if peaks else "of fallback"),
            "why": "P walkers share every data tile, intensity = flop*P/bytes >> fp64 ridge (~6 flop/B): fp64-FMA bound"}
    # ---- CPU baseline beside it: oracle port (-O3 -march=native build), bounded sample
    from oracle import oracle as O
    threads = len(os.sched_getaffinity(0))  # all host cores, even when torchrun pins OMP_NUM_THREADS=1
    big = N >= 100_000
    e1, r1, t1 = _cpu_leg(c, threads, 1 if big else 200)  # calibration pass, then ~12 s of CPU work
    e, r, t = _cpu_leg(c, threads, max(1, int(round(12.0 / max(t1, 1e-3)))) * (1 if big else 200))
    cpu = {"value": e / t, "unit": "evals/s", "cores": threads, "kind": "port", "replacements_per_s": r / t,
           "sample": f"{r} replacements x {MC_STEPS} walk steps on {threads} threads, full {N}-row data, {t:.1f} s; {e} likelihood "
                     f"evaluations performed (proposals outside the box are rejected without one, BS:602-617); {O.FAST_FLAGS}"}
    return {
        "metric": "loglikelihood evals/s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _workload_config(c, K, n_runs, world),
        "replacements_per_s": value / MC_STEPS,
        "evals_note": "every proposal of the batch is scored, in or out of the prior box (the reference skips the latter)",
        "device_walk_ms_per_step": walk_ms / max(graphs, 1),
        "timing": "CUDA events on the library's stream around the K timed steps, max over ranks",
        "wall_ms_per_step": 1e3 * dt_wall / args.steps,
        "gpu_launches": int(launches),
        "clocks": clocks, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "wolfram": _wolfram_probe(),
    }


def _as_primary(extra, args, clocks=None, launches=None):
    """Promote an extras-style measurement to a full contract line (--config C5 / --mode data-sharded)."""
    line = {"metric": extra["metric"], "value": extra["value"], "unit": extra["unit"], "n_gpus": extra["n_gpus"],
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": extra.get("ms_per_step", extra.get("ms_per_iteration")),
            "higher_is_better": True, "scaling": extra["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": extra["config"], "gpu_launches": launches, "clocks": clocks, "wolfram": _wolfram_probe()}
    for k, v in extra.items():
        line.setdefault(k, v)
    return line


def run_ours(args):
    ctx = Ctx()
    eng = ctx.engine
    t_start = time.perf_counter()
    if args.mode == "data-sharded" or args.config == "C5":
        sampler = ClockSampler(ctx.local_rank)
        if ctx.rank == 0:
            sampler.start()
        l0 = eng.launch_count()
        if args.config == "C5":
            ex = measure_c5(ctx, args.steps, args.warmup)
        else:
            ex = measure_data_sharded(ctx, args.rows, iters=max(1, min(args.steps, 3)))
        clocks = sampler.stop() if ctx.rank == 0 else None
        if ctx.rank == 0:
            line = _as_primary(ex, args, clocks, int(eng.launch_count() - l0))
            if args.config == "C5":
                n, t = _cpu_gp_leg(cfg.c5_gp(), len(os.sched_getaffinity(0)), len(os.sched_getaffinity(0)))
                line["cpu_baseline"] = {"value": n / t, "unit": "evals/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                                        "sample": f"{n} covariance matrices of order 4096 (fill + LU + solve, GP:130-141), one per thread, {t:.1f} s"}
            print(json.dumps(line))
        ctx.barrier()
        os._exit(0)

    line = run_primary(ctx, args)
    extras, errors = {}, {}
    done = threading.Event()

    def emit():
        if ctx.rank == 0:
            line["extras"] = extras
            if errors:
                line["extras_errors"] = errors
            line["bench_wall_s"] = time.perf_counter() - t_start
            print(json.dumps(line), flush=True)

    def watchdog():
        if not done.wait(EXTRAS_DEADLINE_S):
            errors["watchdog"] = f"extras did not finish within {EXTRAS_DEADLINE_S:.0f} s; line printed without the missing ones"
            emit()
            os._exit(0)

    if not args.no_extras:
        threading.Thread(target=watchdog, daemon=True).start()
        plan = [("c1", lambda: measure_c1(ctx)), ("c5", lambda: measure_c5(ctx, 3, 1)), ("c4_strong", lambda: measure_c4_strong(ctx)),
                ("c4_strong_T262144", lambda: measure_c4_strong(ctx, reps=1, T=262144)),
                ("data_sharded", lambda: measure_data_sharded(ctx, args.rows))]
        for name, fn in plan:
            # all ranks must agree to enter a collective measurement: a rank-local failure aborts the remaining extras
            try:
                t0 = time.perf_counter()
                r = fn()
                r["measure_wall_s"] = time.perf_counter() - t0
                extras[name] = r
            except Exception as e:  # noqa: BLE001
                errors[name] = f"{type(e).__name__}: {e}"
                break
    done.set()
    emit()
    sys.stdout.flush()
    os._exit(0)  # the extras may leave NCCL / IPC teardown work that is not worth risking a hang on


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=list(WORKLOADS))
    ap.add_argument("--mode", default="run-sharded", choices=["run-sharded", "data-sharded"])
    ap.add_argument("--rows", type=float, default=6.4e7, help="rows of the data-sharded workload")
    ap.add_argument("--runs-per-gpu", type=int, default=1)
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
