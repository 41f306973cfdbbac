"""Data-sharded mode timing: C2-shaped data with N rows split over the ranks (torchrun), one iteration = K walkers x S steps.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/sharded_bench.py [N] [K] [iters]"""
import os, sys, time
sys.path.insert(0, '.')
import torch, torch.distributed as dist
from bayesianinference_b200 import engine, configs as cfg

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
comm = None
if world > 1:
    dist.init_process_group("gloo")
    engine.init(device=int(os.environ["LOCAL_RANK"]))
    comm = engine.Comm(rank, world)
else:
    engine.init(device=0)
c = cfg.c2_polyreg(N=N)
p = engine.Problem.from_config(c, comm=comm)
o = engine.default_options(pool_size=1024, batch_k=K, mc_steps=200, max_iter=10**9, min_iter=10**9, seed=3)
run = engine.RunGroup(p, o)
run.advance(1)
torch.cuda.synchronize()
t0 = time.perf_counter()
run.advance(iters)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / iters
if rank == 0:
    print(f"data-sharded x{world}: N={N} rows, K={K}: {dt*1e3:.2f} ms/iteration, {K*200/dt:.0f} evals/s, "
          f"{K*200/dt*N*9/1e12:.2f} TFLOP/s aggregate")
if world > 1:
    dist.barrier(); comm.close(); dist.destroy_process_group()
