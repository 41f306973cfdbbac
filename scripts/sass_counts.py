"""Per-kernel SASS instruction counts of libbinest.so (cuobjdump -sass): the mnemonics that prove which hardware paths
the hot kernels use — DFMA/DMUL/DADD (fp64 pipe), DMMA (fp64 tensor pipe), UBLKCP (TMA bulk copies), LDGSTS
(cp.async), SYNCS / mbarrier traffic, BAR, ST with .CLUSTER / remote (DSMEM), ATOM/RED, MUFU.
  python scripts/sass_counts.py [out.md]"""
import collections
import re
import subprocess
import sys

LIB = "bayesianinference_b200/libbinest.so"
HOT = ["walk_grid_kernel", "walk_resident_kernel", "ns_loop_kernel", "loglike_stream_kernel", "walk_step_kernel",
       "run_update_kernel", "gp_syrk_kernel", "gp_trsm_kernel", "gp_potf2_reg_kernel", "gp_fill_kernel",
       "shard_reduce_push_kernel", "xchg_push_kernel", "xchg_gather_kernel", "evidence_sampling_kernel"]
PICK = {"C2 walk": "walk_grid_kernelINS_9OpPolyRegILi3EEELi4", "C2 stream": "loglike_stream_kernelINS_9OpPolyRegILi3EEELi8",
        "C3 stream": "loglike_stream_kernelINS_10OpLogisticILi4ELi3EEELi2", "C4 resident": "walk_resident_kernelINS_5OpGbmELi4ELi16",
        "C1 loop": "ns_loop_kernelINS_10OpGaussianELi256", "GP syrk": "gp_syrk_kernel", "GP trsm": "gp_trsm_kernel",
        "GP potf2": "gp_potf2_reg_kernel", "GP fill": "gp_fill_kernel", "shard push (C2)": "shard_reduce_push_kernelINS_9OpPolyRegILi3",
        "xchg push": "xchg_push_kernel", "xchg gather": "xchg_gather_kernel", "update": "run_update_kernel",
        "merge rank (samples)": "rank_kernelINS_9CmpSample", "merge rank (keys)": "rank_kernelINS_6CmpKey",
        "merge scan": "merge_scan_kernel", "merge chunk sort": "chunk_sort_desc_kernel", "merge join": "join_runs_kernel"}
COLS = ["DFMA", "DMUL", "DADD", "DMMA", "UBLKCP", "LDGSTS", "SYNCS", "BAR", "LDS", "STS", "LDG", "STG", "ST.E", "ATOM", "RED", "MUFU", "SHFL", "total"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts, cur = {}, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        base = op.split(".")[0]
        counts[cur][base] += 1
        if op.startswith("ST.E") or op.startswith("ST."):
            counts[cur]["ST.E"] += 1
out = ["| kernel (instantiation) | " + " | ".join(COLS) + " |", "|---|" + "---|" * len(COLS)]
for label, pat in PICK.items():
    hit = [k for k in counts if pat in k]
    if not hit:
        out.append(f"| {label}: not found | " + " | ".join("" for _ in COLS) + " |")
        continue
    k = sorted(hit, key=len)[0]
    c = counts[k]
    out.append(f"| {label} (`{pat}`) | " + " | ".join(str(c.get(col, 0)) for col in COLS) + " |")
text = ("# SASS instruction counts of the hot kernels (static, `cuobjdump -sass bayesianinference_b200/libbinest.so`, sm_100a)\n\n"
        "Static counts per kernel instantiation (not executed counts).  DFMA/DMUL/DADD = fp64 pipe; DMMA = fp64 tensor pipe "
        "(`mma.sync.m8n8k4.f64`); UBLKCP = TMA bulk copy (`cp.async.bulk`), SYNCS = mbarrier operations; LDGSTS = `cp.async`; "
        "ST.E = generic stores (DSMEM `st.async` and P2P stores into peers' buffers among them).\n\n" + "\n".join(out) + "\n")
path = sys.argv[1] if len(sys.argv) > 1 else None
if path:
    open(path, "w").write(text)
print(text)
