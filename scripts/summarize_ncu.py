"""Summarise ncu outputs into small text files for profiles/ (the .ncu-rep files themselves stay in gpurun_out/).
  python scripts/summarize_ncu.py launches <launches.csv> <out.md>
  python scripts/summarize_ncu.py full <report.ncu-rep> <out.md>
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            try:
                agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | mean us | total ms | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:110]}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / 1e6:.3f} | {sum(v) / tot:.3f} |\n")
        f.write(f"\ntotal {tot / 1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches "
                "(ncu: cold-cache, serialised — compare shares, not absolutes)\n")


def full(rep, out):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## {name[:140]}\n\n| metric | unit | value |\n|---|---|---|\n")
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"| {h} | {units[i]} | {r[i]} |\n")
            st = []
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                    try:
                        st.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            f.write("\nwarp stall reasons (warps per issue-active cycle): " +
                    ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]) + "\n\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
