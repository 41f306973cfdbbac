"""Registers / spills per kernel from the -Xptxas -v logs of the last build: python scripts/ptxas_regs.py <pattern>..."""
import glob
import re
import sys

pats = sys.argv[1:] or [""]
for fn in sorted(glob.glob("bayesianinference_b200/csrc/_build/*.ptxas.log")):
    log = open(fn).read()
    for b in re.split(r"(?=ptxas info\s+: Compiling entry function)", log):
        m = re.search(r"Function properties for (\S+)", b)
        if not m or not any(p in m.group(1) for p in pats):
            continue
        regs = re.search(r"Used (\d+) registers", b)
        sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
        print(f"{m.group(1)[:110]:110s} regs {regs.group(1) if regs else '?':>4s} spill {sp.groups() if sp else '?'}")
