#!/usr/bin/env python
"""Where do warps wait?  Per CUDA source line: stall samples by reason (long/short scoreboard, barrier, wait, ...)
from an ncu report captured with --import-source on (dev tool).  usage: ncu_stalls.py report [kernel-regex] [top-n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else "day_kernel"; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}",
                      "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
reasons = ["stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_not_selected", "stall_branch_resolving",
           "stall_math", "stall_mio", "stall_lg", "stall_no_inst", "stall_selected", "stall_dispatch", "stall_membar", "stall_sleep"]
fname = None; hdr = None; lines = []
tot = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        num = lambda x: int(x) if x.isdigit() else 0
        d = {k: num(r[hdr.index(k)]) for k in reasons if k in hdr}
        lines.append((fname, int(r[0]), r[1].strip(), num(r[hdr.index("# Samples")]), d))
        for k, v in d.items(): tot[k] += v
ts = sum(l[3] for l in lines)
print("samples", ts, " by reason:", ", ".join(f"{k[6:]} {100*v/max(ts,1):.1f}%" for k, v in tot.most_common() if v))
for f, n, src, sp, d in sorted(lines, key=lambda l: -l[3])[:topn]:
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100*sp/max(ts,1):5.1f}%  {f}:{n:<4} [{top}]  {src[:90]}")
