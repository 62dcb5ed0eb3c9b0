#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample totals from an ncu report (dev tool).
usage: ncu_lines.py report.ncu-rep [kernel-regex] [min-percent]"""
import csv, subprocess, sys
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else "day_kernel"; minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}",
                      "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; lines = []; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        ie = hdr.index("Instructions Executed"); sp = hdr.index("# Samples")
        num = lambda x: int(x) if x.isdigit() else 0
        lines.append((fname, int(r[0]), r[1].strip(), num(r[ie]), num(r[sp])))
tot = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print(f"total warp-instructions {tot}, samples {ts}")
for f, n, src, ie, sp in lines:
    if ie > tot * minp / 100 or sp > ts * minp / 100:
        print(f"{100*ie/tot:5.1f}% inst {100*sp/max(ts,1):5.1f}% samp  {f}:{n:<4} {src[:120]}")
