#!/usr/bin/env python
"""Key raw metrics of the first matching kernel in an ncu report (dev tool). usage: ncu_raw.py report [kernel-regex]"""
import csv, subprocess, sys
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else "day_kernel"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__cycles_elapsed.avg','sm__cycles_active.avg',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
r = rows[2]
for w in want:
    if w in hdr:
        v = r[hdr.index(w)]
        try:
            if float(v.replace(',', '')) == 0: continue
        except ValueError: pass
        print(f"{w:95s} {v} {units[hdr.index(w)]}")
