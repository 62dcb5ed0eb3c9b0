#!/usr/bin/env python
"""One large device-generated population on one GPU for a few weekdays (profiling target for the large-population path).
usage: run_big.py AGENTS [layout] [weekdays] [H]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1])
if len(sys.argv) > 2:
    os.environ["MVB_LAYOUT"] = sys.argv[2]
days = int(sys.argv[3]) if len(sys.argv) > 3 else 8
H = int(sys.argv[4]) if len(sys.argv) > 4 else 10
import motivate_b200 as mb  # noqa: E402

d = mb.DevicePopulation(n, 4, H, 5, 10, seed=1)
tab = mb.synth_tables(4, H, n, seed=1)
w = mb.weather_pattern(2 * days, seed=0)
s = mb.Simulation(d, tab, mb.make_params())
s.run_async(w, 0, 2 * days)
s.sync()
ms, l = s.last_run_stats()
print(f"{n} agents H={H} {os.environ.get('MVB_LAYOUT', 'auto')}: {ms / l:.3f} ms/weekday, {n * l / ms * 1e3:.3e} agent-days/s, "
      f"{100 * s.algorithmic_bytes_per_day(0) * l / ms * 1e3 / 1e9 / 6560.3:.1f}% of roofline")
