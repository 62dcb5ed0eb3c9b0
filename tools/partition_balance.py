#!/usr/bin/env python
"""Kernel time of every rank's share of ONE agent-partitioned population, measured on one GPU (MVB_DEBUG_NO_EXCHANGE:
no communicator, no exchange, results meaningless after day 1, timing valid) for several settings of the partition work
model.  usage: partition_balance.py AGENTS H WORLD "cold,agent,hub10" ..."""
import os, sys
os.environ["MVB_DEBUG_NO_EXCHANGE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motivate_b200 as mb
n, H, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = mb.DevicePopulation(n, 4, H, 5, 10, seed=0)
tab = mb.synth_tables(4, H, n, seed=0)
w = mb.weather_pattern(20, seed=0)
for setting in sys.argv[4:]:
    c, a, h = setting.split(",")
    os.environ.update(MVB_PART_COLD=c, MVB_PART_AGENT=a, MVB_PART_HUB10=h)
    ms = []
    for r in range(world):
        sim = mb.Simulation.partitioned(d, tab, mb.make_params(), device=0, rank=r, world=world, unique_id=b"\0" * 128)
        sim.run_async(w, 0, 6); sim.sync()
        sim.run_async(w, 6, 14); sim.sync()
        t, l = sim.last_run_stats()
        ms.append(t / l)
        sim.close()
    print(f"cold={c} agent={a} hub10={h}: max {max(ms):.3f} mean {sum(ms) / world:.3f} ms  {[round(x, 3) for x in ms]}", flush=True)
