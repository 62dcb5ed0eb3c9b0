#!/usr/bin/env python
"""Day-kernel time of R replicas x N agents for several forced part counts (MVB_PARTS), one population set."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motivate_b200 as mb
R, N = int(sys.argv[1]), int(sys.argv[2])
pops = [mb.DevicePopulation(N, 4, 10, 5, 10, seed=r) for r in range(R)]
tabs = [mb.synth_tables(4, 10, N, seed=r) for r in range(R)]
w = mb.weather_pattern(80, seed=0)
sim = mb.Simulation(pops, tabs, mb.make_params())
alg = sum(sim.algorithmic_bytes_per_day(r) for r in range(R))
for P in [None] + [int(x) for x in sys.argv[3:]]:
    if P: os.environ["MVB_PARTS"] = str(P)
    sim.run_async(w, 0, 15); sim.sync()
    sim.run_async(w, 15, 71); sim.sync()
    ms, l = sim.last_run_stats()
    print(f"P={P}: {ms / l * 1e3:.1f} us/weekday, {100 * alg * l / ms * 1e3 / 1e9 / 6560.3:.1f}% of roofline", flush=True)
