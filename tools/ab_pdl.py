#!/usr/bin/env python
"""A/B of programmatic dependent launch on the few-replica configurations (same process, same populations):
MVB_NO_PDL is read once per process, so this script re-executes itself for the B arm."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "arm":
    import motivate_b200 as mb
    for R, N, H, S in ((1, 1_000_000, 10, 4), (8, 111_166, 20, 3), (8, 1_000_000, 10, 4), (1, 111_166, 20, 3)):
        pops = [mb.DevicePopulation(N, S, H, 5, 10, seed=r) for r in range(R)]
        tabs = [mb.synth_tables(S, H, N, seed=r) for r in range(R)]
        w = mb.weather_pattern(400, seed=0)
        sim = mb.Simulation(pops, tabs, mb.make_params())
        alg = sum(sim.algorithmic_bytes_per_day(r) for r in range(R))
        sim.run_async(w, 0, 60); sim.sync()
        sim.run_async(w, 60, 400); sim.sync()
        ms, l = sim.last_run_stats()
        print(f"  {R} x {N} (H={H}): {ms / l * 1e3:.2f} us/weekday, {100 * alg * l / ms * 1e3 / 1e9 / 6560.3:.1f}% of roofline", flush=True)
        sim.close()
else:
    for name, env in (("PDL on", {}), ("PDL off", {"MVB_NO_PDL": "1"}), ("PDL on", {}), ("PDL off", {"MVB_NO_PDL": "1"})):
        print(name, flush=True)
        subprocess.run([sys.executable, __file__, "arm"], env={**os.environ, **env}, check=True)
