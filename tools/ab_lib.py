import os, sys
sys.path.insert(0, "/root/repo")
import motivate_b200 as mb
for R, N, H, S in ((8, 111_166, 20, 3), (8, 1_000_000, 10, 4)):
    pops = [mb.DevicePopulation(N, S, H, 5, 10, seed=r) for r in range(R)]
    tabs = [mb.synth_tables(S, H, N, seed=r) for r in range(R)]
    w = mb.weather_pattern(400, seed=0)
    sim = mb.Simulation(pops, tabs, mb.make_params())
    sim.run_async(w, 0, 60); sim.sync(); sim.run_async(w, 60, 400); sim.sync()
    ms, l = sim.last_run_stats()
    print(f"  {os.environ.get('MVB_LIB_PATH', 'NEW').split('_')[-1]}: {R} x {N}: {ms / l * 1e3:.2f} us/weekday", flush=True)
    sim.close()
