"""One population per handle on ONE GPU, the library's own choices (layout, push mode): time per weekday and % of the HBM roofline.
usage: time_big_population.py [agents:H ...]"""
import time, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motivate_b200 as mb
cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(2_000_000, 10), (3_000_000, 10), (6_000_000, 10), (8_000_000, 10), (50_000_000, 64), (50_000_000, 10)]
for n, H in cases:
    d = mb.DevicePopulation(n, 4, H, 5, 10, seed=1)
    tab = mb.synth_tables(4, H, n, seed=1)
    w = mb.weather_pattern(40, seed=0)
    t = time.time(); s = mb.Simulation(d, tab, mb.make_params()); t2 = time.time() - t
    s.run_async(w, 0, 10); s.sync()
    s.run_async(w, 10, 40); s.sync(); ms, l = s.last_run_stats()
    print(f"{n} agents, H = {H}: create {t2:.3f} s, {ms / l:.4f} ms per weekday -> {n * l / ms * 1e3:.3e} agent-days/s, "
          f"{100 * s.algorithmic_bytes_per_day(0) * l / ms * 1e3 / 1e9 / 6560.3:.1f} % of roofline", flush=True)
    s.close(); d.close()
