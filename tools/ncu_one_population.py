import os, sys
sys.path.insert(0, "/root/repo")
import motivate_b200 as mb
n, H = int(sys.argv[1]), int(sys.argv[2])
d = mb.DevicePopulation(n, 4, H, 5, 10, seed=1); tab = mb.synth_tables(4, H, n, seed=1); w = mb.weather_pattern(12, seed=0)
s = mb.Simulation(d, tab, mb.make_params()); s.run_async(w, 0, 12); s.sync(); s.close(); d.close()
