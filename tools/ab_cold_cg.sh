# A/B on one GPU: cold gathers through ld.global.nc (__ldg, MVB_COLD_CG=0) vs ld.global.cg (MVB_COLD_CG=1); unset = the
# library's own choice by the size of the unstaged part of the mode vector (create.cuh).
cd /root/repo
cat > /tmp/t8m.py <<'P'
import os, sys
sys.path.insert(0, "/root/repo")
import motivate_b200 as mb
for n, H in ((8_000_000, 10), (2_000_000, 10), (1_000_000, 10)):
    d = mb.DevicePopulation(n, 4, H, 5, 10, seed=1); tab = mb.synth_tables(4, H, n, seed=1); w = mb.weather_pattern(60, seed=0)
    s = mb.Simulation(d, tab, mb.make_params()); s.run_async(w, 0, 10); s.sync(); s.run_async(w, 10, 60); s.sync(); ms, l = s.last_run_stats()
    print("MVB_COLD_CG=" + os.environ.get("MVB_COLD_CG", "auto"), n, H, "%.4f ms/day %.1f%% of roofline" % (ms / l, 100 * s.algorithmic_bytes_per_day(0) * l / ms * 1e3 / 1e9 / 6560.3), flush=True)
    s.close(); d.close()
P
for k in 1 2; do
MVB_COLD_CG=0 python /tmp/t8m.py
MVB_COLD_CG=1 python /tmp/t8m.py
python /tmp/t8m.py
done
