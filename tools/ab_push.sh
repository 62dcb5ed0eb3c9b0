# A/B on one GPU: the cold links of large single populations through the pull path (MVB_PUSH=0) and through push_kernel (MVB_PUSH=1)
cd /root/repo
cat > /tmp/tpush.py <<'P'
import os, sys, time
sys.path.insert(0, "/root/repo")
import motivate_b200 as mb
for n, H in ((8_000_000, 10), (2_000_000, 10), (50_000_000, 64), (50_000_000, 10)):
    d = mb.DevicePopulation(n, 4, H, 5, 10, seed=1); tab = mb.synth_tables(4, H, n, seed=1); w = mb.weather_pattern(40, seed=0)
    for push in ("0", "1"):
        os.environ["MVB_PUSH"] = push
        t = time.time(); s = mb.Simulation(d, tab, mb.make_params()); tc = time.time() - t
        s.run_async(w, 0, 10); s.sync(); s.run_async(w, 10, 40); s.sync(); ms, l = s.last_run_stats()
        print("MVB_PUSH=" + push, n, H, "create %.2f s, %.4f ms/day %.1f%% of roofline" % (tc, ms / l, 100 * s.algorithmic_bytes_per_day(0) * l / ms * 1e3 / 1e9 / 6560.3), flush=True)
        s.close()
    d.close()
P
python /tmp/tpush.py
