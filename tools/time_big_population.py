import time, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motivate_b200 as mb
for n in (8_000_000, 50_000_000):
    d=mb.DevicePopulation(n,4,10,5,10,seed=1)
    tab=mb.synth_tables(4,10,n,seed=1)
    w=mb.weather_pattern(24,seed=0)
    for layout in ("global","local"):
        os.environ["MVB_LAYOUT"]=layout
        t=time.time(); s=mb.Simulation(d,tab,mb.make_params()); t2=time.time()-t
        s.run_async(w,0,10); s.sync()
        s.run_async(w,10,24); s.sync(); ms,l=s.last_run_stats()
        print(n, layout, "create %.3f s, %d weekdays %.3f ms/day -> %.3e agent-days/s, %.1f%% of roofline"%(t2,l,ms/l,n*l/ms*1e3, 100*s.algorithmic_bytes_per_day(0)*l/ms*1e3/1e9/6560.3), flush=True)
        s.close()
    d.close()
