#!/usr/bin/env python
"""The day loop on the device, seen from inside the kernel: for every day-kernel launch the time its first block started
and its last block ended (%globaltimer, ns; debug hook mvbx_launch_stamps).  gap(d) = start(d+1) - end(d): positive =
the GPU idles between two weekdays, negative = tomorrow's first blocks were already running their launch-independent
prologue while today's last blocks finished (programmatic dependent launch).  Re-executes itself with MVB_NO_PDL=1."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the stamps live in a debug build of the library (make -C motivate_b200/csrc stamps, here, before gpurun)
os.environ.setdefault("MVB_LIB_PATH", os.path.join(ROOT, "motivate_b200", "libmotivate_b200_stamps.so"))
if len(sys.argv) > 1 and sys.argv[1] == "arm":
    import numpy as np
    import motivate_b200 as mb
    from motivate_b200 import _ffi
    lib = _ffi.load()
    fn = lib.mvbx_launch_stamps
    fn.restype, fn.argtypes = C.c_int, [_ffi.SimP, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    for R, N, H, S in ((1, 1_000_000, 10, 4), (8, 111_166, 20, 3), (64, 1_000_000, 10, 4)):
        pops = [mb.DevicePopulation(N, S, H, 5, 10, seed=r) for r in range(R)]
        tabs = [mb.synth_tables(S, H, N, seed=r) for r in range(R)]
        w = mb.weather_pattern(200, seed=0)
        sim = mb.Simulation(pops, tabs, mb.make_params())
        sim.run_async(w, 0, 200); sim.sync()          # (also makes room for the rows: the record buffer must not move below)
        sim.run_async(w, 0, 30); sim.sync()
        assert fn(sim._h, 100, None, None) == 0
        sim.run_async(w, 30, 170); sim.sync()
        ms, l = sim.last_run_stats()
        buf = np.zeros(400, np.uint64); n = C.c_uint32()
        assert fn(sim._h, 0, buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(n)) == 0
        st = buf[:2 * n.value].reshape(-1, 2).astype(np.int64)
        gaps = (st[1:, 0] - st[:-1, 1]) / 1e3
        dur = (st[:, 1] - st[:, 0]) / 1e3
        print(f"  {R} x {N}: {ms / l * 1e3:.2f} us per weekday (CUDA events); first-block-start to last-block-end {np.median(dur):.2f} us; "
              f"gap to the next launch: median {np.median(gaps):+.2f} us, p10 {np.percentile(gaps, 10):+.2f}, p90 {np.percentile(gaps, 90):+.2f}  ({n.value} launches)", flush=True)
        sim.close()
else:
    for name, env in (("programmatic dependent launch ON", {}), ("OFF (MVB_NO_PDL=1)", {"MVB_NO_PDL": "1"})):
        print(name, flush=True)
        subprocess.run([sys.executable, __file__, "arm"], env={**os.environ, **env}, check=True)
