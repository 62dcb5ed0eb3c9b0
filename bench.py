#!/usr/bin/env python
"""bench.py — agent-days/s of the per-day agent update on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 64 independent replicas x 1M
synthetic agents (S=4 subcultures, H=10 neighbourhoods, Barabási–Albert links with minimum 5 social + 10
neighbour), sharded over the N ranks with no communication (each rank steps its 64/N replicas in one launch per
weekday).  A STEP is one simulated weekday of every replica: update_habit + congestion + choose + statistics
(src/simulation.rs:108-135).  Inputs are far larger than L2 (>= 1.2 GB per rank per step), so no flush is needed.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
Under torchrun (N > 1) one process per GPU; time = max over ranks of the CUDA-event time of the K launches.

The same JSON line carries an `extra` block with the other BASELINE.json configurations, each on its own named
calendar (device-timed, not part of `value`):
  N = 1: c1 (the shipped configuration, 8 replicas x 111 166 agents x 5 years, intervention on day 365),
         c2 (1 x 1M agents x 2 years), c5 (256 scenario variants x 100k agents over one shared population, 2 years,
         every variant's day-365 intervention firing inside the timed run)
  N > 1: c4 (ONE population agent-partitioned over the N ranks): first bit-exact checks of the partitioned run at
         world = N against the CPU oracle (20k agents) and against a 1-rank GPU run (8M agents), then the timed run
         on 50M agents (N = 8) or 6M x N agents, with per-rank times.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "agent_days_per_sec"
UNIT = "agent-days/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=521, help="timed weekdays (521 = the 2-year calendar of BASELINE config 3)")
    ap.add_argument("--warmup", type=int, default=5, help="untimed weekdays")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--replicas", type=int, default=64, help="total replicas over all ranks (BASELINE config 3)")
    ap.add_argument("--agents", type=int, default=1_000_000)
    ap.add_argument("--partitioned", action="store_true",
                    help="BASELINE config 4: ONE population of --agents agents partitioned over the ranks (modes and counters exchanged every weekday: "
                         "inside the weekday kernel over peer memory, or over NCCL with MVB_EXCHANGE=nccl)")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE config 5: --replicas scenario variants (tables + ownership differ) over ONE shared population")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (c1, c2, c5 at N=1; c4 at N>1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-days", type=int, default=100)
    return ap.parse_args()


def weekdays_calendar(n_weekdays):
    """Smallest calendar [0, n_days) whose day loop (days 1..n_days-1, day%7<5) holds n_weekdays weekdays."""
    n_days, k = 1, 0
    while k < n_weekdays:
        k += (n_days % 7 < 5)
        n_days += 1
    return n_days


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(args, replica_ids, pinned=False):
    """Synthetic populations (master seed 0, replica r -> seed r), generated on all host cores; with pinned=True the
    arrays live in page-locked host memory so mvb_create_batch's H2D copies run at full PCIe speed."""
    from concurrent.futures import ThreadPoolExecutor

    import motivate_b200 as mb

    def one(r):
        p = mb.synth_population(args.agents, 4, 10, 5, 10, seed=r)        # ctypes call: releases the GIL
        if pinned:
            import torch
            p._pins = []
            for k, v in list(p.arrays.items()):
                if v is None or v.nbytes == 0:
                    continue
                t = torch.empty(v.nbytes, dtype=torch.uint8, pin_memory=True)      # page-locked host buffer
                a = t.numpy().view(v.dtype).reshape(v.shape)
                a[...] = v
                p.arrays[k] = a
                p._pins.append(t)
            p._struct = None
        return p, mb.synth_tables(4, 10, args.agents, seed=r)

    if args.sweep:
        # config 5: one population; variant v has its own tables (supportiveness +-0.1 on a subset of neighbourhoods,
        # Car capacity cut) and its own car / bike ownership (+-5 %); the link graph arrays are shared objects, so the
        # library stores the graph once on the device
        base, tab0 = one(0)
        pops, tabs = [], []
        rng = np.random.default_rng(5)
        for v in replica_ids:
            p = mb.PopulationData(dict(base.arrays))             # same array objects: same pointers
            for k in ("owns_car", "owns_bike"):
                a = base[k].copy()
                flip = rng.choice(args.agents, args.agents // 20, replace=False)
                a[flip] = (v + (k == "owns_bike")) % 2
                p.arrays[k] = a
            t = tab0.copy()
            sel = rng.random(t.H) < 0.5
            t.supportiveness[sel, 2:] += np.float32(0.1); t.supportiveness[sel, 0] -= np.float32(0.1)
            t.capacity[sel, 0] = (t.capacity[sel, 0] * 0.8).astype(np.uint32)
            pops.append(p); tabs.append(t)
        return pops, tabs
    with ThreadPoolExecutor(max_workers=min(len(replica_ids), os.cpu_count() or 1)) as ex:
        res = list(ex.map(one, replica_ids))
    return [p for p, _ in res], [t for _, t in res]


def cpu_reference_throughput(args, n_threads, n_days, warm_days=1, shape="dense", agents=None):
    """The reference's CPU path (one single-threaded replica per rayon worker, src/main.rs:89-121) on this box's
    cores, via the oracle port rebuilt here with -O3 -march=native: agent-days/s over n_days weekdays.
    shape="reference": the reference-shaped twin of the oracle (heap objects, a map per intermediate result)."""
    import motivate_b200 as mb
    from oracle.binding import OracleSimulation, synth_population
    agents = agents or args.agents
    base = [synth_population(agents, 4, 10, 5, 10, seed=s) for s in range(min(2, n_threads))]
    tabs = [mb.synth_tables(4, 10, agents, seed=s) for s in range(len(base))]
    prm = mb.make_params()
    sims = [OracleSimulation(base[i % len(base)], tabs[i % len(base)], prm, native=True, shape=shape) for i in range(n_threads)]
    w = mb.weather_pattern(warm_days + n_days + 2, seed=0)

    def work(sim, lo, hi):
        for d in range(lo, hi):
            sim.step(int(w[d]), int(w[d] != w[d - 1]))          # ctypes releases the GIL

    def run(lo, hi):
        th = [threading.Thread(target=work, args=(s, lo, hi)) for s in sims]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        return time.perf_counter() - t0

    run(1, 1 + warm_days)
    dt = run(1 + warm_days, 1 + warm_days + n_days)
    for s in sims:
        s.close()
    return n_threads * agents * n_days / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: one replica per core, K weekdays each (the reference's own parallelism)
    import motivate_b200 as mb                      # python helpers only: this arm never loads libmotivate_b200.so
    from oracle.binding import OracleSimulation, synth_population
    base = [synth_population(args.agents, 4, 10, 5, 10, seed=s) for s in range(min(2, cores))]      # oracle/libmotivate_synthhost.so
    tabs = [mb.synth_tables(4, 10, args.agents, seed=s) for s in range(len(base))]
    prm = mb.make_params()
    sims = [OracleSimulation(base[i % len(base)], tabs[i % len(base)], prm, native=True) for i in range(cores)]
    n_days = weekdays_calendar(args.warmup + args.steps)
    w = mb.weather_pattern(n_days, seed=0)
    wk = [d for d in range(1, n_days) if d % 7 < 5]
    prev = [int(w[0])]

    def step_all(d):
        th = [threading.Thread(target=s.step, args=(int(w[d]), int(w[d] != prev[0]))) for s in sims]
        [t.start() for t in th]; [t.join() for t in th]
        prev[0] = int(w[d])

    for d in wk[:args.warmup]:
        step_all(d)
    t0 = time.perf_counter()
    for d in wk[args.warmup:args.warmup + args.steps]:
        step_all(d)
    dt = time.perf_counter() - t0
    value = cores * args.agents * args.steps / dt
    sample = f"{cores} replicas x {args.agents} agents x {args.steps} weekdays, one single-threaded replica per core"
    try:
        product_loaded = "libmotivate_b200.so" in open("/proc/self/maps").read()
    except OSError:
        product_loaded = None
    assert not product_loaded, "the CPU arm must not load the product library"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, note="CPU arm: bounded sample, " + sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "product_library_loaded": product_loaded,
    }))


def l2_note(args):
    """Timing rule: inputs larger than L2 between timed iterations, or say that they are not."""
    per_rank = 171.0 * args.agents * -(-args.replicas // max(args.gpus, 1))          # algorithmic bytes per step per rank
    if per_rank >= 2 * 126e6:
        return f"inputs larger than L2 ({per_rank / 1e9:.2f} GB streamed per rank per step vs 126 MB); no flush needed"
    return (f"working set per rank per step is {per_rank / 1e6:.0f} MB, i.e. it (partly) stays in the 126 MB L2 between steps: "
            "a small-population number, not an HBM-streaming measurement")


def workload_config(args, note=None):
    if args.sweep:
        return {"workload": f"C5: {args.replicas} scenario variants (tables + car/bike ownership) x {args.agents} agents over ONE "
                            "shared population and link graph, 4 subcultures, 10 neighbourhoods, BA links 5+10",
                "replicas": args.replicas, "agents_per_replica": args.agents,
                "parallelism": f"variant-sharded x{args.gpus}, no communication",
                "l2": "link codes of the shared graph are L2-resident by design; per-variant state streams from HBM",
                **({"note": note} if note else {})}
    c = {"workload": f"C3: {args.replicas} independent replicas x {args.agents} agents, 4 subcultures, 10 neighbourhoods, "
                     "BA links min 5 social + 10 neighbour (mean 30/agent), steps = weekdays of the 2-year calendar",
         "replicas": args.replicas, "agents_per_replica": args.agents, "parallelism": f"replica-sharded x{args.gpus}, no communication",
         "l2": l2_note(args)}
    if note:
        c["note"] = note
    return c


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks", file=sys.stderr)
            sys.exit(2)
    # rank 0 prints exactly one line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION writes it to stdout) out of it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
        os.environ["NCCL_DEBUG"] = "NONE"               # NCCL prints the banner at VERSION and at WARN
    import torch
    import torch.distributed as dist

    import motivate_b200 as mb

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.partitioned:
        return run_partitioned(args, rank, local_rank, world, torch, dist, mb)
    from motivate_b200.sharding import replicas_of_rank
    my_ids = replicas_of_rank(args.replicas, rank, world)        # contiguous share, no communication
    per = len(my_ids)
    pops, tabs = make_inputs(args, my_ids, pinned=not args.no_e2e)
    prm = mb.make_params()
    n_days = weekdays_calendar(args.warmup + args.steps)
    w = mb.weather_pattern(n_days, seed=0)
    warm_end = weekdays_calendar(args.warmup)               # days [0, warm_end) hold the warm-up weekdays

    sim = mb.Simulation(pops, tabs, prm, device=local_rank)
    alg_bytes = sum(sim.algorithmic_bytes_per_day(r) for r in range(per))
    if args.warmup:
        sim.run_async(w, 0, warm_end)
    sim.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sim.run_async(w, warm_end if args.warmup else 0, n_days)
    sim.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms, launches = sim.last_run_stats()
    clocks = sampler.stop()
    assert launches == args.steps * -(-per // 64), (launches, args.steps)       # 64 replicas per launch
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    # sanity: statistics rows are populated (work was not skipped)
    days, rows = sim.fetch_rows()
    assert len(days) == 1 + args.warmup + args.steps
    for r in rows:          # work was not skipped: every row counts active agents, and the per-neighbourhood columns add up to ActiveMode
        assert int(r[-1, 0]) > 0 and int(r[-1, 0]) == int(r[-1, 7 + 4:].sum()) and not np.array_equal(r[0], r[-1])
    sim.close()

    total_agents = args.agents * args.replicas
    value = total_agents * args.steps / (max_ms * 1e-3)
    peak, peak_src = peaks()
    # per-GPU roofline of the slowest rank: replicas are dealt evenly, so every rank moves alg_bytes per launch
    ach = alg_bytes * args.steps / (max_ms * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": profile_traffic(per) if args.agents == 1_000_000 else None,
                     "traffic_source": traffic_source() if args.agents == 1_000_000 else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel": "day kernel (one launch per weekday per rank); time = max over ranks"},
        "clocks": clocks,
    }

    if not args.no_e2e:
        out["e2e"] = e2e(args, pops, tabs, prm, w, n_days, local_rank, world, dist if world > 1 else None, torch)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, dt = cpu_reference_throughput(args, cores, args.cpu_baseline_days)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{cores} replicas x {args.agents} agents x {args.cpu_baseline_days} weekdays "
                                         f"({dt:.1f} s), one single-threaded replica per core, oracle -O3 -march=native"}
        # the same update in the reference's own shape (heap objects, a map per intermediate result): what the dense
        # port above leaves out of the reference's cost; a smaller sample, it is ~60x slower per core
        rs_agents = min(args.agents, 50_000)
        v2, dt2 = cpu_reference_throughput(args, cores, 4, shape="reference", agents=rs_agents)
        out["cpu_baseline"]["reference_shaped"] = {
            "value": v2, "unit": UNIT, "cores": cores,
            "sample": f"{cores} replicas x {rs_agents} agents x 4 weekdays ({dt2:.1f} s), oracle/reference_shaped.cpp -O3 -march=native"}
    if not args.no_extra:
        del pops, tabs
        out["extra"] = extra_legs(args, rank, local_rank, world, torch, dist, mb)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def device_run(sim, w, n_days, interventions=None):
    """The whole calendar [0, n_days) on the device, device-timed; returns (ms, launches, days, rows)."""
    sim.run_async(w, 0, n_days, interventions)
    sim.sync()
    ms, launches = sim.last_run_stats()
    days, rows = sim.fetch_rows()
    return ms, launches, days, rows


def leg_line(n_agents_total, weekdays, ms, alg_bytes_per_day, peak, what, **more):
    return {"workload": what, "agent_days_per_sec": n_agents_total * weekdays / (ms * 1e-3), "weekdays": weekdays,
            "ms_per_weekday": ms / weekdays, "roofline_frac": alg_bytes_per_day * weekdays / (ms * 1e-3) / 1e9 / peak, **more}


def extra_legs(args, rank, local_rank, world, torch, dist, mb):
    peak, _ = peaks()
    prm = mb.make_params()
    extra = {}
    if world > 1:
        extra["c4"] = leg_c4(args, rank, local_rank, world, torch, dist, mb, peak)
        return extra
    # ---- c2: BASELINE configs[1] -------------------------------------------------------------------------------
    n_days = 730
    w = mb.weather_pattern(n_days, seed=0)
    pop, tab = mb.synth_population(1_000_000, 4, 10, 5, 10, seed=0), mb.synth_tables(4, 10, 1_000_000, seed=0)
    with mb.Simulation(pop, tab, prm, device=local_rank) as sim:
        device_run(sim, w, 60)                                                   # warm-up
        ms, launches, days, rows = device_run(sim, w, n_days)
        extra["c2"] = leg_line(1_000_000, launches, ms, sim.algorithmic_bytes_per_day(0), peak,
                               "C2: 1 replica x 1M agents (S=4, H=10, BA links 5+10) x 2 years = 521 weekdays on one B200",
                               l2="171 MB per weekday: partly L2-resident between launches, a small-population number")
        assert launches == 521 and int(rows[0][-1, 0]) == int(rows[0][-1, 11:].sum()) > 0
    del pop
    # ---- c5: BASELINE configs[4], every variant's intervention fires on day 365, inside the timed run ----------
    V, n = 256, 100_000
    base, tab0 = mb.synth_population(n, 4, 10, 5, 10, seed=0), mb.synth_tables(4, 10, n, seed=0)
    rng = np.random.default_rng(5)
    pops, tabs, ivs = [], [], []
    for v in range(V):
        p = mb.PopulationData(dict(base.arrays))                                 # same graph arrays: stored once on the device
        t = tab0.copy()
        sel = rng.random(t.H) < 0.5
        t.supportiveness[sel, 2:] += np.float32(0.05 * (v % 3))
        pops.append(p); tabs.append(t)
        # day 365 (src/simulation.rs:103-105): supportiveness +-0.1 and a Car capacity cut on a subset of neighbourhoods,
        # cars / bikes +-5 % of the agents (the shape of config/scenario.yaml:3-43, README.md:48-76)
        t2 = t.copy()
        sel2 = rng.random(t.H) < 0.5
        t2.supportiveness[sel2, 2:] += np.float32(0.1); t2.supportiveness[sel2, 0] -= np.float32(0.1)
        t2.capacity[sel2, 0] = (t2.capacity[sel2, 0] * 0.8).astype(np.uint32)
        own = {}
        for k, sign in (("owns_car", (-1, 0, 1)[v % 3]), ("owns_bike", (1, -1, 0)[v % 3])):
            a = base[k].copy()
            if sign:
                pool = np.flatnonzero(a == (0 if sign > 0 else 1))
                a[rng.choice(pool, n // 20, replace=False)] = 1 if sign > 0 else 0
            own[k] = a
        ivs.append(mb.InterventionData(day=365, tables=t2, owns_car=own["owns_car"], owns_bike=own["owns_bike"]))
    with mb.Simulation(pops, tabs, prm, device=local_rank) as sim:
        alg = sum(sim.algorithmic_bytes_per_day(r) for r in range(V))
        device_run(sim, w, 40)                                                   # warm-up
        ms, launches, days, rows = device_run(sim, w, n_days, ivs)
        k365 = int(np.searchsorted(days, 365))
        moved = sum(int(not np.array_equal(r[k365 - 1, 11:], r[k365 + 4, 11:])) for r in rows)
        extra["c5"] = leg_line(V * n, launches // -(-V // 64), ms, alg, peak,
                               "C5: 256 scenario variants x 100k agents over ONE shared population and link graph, 2 years; every variant's "
                               "intervention (tables + car/bike ownership) fires on day 365, inside the timed run",
                               intervention={"day": 365, "variants": V, "inside_timed_region": True, "rows_changed_after": moved,
                                             "how": "staged before the loop, applied by ONE apply_interventions_kernel launch"},
                               l2="link codes of the shared graph are L2-resident by design; per-variant state streams from HBM")
        assert moved == V
    del pops, tabs, ivs, base
    # ---- c1: BASELINE configs[0], the shipped configuration + the three documented fixes -------------------------
    import tempfile
    from motivate_b200 import c1, simulation
    from motivate_b200.scenario import Parameters
    from motivate_b200.weather import make_pattern
    d = tempfile.mkdtemp()
    pp, sp = c1.materialise(d)
    P = Parameters.from_file(pp)
    preps = [simulation.prepare(str(i), True, "", sp, P.number_of_people, P.social_connectivity, P.subculture_connectivity,
                                P.neighbourhood_connectivity, P.number_of_neighbour_links, P.days_in_habit_average, P.distributions,
                                number_of_social_network_links=P.number_of_social_network_links)
             for i in range(1, P.number_of_simulations + 1)]
    n_days = 365 * P.total_years
    w = make_pattern(n_days)                                                     # the reference's own rain sequence (HC-128, zero seed)
    with mb.Simulation([p.population for p in preps], [p.tables for p in preps], preps[0].params, device=local_rank) as sim:
        alg = sum(sim.algorithmic_bytes_per_day(r) for r in range(len(preps)))
        device_run(sim, w, 60)
        ms, launches, days, rows = device_run(sim, w, n_days, [p.intervention for p in preps])
        extra["c1"] = leg_line(P.number_of_people * len(preps), launches, ms, alg, peak,
                               f"C1: the reference's shipped configuration (examples/c1: {P.number_of_people} agents, 3 subcultures, 20 "
                               f"neighbourhoods, intervention on day 365) x {len(preps)} simulations x {P.total_years} years, HC-128 rain sequence",
                               l2="19 MB per replica per weekday: L2-resident, a small-population number",
                               parity="tests/test_gpu_c1.py: bit-exact against the oracle every weekday of years 1-2")
        assert launches == 1304
    return extra


def leg_c4(args, rank, local_rank, world, torch, dist, mb, peak):
    """BASELINE configs[3]: ONE population agent-partitioned over the ranks.  Populations are generated on each rank's own
    GPU from the same seed (mvb_synth_create_device: identical on every rank up to the order inside a link list, which the
    update does not depend on)."""
    from oracle.binding import OracleSimulation
    prm = mb.make_params()
    dev = torch.device("cuda", local_rank)

    def uid():
        u = [mb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(u, src=0)
        return u[0]

    def all_ok(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def same_state(a, b, owner_mask):
        return (np.array_equal(a["current_mode"], b["current_mode"]) and np.array_equal(a["norm"], b["norm"]) and
                np.array_equal(a["hist"], b["hist"]) and
                np.array_equal(a["habit"][owner_mask].view(np.uint32), b["habit"][owner_mask].view(np.uint32)))

    out = {}
    # 1. against the CPU oracle: 20k agents, 6 weeks, an intervention on a Sunday; once staged whole (GLOBAL layout) and once
    #    with a small staged set in the LOCAL layout (hub region + own neighbourhood prefix: the large-population path)
    n = 20_000
    d = mb.DevicePopulation(n, 4, 10, 5, 10, seed=11, device=local_rank)
    p = d.to_host()
    tab = mb.synth_tables(4, 10, n, seed=11)
    w = mb.weather_pattern(43, seed=11)
    tab2 = tab.copy(); tab2.capacity[:, 0] //= 2; tab2.supportiveness[::2, 2:] += np.float32(0.1)
    iv = mb.InterventionData(day=20, tables=tab2, owns_car=1 - p["owns_car"], owns_bike=None)
    o = OracleSimulation(p, tab, prm)
    d2, r2 = o.run(w, intervention=iv)
    so = o.get_state()
    ok = True
    modes_seen = set()
    # both layouts over the default exchange (peer-to-peer stores fused into the weekday kernel where the GPUs can map each
    # other's memory), and the LOCAL layout once more over the NCCL all-gather / all-reduce path
    for env in ({}, {"MVB_LAYOUT": "local", "MVB_STAGE_AGENTS": "2048"}, {"MVB_LAYOUT": "local", "MVB_STAGE_AGENTS": "2048", "MVB_EXCHANGE": "nccl"}):
        os.environ.update(env)
        sim = mb.Simulation.partitioned(d, tab, prm, device=local_rank, rank=rank, world=world, unique_id=uid())
        for k in env:
            os.environ.pop(k)
        modes_seen.add(sim.exchange_mode())
        days, rows = sim.run(w, interventions=iv)
        st = sim.get_state(0)
        mine = sim.partition_owner() == rank
        sim.close()
        ok = ok and np.array_equal(days, d2) and np.array_equal(rows[0], r2) and same_state(st, so, mine)
    out["partitioned_parity_vs_oracle"] = {"agents": n, "weekdays": len(days) - 1, "world": world, "layouts": ["global", "local"],
                                            "exchange_paths": sorted(modes_seen),
                                            "result": "bit-exact" if all_ok(ok) else "MISMATCH",
                                            "compared": "statistics rows, every agent's mode and norm, habit bits of the agents each rank owns"}
    d.close()
    # 2. against a 1-rank GPU run: 8M agents (most links gathered from L2: the large-population path), 12 weekdays
    n = 8_000_000
    d = mb.DevicePopulation(n, 4, 10, 5, 10, seed=12, device=local_rank)
    tab = mb.synth_tables(4, 10, n, seed=12)
    w = mb.weather_pattern(17, seed=12)
    sim = mb.Simulation.partitioned(d, tab, prm, device=local_rank, rank=rank, world=world, unique_id=uid())
    exchange_8m = sim.exchange_mode()
    days, rows = sim.run(w)
    st = sim.get_state(0)
    mine = sim.partition_owner() == rank
    sim.close()
    one = mb.Simulation(d, tab, prm, device=local_rank)
    d1, r1 = one.run(w)
    s1 = one.get_state(0)
    one.close()
    ok = np.array_equal(days, d1) and np.array_equal(rows[0], r1[0]) and same_state(st, s1, mine)
    import zlib
    out["partitioned_parity_vs_one_gpu"] = {"agents": n, "weekdays": len(days) - 1, "world": world, "exchange": exchange_8m,
                                             "result": "bit-exact" if all_ok(ok) else "MISMATCH",
                                             "mode_vector_crc32": zlib.crc32(st["current_mode"].tobytes()),
                                             "norm_vector_crc32": zlib.crc32(st["norm"].tobytes())}
    out["partitioned_parity"] = ("bit-exact" if out["partitioned_parity_vs_oracle"]["result"] == "bit-exact" and
                                 out["partitioned_parity_vs_one_gpu"]["result"] == "bit-exact" else "MISMATCH")
    d.close()
    del st, s1
    # 3. the timed runs: H = 10 neighbourhoods (the shape of C2, as in round 1: a neighbourhood of 5M agents is 5x what shared
    #    memory holds) and H = 64 (781k agents per neighbourhood at 50M: every neighbour link is a shared-memory gather)
    n = 50_000_000 if world == 8 else 6_000_000 * world
    n_days = weekdays_calendar(args.warmup + args.steps)
    w = mb.weather_pattern(n_days, seed=0)
    warm_end = weekdays_calendar(args.warmup)

    def timed(H, exchange=None):
        d = mb.DevicePopulation(n, 4, H, 5, 10, seed=0, device=local_rank)
        tab = mb.synth_tables(4, H, n, seed=0)
        t0 = time.perf_counter()
        if exchange:
            os.environ["MVB_EXCHANGE"] = exchange
        sim = mb.Simulation.partitioned(d, tab, prm, device=local_rank, rank=rank, world=world, unique_id=uid())
        os.environ.pop("MVB_EXCHANGE", None)
        t_create = time.perf_counter() - t0
        alg = sim.algorithmic_bytes_per_day(0)
        mode = sim.exchange_mode()
        sim.run_async(w, 0, warm_end); sim.sync()
        dist.barrier(); torch.cuda.synchronize()
        sim.run_async(w, warm_end, n_days); sim.sync()
        torch.cuda.synchronize(); dist.barrier()
        ms, launches = sim.last_run_stats()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        per_rank = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(per_rank, t)
        max_ms = max(float(x.item()) for x in per_rank)
        free_b, total_b = torch.cuda.mem_get_info(dev)
        sim.close(); d.close()
        return {"neighbourhoods": H, "exchange": mode, "agent_days_per_sec": n * args.steps / (max_ms * 1e-3), "weekdays": args.steps,
                "ms_per_weekday": max_ms / args.steps, "ms_per_weekday_by_rank": [float(x.item()) / args.steps for x in per_rank],
                "roofline_frac_per_gpu": alg * args.steps / (max_ms * 1e-3) / 1e9 / world / peak,
                "create_seconds_rank0": t_create, "device_memory_in_use_gb_rank0": (total_b - free_b) / 1e9}

    out["workload"] = (f"C4: ONE population of {n} agents (S=4, BA links 5+10) agent-partitioned over {world} B200; the exchange of "
                       "the packed 2-bit modes and of the day's counters is fused into the weekday kernel (stores into every rank's "
                       "vector over NVLink peer memory, flags; `exchange` = p2p) or, as the baseline beside it, one NCCL all-gather + "
                       "one all-reduce per weekday (`exchange_nccl`); populations generated on the GPUs")
    out["agents"] = n
    out.update(timed(10))
    out["h64"] = timed(64)
    nccl = timed(64, "nccl")
    out["h64"]["exchange_nccl"] = {k: nccl[k] for k in ("exchange", "agent_days_per_sec", "ms_per_weekday", "roofline_frac_per_gpu")}
    return out


def run_partitioned(args, rank, local_rank, world, torch, dist, mb):
    """BASELINE config 4: one population, agents partitioned over the ranks; per weekday each rank updates its agents,
    and the ranks exchange packed modes and counters (inside the weekday kernel over peer memory, or NCCL all-gather + all-reduce)."""
    pop = mb.synth_population(args.agents, 4, 10, 5, 10, seed=0)          # same population on every rank
    tab = mb.synth_tables(4, 10, args.agents, seed=0)
    prm = mb.make_params()
    n_days = weekdays_calendar(args.warmup + args.steps)
    w = mb.weather_pattern(n_days, seed=0)
    warm_end = weekdays_calendar(args.warmup)
    uid = [mb.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    sim = mb.Simulation.partitioned(pop, tab, prm, device=local_rank, rank=rank, world=world, unique_id=uid[0])
    alg_bytes = sim.algorithmic_bytes_per_day(0)
    exchange = sim.exchange_mode()
    if args.warmup:
        sim.run_async(w, 0, warm_end)
    sim.sync()
    sampler = ClockSampler(local_rank); sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sim.run_async(w, warm_end if args.warmup else 0, n_days)
    sim.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms, launches = sim.last_run_stats()
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    per_rank = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:
        per_rank = [t]
    max_ms = float(t.item())
    days, rows = sim.fetch_rows()
    sim.close()
    peak, peak_src = peaks()
    ach = alg_bytes * args.steps / (max_ms * 1e-3) / 1e9 / world          # per GPU: every rank moves 1/world of the bytes
    out = {"metric": METRIC, "value": args.agents * args.steps / (max_ms * 1e-3), "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"C4: ONE population of {args.agents} agents (4 subcultures, 10 neighbourhoods, BA links 5+10) "
                                  f"agent-partitioned over {world} ranks; exchange of the packed 2-bit modes and the day's counters: "
                                  + ("fused into the weekday kernel over NVLink peer memory" if exchange == "p2p" else
                                     "one NCCL all-gather + one all-reduce per weekday"), "agents": args.agents, "exchange": exchange,
                      "parallelism": f"agent-partitioned x{world}", "l2": "inputs larger than L2; no flush needed"},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / world,
                        "kernel": "push kernel + day kernel + exchange, per rank"},
           "clocks": clocks, "active_mode_last_row": int(rows[0][-1, 0]),
           "ms_per_step_by_rank": [float(x.item()) / args.steps for x in per_rank]}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def traffic_source():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        j = json.load(open(p))
        return ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of the day kernel from the committed ncu --set full "
                f"capture {j.get('capture', 'under profiles/')}, per 1M-agent replica, times the replicas per launch")
    except Exception:
        return None


def profile_traffic(replicas_on_rank):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the day kernel, from the committed `ncu --set full`
    capture (profiles/traffic.json holds it per replica of 1M agents; a launch steps `replicas_on_rank` of them)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["dram_bytes_per_replica_launch"] * replicas_on_rank
        except Exception:
            return None
    return None


def e2e(args, pops, tabs, prm, w, n_days, local_rank, world, dist, torch):
    """Same metric through the public call a host makes, with HOST buffers: mvb_create_batch (population H2D from
    pinned memory) + mvb_run over the same calendar (weather in, statistics rows D2H) + mvb_destroy."""
    import motivate_b200 as mb
    h2d = sum(sum(v.nbytes for v in p.arrays.values() if v is not None) for p in pops) + n_days
    # context for the number below: what the same pinned buffers need as plain copies on this box (the PCIe floor of set-up)
    floor = None
    pins = [t for p in pops for t in getattr(p, "_pins", [])]
    if pins:
        dst = torch.empty(max(t.numel() for t in pins), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        tf = time.perf_counter()
        for t in pins:
            dst[:t.numel()].copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        floor = time.perf_counter() - tf
        del dst
    # one untimed pass of the same call sequence first (the warm-up rule: the GPU's clocks drop while the host generates
    # populations, and a create that starts on an idle GPU runs its layout kernels at idle clocks: 0.37 s instead of 0.21 s)
    for timed in (False, True):
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sim = mb.Simulation(pops, tabs, prm, device=local_rank)
        t1 = time.perf_counter()
        days, rows = sim.run(w, n_days)
        t2 = time.perf_counter()
        sim.close()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t3 = t0 + dt                                       # rank-local end, before the max over ranks
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    weekdays = len(days) - 1
    d2h = sum(r.nbytes for r in rows)
    return {"value": args.agents * args.replicas * weekdays / dt, "unit": UNIT,
            "h2d_bytes_per_step": h2d / weekdays, "d2h_bytes_per_step": d2h / weekdays,
            "weekdays": weekdays, "seconds": dt,
            "seconds_rank0": {"create": t1 - t0, "run": t2 - t1, "destroy": t3 - t2, "plain_h2d_of_the_same_buffers": floor},
            "what": "mvb_create_batch (H2D of populations from pinned host memory) + mvb_run (rows D2H) + mvb_destroy; second of two "
                    "identical passes (the first is the warm-up)"}


if __name__ == "__main__":
    main()
