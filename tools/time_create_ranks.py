#!/usr/bin/env python
"""mvb_create_batch when N ranks of one box create at the same time (the e2e leg of bench.py at N GPUs): every rank makes
`--replicas` populations of 1M agents in pinned host memory; then, several times: barrier, plain copies of the buffers,
barrier, create (MVB_TIMING=1 prints the library's own breakdown on stderr), destroy.  usage: torchrun ... time_create_ranks.py"""
import argparse, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motivate_b200 as mb  # noqa: E402
ap = argparse.ArgumentParser(); ap.add_argument("--replicas", type=int, default=8); ap.add_argument("--repeat", type=int, default=4)
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pops = []
for r in range(a.replicas):
    p = mb.synth_population(1_000_000, 4, 10, 5, 10, seed=rank * 100 + r); p._pins = []
    for k, v in list(p.arrays.items()):
        if v is None or v.nbytes == 0: continue
        t = torch.empty(v.nbytes, dtype=torch.uint8, pin_memory=True)
        arr = t.numpy().view(v.dtype).reshape(v.shape); arr[...] = v
        p.arrays[k] = arr; p._pins.append(t)
    p._struct = None; pops.append(p)
tabs = [mb.synth_tables(4, 10, 1_000_000, seed=r) for r in range(a.replicas)]
dst = torch.empty(max(t.numel() for p in pops for t in p._pins), dtype=torch.uint8, device="cuda")
for rep in range(a.repeat):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for p in pops:
        for t in p._pins: dst[:t.numel()].copy_(t, non_blocking=True)
    torch.cuda.synchronize(); t_copy = time.perf_counter() - t0
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    sim = mb.Simulation(pops, tabs, mb.make_params(), device=lr)
    t1 = time.perf_counter(); sim.close(); t2 = time.perf_counter()
    ts = torch.tensor([t_copy, t1 - t0, t2 - t1], dtype=torch.float64, device="cuda")
    allt = [torch.zeros_like(ts) for _ in range(world)]; dist.all_gather(allt, ts)
    if rank == 0:
        print(f"pass {rep}: plain copy {[round(float(x[0]) * 1e3, 1) for x in allt]} ms; create {[round(float(x[1]) * 1e3, 1) for x in allt]} ms; destroy {[round(float(x[2]) * 1e3, 1) for x in allt]} ms", flush=True)
dist.destroy_process_group()
