#!/usr/bin/env python
"""ONE device-generated population agent-partitioned over the ranks of a torchrun launch (timing only).
usage: torchrun ... tools/run_partitioned.py AGENTS [H] [weekdays]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motivate_b200 as mb  # noqa: E402

n = int(sys.argv[1]); H = int(sys.argv[2]) if len(sys.argv) > 2 else 10; days = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["NCCL_DEBUG"] = "NONE"
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
u = [mb.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(u, src=0)
d = mb.DevicePopulation(n, 4, H, 5, 10, seed=0, device=lr)
tab = mb.synth_tables(4, H, n, seed=0)
sim = mb.Simulation.partitioned(d, tab, mb.make_params(), device=lr, rank=rank, world=world, unique_id=u[0])
w = mb.weather_pattern(4 * days, seed=0)
sim.run_async(w, 0, 8); sim.sync()
# kernel time of this rank alone: a few weekdays WITHOUT the exchange would desynchronise the ranks, so time whole days
dist.barrier(); torch.cuda.synchronize()
sim.run_async(w, 8, 8 + days); sim.sync()
ms, l = sim.last_run_stats()
t = torch.tensor([ms / l], dtype=torch.float64, device="cuda")
allt = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(allt, t)
own = (sim.partition_owner() == rank).sum()
cnt = torch.tensor([int(own)], dtype=torch.int64, device="cuda")
allc = [torch.zeros_like(cnt) for _ in range(world)]
dist.all_gather(allc, cnt)
if rank == 0:
    m = max(float(x) for x in allt)
    print("per rank ms/weekday", [round(float(x), 3) for x in allt])
    print(f"{n} agents H={H} world={world} exchange={sim.exchange_mode()}: {m:.3f} ms/weekday = {n / m * 1e3:.3e} agent-days/s, "
          f"{100 * sim.algorithmic_bytes_per_day(0) / world / m * 1e3 / 1e9 / 6560.3:.1f}% of roofline per GPU; agents per rank {[int(c) for c in allc]}")
sim.close()
dist.destroy_process_group()
