#!/usr/bin/env python
"""Where mvb_create_batch's time goes: R replicas x N agents from pinned host buffers, several times, next to the
time the same bytes need as plain cudaMemcpy (the PCIe floor of set-up from host-resident populations)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motivate_b200 as mb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=64)
ap.add_argument("--agents", type=int, default=1_000_000)
ap.add_argument("--repeat", type=int, default=3)
a = ap.parse_args()
from concurrent.futures import ThreadPoolExecutor


def one(r):
    p = mb.synth_population(a.agents, 4, 10, 5, 10, seed=r)
    p._pins = []
    for k, v in list(p.arrays.items()):
        if v is None or v.nbytes == 0:
            continue
        t = torch.empty(v.nbytes, dtype=torch.uint8, pin_memory=True)
        arr = t.numpy().view(v.dtype).reshape(v.shape)
        arr[...] = v
        p.arrays[k] = arr
        p._pins.append(t)
    p._struct = None
    return p


with ThreadPoolExecutor(os.cpu_count()) as ex:
    pops = list(ex.map(one, range(a.replicas)))
tabs = [mb.synth_tables(4, 10, a.agents, seed=r) for r in range(a.replicas)]
nbytes = sum(sum(v.nbytes for v in p.arrays.values() if v is not None) for p in pops)
dst = torch.empty(max(t.numel() for p in pops for t in p._pins), dtype=torch.uint8, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for p in pops:
        for t in p._pins:
            dst[:t.numel()].copy_(t, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"plain H2D of the same buffers: {nbytes / 1e9:.2f} GB in {dt:.3f} s = {nbytes / dt / 1e9:.1f} GB/s", flush=True)
for rep in range(a.repeat):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sim = mb.Simulation(pops, tabs, mb.make_params())
    t1 = time.perf_counter()
    sim.close()
    print(f"mvb_create_batch: {t1 - t0:.3f} s, destroy {time.perf_counter() - t1:.3f} s", flush=True)
