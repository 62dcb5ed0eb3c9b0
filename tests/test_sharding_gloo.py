"""N > 1 path on the CPU: two gloo ranks shard the replicas exactly like bench.py / a multi-GPU host does, run their
shard (here on the oracle, there is no GPU), and the gathered statistics equal a serial run of all replicas."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import motivate_b200 as mb
from motivate_b200.sharding import replicas_of_rank
from oracle.binding import OracleSimulation

N_REPLICAS, N_AGENTS, N_DAYS = 5, 600, 12


def run_replica(r):
    pop = mb.synth_population(N_AGENTS, 3, 4, 3, 4, seed=r)
    tab = mb.synth_tables(3, 4, N_AGENTS, seed=r)
    o = OracleSimulation(pop, tab, mb.make_params())
    days, rows = o.run(mb.weather_pattern(N_DAYS, seed=0))
    return rows


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = replicas_of_rank(N_REPLICAS, rank, world)
    rows = {r: run_replica(r) for r in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    dist.barrier()
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        np.savez(out, **{str(k): v for k, v in merged.items()})
    dist.destroy_process_group()


def test_replica_split_is_a_partition():
    for n in (1, 5, 8, 64):
        for w in (1, 2, 3, 4, 8):
            parts = [replicas_of_rank(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_two_ranks_equal_serial(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "rows.npz")
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    assert sorted(z.files) == [str(r) for r in range(N_REPLICAS)]
    for r in range(N_REPLICAS):
        assert np.array_equal(z[str(r)], run_replica(r))
