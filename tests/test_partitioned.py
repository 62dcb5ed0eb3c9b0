"""Agent-partitioned mode (BASELINE config 4): one population, `world` ranks, each updating its own
agents; after every weekday the ranks exchange packed modes and all-reduce the counters.

CPU part: the partition plan (host logic of the library, no device) and a world_size-2 gloo emulation of the exchange
protocol on the oracle.  GPU part (needs 2 devices; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k partitioned`):
2 ranks against one oracle run, bit for bit, over both exchange paths (peer-to-peer stores fused into the weekday kernel; NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import motivate_b200 as mb
from motivate_b200 import _ffi
from common import assert_state_equal
from oracle.binding import OracleSimulation

N, S, H, N_DAYS = 6000, 3, 5, 16


def inputs():
    return (mb.synth_population(N, S, H, 4, 6, seed=77), mb.synth_tables(S, H, N, seed=77), mb.make_params(),
            mb.weather_pattern(N_DAYS, seed=77, p_good_given_good=0.6, p_good_given_bad=0.5))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_partition_plan_is_balanced_partition():
    pop, _, _, _ = inputs()
    deg = (np.diff(pop["social_rowptr"]) + np.diff(pop["neighbour_rowptr"])).astype(np.int64)
    for world in (1, 2, 3, 8):
        owner = mb.partition_plan(pop, S, H, world)
        assert owner.min() == 0 and owner.max() == world - 1           # everyone is owned, every rank owns someone
        # the plan's work measure: links (a hub's links count 2.5 times, layout.cpp) + per-agent work
        links = np.bincount(owner, weights=np.where(deg > 252, 2.5 * deg, deg) + 64, minlength=world)
        if world <= 3:                                                  # (slice granularity: 6 000 agents are 190 slices,
            assert links.max() / links.mean() < 1.15, (world, links)    #  the first of them all hubs)
    assert np.array_equal(mb.partition_plan(pop, S, H, 2), mb.partition_plan(pop, S, H, 2))


def gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pop, tab, prm, w = inputs()
    owner = mb.partition_plan(pop, S, H, world)
    mine = owner == rank
    # every rank keeps the whole population (as the GPU ranks do) but only its own agents' results count
    o = OracleSimulation(pop, tab, prm)
    rows = [o.stats_row()]
    prev = w[0]
    for day in range(1, N_DAYS):
        if day % 7 >= 5:
            continue
        o.step(w[day], prev != w[day]); prev = w[day]
        st = o.get_state()
        # exchange, as launch_day does it on the GPUs: every rank packs its own piece [modes | norms] padded to the longest
        # piece, ONE all-gather, every rank unpacks the others' pieces; then the counters are summed (all-reduce)
        ids = [np.flatnonzero(owner == r) for r in range(world)]
        M = max(len(x) for x in ids)
        piece = np.zeros(2 * M, np.uint8)
        piece[:len(ids[rank])] = st["current_mode"][ids[rank]]
        piece[M:M + len(ids[rank])] = st["norm"][ids[rank]]
        got = [torch.zeros(2 * M, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(got, torch.from_numpy(piece))
        mode, norm = np.zeros(N, np.uint8), np.zeros(N, np.uint8)
        for r in range(world):
            g = got[r].numpy()
            mode[ids[r]] = g[:len(ids[r])]; norm[ids[r]] = g[M:M + len(ids[r])]
        assert np.array_equal(mode, st["current_mode"]) and np.array_equal(norm, st["norm"])
        part = np.zeros(7 + S + H, np.int64)
        a, an = st["current_mode"] >= 2, st["norm"] >= 2
        for i in np.flatnonzero(mine):
            part[0] += a[i]; part[1] += an[i]; part[2] += a[i] and not an[i]; part[3] += (not a[i]) and an[i]
            if a[i]:
                part[4 + pop["commute"][i]] += 1; part[7 + pop["subculture"][i]] += 1; part[7 + S + pop["neighbourhood"][i]] += 1
        t = torch.from_numpy(part); dist.all_reduce(t)
        assert np.array_equal(t.numpy(), o.stats_row().astype(np.int64))
        rows.append(o.stats_row())
    if rank == 0:
        np.save(out, np.array(rows))
    dist.destroy_process_group()


def test_exchange_protocol_two_gloo_ranks(tmp_path):
    out = str(tmp_path / "rows.npy")
    mp.spawn(gloo_worker, args=(2, free_port(), out), nprocs=2, join=True)
    pop, tab, prm, w = inputs()
    _, rows = OracleSimulation(pop, tab, prm).run(w)
    assert np.array_equal(np.load(out), rows)


def p2p_worker(rank, world, port, vec, normvec, xrec, flags, out):
    """The peer-to-peer exchange protocol of day_kernel<.., XCHG> + exchange_finish_kernel (day_kernels.cuh), emulated on
    the oracle with shared host memory standing in for the ranks' mapped device memory: vec[p][parity] / normvec[p] /
    xrec[p][epoch parity][source] / flags[p][source] are rank p's buffers, which every rank can write."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)        # start-up barrier only: the protocol itself uses the flags
    import time
    pop, tab, prm, w = inputs()
    owner = mb.partition_plan(pop, S, H, world)
    mine = np.flatnonzero(owner == rank)
    o = OracleSimulation(pop, tab, prm)
    st = o.get_state()
    parity, epoch, rows, prev = 0, 0, [o.stats_row()], w[0]
    vec[rank][parity][:] = torch.from_numpy(st["current_mode"])          # every rank starts from the same state
    dist.barrier()
    for day in range(1, N_DAYS):
        if day % 7 >= 5:
            continue
        # the rank's copy of yesterday's modes is what the others stored into it: the update must see exactly the oracle's state
        assert np.array_equal(vec[rank][parity].numpy(), o.get_state()["current_mode"]), (rank, day)
        o.step(w[day], prev != w[day]); prev = w[day]
        st = o.get_state()
        epoch += 1
        a, an = st["current_mode"] >= 2, st["norm"] >= 2
        part = np.zeros(7 + S + H, np.int64)                             # this rank's partial statistics row: its own agents only
        for i in mine:
            part[0] += a[i]; part[1] += an[i]; part[2] += a[i] and not an[i]; part[3] += (not a[i]) and an[i]
            if a[i]:
                part[4 + pop["commute"][i]] += 1; part[7 + pop["subculture"][i]] += 1; part[7 + S + pop["neighbourhood"][i]] += 1
        for p in range(world):                                           # the kernel: new mode words into EVERY rank's vector,
            vec[p][parity ^ 1][mine] = torch.from_numpy(st["current_mode"][mine])
            normvec[p][mine] = torch.from_numpy(st["norm"][mine])
        for p in range(world):                                           # its last block: partial counters, then the flag
            xrec[p][epoch & 1][rank][:] = torch.from_numpy(part)
        for p in range(world):
            flags[p][rank] = epoch
        t0 = time.time()                                                 # exchange_finish_kernel: wait for every rank's flag ...
        while any(int(flags[rank][q]) - epoch < 0 for q in range(world)):
            assert time.time() - t0 < 60, "a rank never raised its flag"
            time.sleep(0.0005)
        total = xrec[rank][epoch & 1].sum(dim=0).numpy()                 # ... and add the partial rows up
        assert np.array_equal(total, o.stats_row().astype(np.int64)), (rank, day)
        rows.append(total.copy())
        parity ^= 1
    # nobody may be a whole weekday ahead while a rank still reads: that is what the flag wait guarantees; final state
    assert np.array_equal(vec[rank][parity].numpy(), o.get_state()["current_mode"])
    dist.barrier()
    assert np.array_equal(normvec[rank].numpy(), o.get_state()["norm"])
    if rank == 0:
        np.save(out, np.array(rows))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_to_peer_exchange_protocol_on_shared_memory(tmp_path, world):
    """CPU, `world` processes: double buffering by day parity (vectors) and epoch parity (landing zones) with the flags as
    the only synchronisation reproduces the oracle's rows and state on every rank."""
    out = str(tmp_path / "rows.npy")
    W = 7 + S + H
    vec = [[torch.zeros(N, dtype=torch.uint8).share_memory_() for _ in range(2)] for _ in range(world)]
    normvec = [torch.zeros(N, dtype=torch.uint8).share_memory_() for _ in range(world)]
    xrec = [torch.zeros(2, world, W, dtype=torch.int64).share_memory_() for _ in range(world)]
    flags = [torch.zeros(world, dtype=torch.int64).share_memory_() for _ in range(world)]
    mp.spawn(p2p_worker, args=(world, free_port(), vec, normvec, xrec, flags, out), nprocs=world, join=True)
    pop, tab, prm, w = inputs()
    _, rows = OracleSimulation(pop, tab, prm).run(w)
    assert np.array_equal(np.load(out), rows)


def wait_for_id(rank, idfile):
    if rank == 0:
        open(idfile + ".tmp", "wb").write(mb.comm_unique_id()); os.rename(idfile + ".tmp", idfile)
    import time
    while not os.path.exists(idfile):
        time.sleep(0.05)
    return open(idfile, "rb").read()


def nccl_worker(rank, world, idfile, out, exchange):
    if exchange == "nccl":
        os.environ["MVB_EXCHANGE"] = "nccl"                       # force the NCCL all-gather / all-reduce path
    if exchange == "p2p-push":                                    # + the cold links through push_kernel (layout.h), small staged set and chunks
        os.environ.update(MVB_PUSH="1", MVB_PUSH_CHUNK="512", MVB_STAGE_AGENTS="1024", MVB_LAYOUT="local")
    pop, tab, prm, w = inputs()
    sim = mb.Simulation.partitioned(pop, tab, prm, device=rank, rank=rank, world=world, unique_id=wait_for_id(rank, idfile))
    days, rows = sim.run(w)
    st = sim.get_state(0)
    np.savez(out + f".{rank}.npz", days=days, rows=rows[0], owner=sim.partition_owner(), exchange=sim.exchange_mode(), **st)
    sim.close()
    # a second handle (its own communicator), driven in pieces: a run, single steps, another run.  Epochs and flags of
    # the peer-to-peer exchange carry over from call to call; the norm words travel with the last weekday of each piece
    sim = mb.Simulation.partitioned(pop, tab, prm, device=rank, rank=rank, world=world, unique_id=wait_for_id(rank, idfile + "2"))
    sim.run_async(w, 0, 6)
    prev = w[max(d for d in range(1, 6) if d % 7 < 5)]            # the weather of the last weekday stepped
    for day in range(6, 10):
        if day % 7 < 5:
            sim.step(w[day], prev != w[day]); prev = w[day]
    sim.sync()
    mid = sim.get_state(0)
    sim.run_async(w, 10, N_DAYS)
    sim.sync()
    st2 = sim.get_state(0)
    np.savez(out + f".pieces.{rank}.npz", mid_mode=mid["current_mode"], mid_norm=mid["norm"], **st2)
    sim.close()


@pytest.mark.gpu
@pytest.mark.skipif(_ffi.load().mvb_device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange", ["p2p", "nccl", "p2p-push"])
def test_partitioned_two_gpus_equal_oracle(tmp_path, exchange):
    out, idfile = str(tmp_path / "state"), str(tmp_path / "nccl_id")
    mp.spawn(nccl_worker, args=(2, idfile, out, exchange), nprocs=2, join=True)
    pop, tab, prm, w = inputs()
    o = OracleSimulation(pop, tab, prm)
    days, rows = o.run(w)
    ref = o.get_state()
    z = [np.load(out + f".{r}.npz") for r in range(2)]
    assert np.array_equal(z[0]["owner"], z[1]["owner"]) and set(np.unique(z[0]["owner"])) == {0, 1}
    for r in range(2):
        assert str(z[r]["exchange"]) == exchange.split("-")[0], "the GPUs of one B200 box can map each other's memory"
        assert np.array_equal(z[r]["days"], days) and np.array_equal(z[r]["rows"], rows), f"rows on rank {r}"
        for k in ("current_mode", "norm", "hist"):
            assert np.array_equal(z[r][k], ref[k]), (r, k)              # every rank holds the full vectors
        mine = z[r]["owner"] == r
        assert np.array_equal(z[r]["habit"][mine].view(np.uint32), ref["habit"][mine].view(np.uint32))   # own habits
    # the handle driven in pieces: state after day 9 and at the end
    o2 = OracleSimulation(pop, tab, prm)
    prev = w[0]
    for day in range(1, 10):
        if day % 7 < 5:
            o2.step(w[day], prev != w[day]); prev = w[day]
    mid = o2.get_state()
    for r in range(2):
        zp = np.load(out + f".pieces.{r}.npz")
        assert np.array_equal(zp["mid_mode"], mid["current_mode"]) and np.array_equal(zp["mid_norm"], mid["norm"]), r
        for k in ("current_mode", "norm", "hist"):
            assert np.array_equal(zp[k], ref[k]), (r, k, "pieces")
        mine = z[r]["owner"] == r
        assert np.array_equal(zp["habit"][mine].view(np.uint32), ref["habit"][mine].view(np.uint32))


@pytest.mark.gpu
def test_partitioned_world_1_is_a_plain_run():
    pop, tab, prm, w = inputs()
    sim = mb.Simulation.partitioned(pop, tab, prm, device=0, rank=0, world=1, unique_id=None)
    days, rows = sim.run(w)
    o = OracleSimulation(pop, tab, prm)
    d2, r2 = o.run(w)
    assert np.array_equal(rows[0], r2)
    assert_state_equal(sim.get_state(0), o.get_state(), "world 1")
    assert np.all(sim.partition_owner() == 0)
    sim.close()
