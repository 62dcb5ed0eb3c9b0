"""Set-up on the device (motivate_b200/csrc/create.cuh, setup_kernels.cuh): the layout the GPU builds is the one the host
builder (layout.cpp, the specification; its invariants are tested on the CPU in test_layout.py) builds, entry for
entry; populations staged only partly in shared memory (cold links) give the oracle's results at small sizes too."""
import ctypes as C
import os

import numpy as np
import pytest

import motivate_b200 as mb
from common import assert_state_equal
from motivate_b200 import _ffi
from oracle.binding import OracleSimulation

pytestmark = pytest.mark.gpu


def compare_with_host(sim, pop, r=0):
    lib = _ffi.load()
    fn = lib.mvbx_compare_layout_with_host
    fn.restype, fn.argtypes = C.c_int, [_ffi.SimP, C.c_uint32, _ffi.PopP, C.c_char_p, C.c_size_t]
    msg = C.create_string_buffer(256)
    rc = fn(sim._h, r, C.byref(pop.struct()), msg, 256)
    assert rc == 0, (rc, msg.value.decode())


@pytest.fixture
def stage_env():
    yield lambda v: os.environ.__setitem__("MVB_STAGE_AGENTS", str(v))
    os.environ.pop("MVB_STAGE_AGENTS", None)


# (agents, subcultures, neighbourhoods, staged agents or None = what the GPU's shared memory holds, seed)
@pytest.mark.parametrize("n,s,h,stage,seed", [
    (3000, 3, 4, None, 1),            # everything staged
    (40000, 4, 10, 8192, 2),          # most agents cold
    (40000, 2, 300, None, 3),         # many small neighbourhoods (regions padded to 64)
    (200000, 4, 10, 64000, 4),        # hubs + partial hot set
    (5000, 1, 1, 64, 5),              # smallest staged set: more hubs than fit
    (1_000_000, 4, 10, None, 6),      # BASELINE shape: a few percent cold at the real capacity
])
def test_device_layout_equals_host_layout(stage_env, n, s, h, stage, seed):
    if stage:
        stage_env(stage)
    pop = mb.synth_population(n, s, h, 5, 10, seed=seed)
    tab = mb.synth_tables(s, h, n, seed=seed)
    with mb.Simulation(pop, tab, mb.make_params()) as sim:
        compare_with_host(sim, pop)


def test_batch_with_shared_graph_layouts(stage_env):
    pops = [mb.synth_population(20000, 4, 10, 5, 10, seed=10), mb.synth_population(9000, 3, 6, 4, 6, seed=11)]
    tabs = [mb.synth_tables(4, 10, 20000, seed=10), mb.synth_tables(3, 6, 9000, seed=11)]
    pop3 = pops[0].copy()
    for k in ("social_rowptr", "social_col", "neighbour_rowptr", "neighbour_col", "neighbourhood"):
        pop3.arrays[k] = pops[0][k]
    pops.append(pop3); tabs.append(mb.synth_tables(4, 10, 20000, seed=12))
    with mb.Simulation(pops, tabs, mb.make_params()) as sim:
        for r in range(3):
            compare_with_host(sim, pops[r], r)


@pytest.mark.parametrize("stage", [64, 1024, 6400])
def test_partly_staged_small_population_equals_oracle(stage_env, stage):
    """Cold lists (targets gathered from L2 instead of shared memory), hubs that are not staged: every day, full state."""
    stage_env(stage)
    n = 12000
    pop = mb.synth_population(n, 4, 6, 5, 10, seed=stage)
    tab = mb.synth_tables(4, 6, n, seed=stage)
    w = mb.weather_pattern(30, seed=stage)
    g = mb.Simulation(pop, tab, mb.make_params())
    compare_with_host(g, pop)
    o = OracleSimulation(pop, tab, mb.make_params())
    prev = w[0]
    for day in range(1, 30):
        if day % 7 >= 5:
            continue
        g.step(w[day], prev != w[day]); o.step(w[day], prev != w[day])
        prev = w[day]
        assert_state_equal(g.get_state(0), o.get_state(), f"day {day}")
    g.close()
