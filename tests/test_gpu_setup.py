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
    def set_env(v, layout=None):
        os.environ["MVB_STAGE_AGENTS"] = str(v)
        if layout:
            os.environ["MVB_LAYOUT"] = layout
    yield set_env
    for k in ("MVB_STAGE_AGENTS", "MVB_LAYOUT", "MVB_PUSH", "MVB_PUSH_CHUNK"):
        os.environ.pop(k, None)


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


@pytest.mark.parametrize("n,s,h,stage,seed", [(40000, 4, 10, 8192, 2), (200000, 4, 10, 64000, 4), (5000, 1, 1, 64, 5), (30000, 3, 40, 4096, 7)])
def test_device_layout_equals_host_layout_local_mode(stage_env, n, s, h, stage, seed):
    """LOCAL staging (layout.h): hub region + the prefix of the block's own neighbourhood."""
    stage_env(stage, "local")
    pop = mb.synth_population(n, s, h, 5, 10, seed=seed)
    tab = mb.synth_tables(s, h, n, seed=seed)
    with mb.Simulation(pop, tab, mb.make_params()) as sim:
        compare_with_host(sim, pop)


@pytest.mark.parametrize("layout", ["global", "local"])
@pytest.mark.parametrize("stage", [64, 1024, 6400])
def test_partly_staged_small_population_equals_oracle(stage_env, stage, layout):
    """Cold lists (targets gathered from L2 instead of shared memory), hubs that are not staged, both staging layouts:
    every day, full state."""
    stage_env(stage, layout)
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


@pytest.mark.parametrize("layout", ["global", "local"])
@pytest.mark.parametrize("stage,chunk", [(64, 96), (1024, 4096), (6400, 24576)])
def test_push_mode_small_population_equals_oracle(stage_env, stage, chunk, layout):
    """PUSH mode (layout.h): the cold links of the SELL agents are counted by push_kernel from chunk-wise target-sorted
    lists instead of being gathered by their source's lane.  Forced on a small population with a small staged set and
    small chunks (many chunks per block, a ragged last chunk); every day, full state, an intervention on the way."""
    stage_env(stage, layout)
    os.environ["MVB_PUSH"] = "1"
    os.environ["MVB_PUSH_CHUNK"] = str(chunk)
    n = 12000
    pop = mb.synth_population(n, 4, 6, 5, 10, seed=stage + 1)
    tab = mb.synth_tables(4, 6, n, seed=stage + 1)
    w = mb.weather_pattern(30, seed=stage)
    g = mb.Simulation(pop, tab, mb.make_params())
    compare_with_host(g, pop)
    o = OracleSimulation(pop, tab, mb.make_params())
    prev = w[0]
    for day in range(1, 30):
        if day == 12:
            car = (1 - pop["owns_car"]).astype(np.uint8)
            g.set_ownership(0, car, None); o.set_ownership(car, None)
        if day % 7 >= 5:
            continue
        g.step(w[day], prev != w[day]); o.step(w[day], prev != w[day])
        prev = w[day]
        assert_state_equal(g.get_state(0), o.get_state(), f"day {day}")
        assert np.array_equal(g.stats_row(0), o.stats_row())
    g.close()


def test_push_mode_with_per_agent_weights(stage_env):
    """The general kernel (per-agent weights, plain IEEE division) in its single-population instantiation: unstructured
    inputs (duplicate links, empty lists, exact ties), a 64-agent staged set, push mode, several chunks; every day."""
    from common import random_population, random_tables
    stage_env(64, "local")
    os.environ["MVB_PUSH"] = "1"
    os.environ["MVB_PUSH_CHUNK"] = "256"
    n, S, H, n_days = 3000, 5, 7, 25
    pop = random_population(n, S, H, 3, per_agent=True, max_deg=9)
    tab = random_tables(S, H, n, 3)
    w = mb.weather_pattern(n_days, seed=3, p_good_given_good=0.6, p_good_given_bad=0.5)
    prm = mb.make_params(days_in_habit_average=4)
    g = mb.Simulation(pop, tab, prm)
    o = OracleSimulation(pop, tab, prm)
    prev = w[0]
    for day in range(1, n_days):
        if day % 7 >= 5:
            continue
        g.step(w[day], prev != w[day]); o.step(w[day], prev != w[day])
        prev = w[day]
        assert_state_equal(g.get_state(0), o.get_state(), f"day {day}")
    g.close()


def test_push_mode_run_equals_pull_run_200k(stage_env):
    """The same 200k-agent population through the pull path and through the push path (one third of the agents staged,
    LOCAL layout, default chunk size): rows and full state identical after a 3-week run with mvb_run."""
    n = 200_000
    pop = mb.synth_population(n, 4, 10, 5, 10, seed=31)
    tab = mb.synth_tables(4, 10, n, seed=31)
    w = mb.weather_pattern(22, seed=31)
    out = []
    for push in ("0", "1"):
        stage_env(64000, "local")
        os.environ["MVB_PUSH"] = push
        with mb.Simulation(pop, tab, mb.make_params()) as sim:
            days, rows = sim.run(w)
            out.append((rows[0].copy(), sim.get_state(0)))
    assert np.array_equal(out[0][0], out[1][0])
    assert_state_equal(out[0][1], out[1][1], "push vs pull")


def test_device_generated_population_statistics_and_structure():
    """mvb_synth_create_device: the distributions of src/agent_generation.rs (exact owner counts, initial mode by
    ownership, habit one-hot, commute classes of the GMM), Barabási–Albert graphs with duplicates kept (symmetric,
    2·min links per agent on average, heavy tail), neighbour links inside the neighbourhood."""
    n, S, H = 300_000, 4, 10
    d = mb.DevicePopulation(n, S, H, 5, 10, seed=3)
    p = d.to_host()
    assert p.n_agents == n
    assert int(p["owns_car"].sum()) == round(64926 / 111166 * n) == int(p["owns_bike"].sum())
    assert abs(np.corrcoef(p["owns_car"], p["owns_bike"])[0, 1]) < 0.01             # independent samples
    assert np.bincount(p["subculture"], minlength=S).min() > 0.95 * n / S and p["subculture"].max() == S - 1
    assert np.bincount(p["neighbourhood"], minlength=H).min() > 0.95 * n / H and p["neighbourhood"].max() == H - 1
    assert 0.49 < p["weather_sensitivity"].mean() < 0.51 and p["weather_sensitivity"].max() < 1.0
    cm = np.bincount(p["commute"], minlength=3) / n                                  # |GMM| vs the 4241 m / 19457 m thresholds
    host = np.bincount(mb.synth_population(n, S, H, 2, 2, seed=1)["commute"], minlength=3) / n
    assert np.allclose(cm, host, atol=0.01), (cm, host)
    assert np.array_equal(p["habit"].argmax(1), p["current_mode"]) and np.all(p["habit"].sum(1) == 1.0)
    assert np.array_equal(p["norm"], p["current_mode"])
    both = (p["owns_car"] == 1) & (p["owns_bike"] == 1)
    share = np.bincount(p["current_mode"][both], minlength=4) / both.sum()
    assert np.allclose(share, [0.40, 0.15, 0.30, 0.15], atol=0.01), share            # agent_generation.rs:286-325
    none = (p["owns_car"] == 0) & (p["owns_bike"] == 0)
    assert np.allclose(np.bincount(p["current_mode"][none], minlength=4) / none.sum(), [0, 0.5, 0, 0.5], atol=0.01)
    for g, m in (("social", 5), ("neighbour", 10)):
        rp, col = p[g + "_rowptr"].astype(np.int64), p[g + "_col"].astype(np.int64)
        deg = np.diff(rp)
        assert rp[0] == 0 and rp[-1] == len(col) and col.max() < n
        assert abs(deg.mean() - 2 * m) < 0.02 and deg.min() >= m - 1 + (g == "neighbour") * 0
        assert deg.max() > 40 * m                                                     # preferential attachment: hubs
        src = np.repeat(np.arange(n), deg)
        # symmetric WITH multiplicity: the multiset of (a, b) equals the multiset of (b, a)
        fw = np.sort(src * n + col); bw = np.sort(col * n + src)
        assert np.array_equal(fw, bw)
        assert not np.any(src == col)                                                 # no self links
        if g == "neighbour":
            assert np.array_equal(p["neighbourhood"][src], p["neighbourhood"][col])  # simulation.rs:165-173
    # the degree tail follows the BA law P(k) ~ 2 m^2 / k^3: the share of agents with degree >= 4m is about (m / 4m)^2
    deg = np.diff(p["social_rowptr"].astype(np.int64))
    assert 0.04 < (deg >= 20).mean() < 0.09
    # different seeds give different populations, the same seed the same one
    q = mb.DevicePopulation(n, S, H, 5, 10, seed=3).to_host()
    assert np.array_equal(q["social_rowptr"], p["social_rowptr"]) and np.array_equal(q["current_mode"], p["current_mode"])
    rows = np.repeat(np.arange(n), np.diff(p["social_rowptr"].astype(np.int64)))
    canon = lambda c: c[np.lexsort((c, rows))]                                       # (the order inside a row is not fixed)
    assert np.array_equal(canon(q["social_col"]), canon(p["social_col"]))
    r = mb.DevicePopulation(n, S, H, 5, 10, seed=4).to_host()
    assert not np.array_equal(r["current_mode"], p["current_mode"])


def test_device_resident_population_runs_like_its_host_copy():
    """mvb_create* on a population that lives in device memory (no host copy made) against the oracle on its host copy,
    and against a handle created from the host copy: same layout, same results."""
    n = 50_000
    d = mb.DevicePopulation(n, 4, 10, 5, 10, seed=8)
    p = d.to_host()
    tab = mb.synth_tables(4, 10, n, seed=8)
    w = mb.weather_pattern(40, seed=8)
    with mb.Simulation(d, tab, mb.make_params()) as a, mb.Simulation(p, tab, mb.make_params()) as b:
        compare_with_host(a, p)
        da, ra = a.run(w)
        db, rb = b.run(w)
        assert np.array_equal(ra[0], rb[0])
        o = OracleSimulation(p, tab, mb.make_params())
        do, ro = o.run(w)
        assert np.array_equal(ra[0], ro)
        assert_state_equal(a.get_state(0), o.get_state(), "device-resident population")


def test_batch_mixing_staging_layouts_equals_oracle(stage_env):
    """One launch stepping a partly staged population (LOCAL layout: common set + own neighbourhood piece) next to populations
    that fit shared memory whole (GLOBAL layout): every replica equals its own oracle run."""
    stage_env(2048)
    pops = [mb.synth_population(15000, 4, 6, 5, 10, seed=31), mb.synth_population(900, 3, 4, 4, 6, seed=32),
            mb.synth_population(9000, 2, 12, 5, 10, seed=33)]
    tabs = [mb.synth_tables(4, 6, 15000, seed=31), mb.synth_tables(3, 4, 900, seed=32), mb.synth_tables(2, 12, 9000, seed=33)]
    w = mb.weather_pattern(45, seed=31)
    with mb.Simulation(pops, tabs, mb.make_params()) as g:
        for r in range(3):
            compare_with_host(g, pops[r], r)
        days, rows = g.run(w)
        for r in range(3):
            o = OracleSimulation(pops[r], tabs[r], mb.make_params())
            d2, r2 = o.run(w)
            assert np.array_equal(rows[r], r2), f"replica {r}"
            assert_state_equal(g.get_state(r), o.get_state(), f"replica {r}")


def test_many_neighbourhoods_local_layout_equals_oracle(stage_env):
    """LOCAL layout with more neighbourhoods than a launch would otherwise have parts (H = 40), hubs included."""
    stage_env(4096, "local")
    n = 30000
    pop = mb.synth_population(n, 3, 40, 5, 10, seed=41)
    tab = mb.synth_tables(3, 40, n, seed=41)
    w = mb.weather_pattern(25, seed=41)
    with mb.Simulation(pop, tab, mb.make_params()) as g:
        compare_with_host(g, pop)
        days, rows = g.run(w)
        o = OracleSimulation(pop, tab, mb.make_params())
        d2, r2 = o.run(w)
        assert np.array_equal(rows[0], r2)
        assert_state_equal(g.get_state(0), o.get_state(), "H = 40, LOCAL layout")


def test_more_replicas_than_one_launch_holds():
    """70 replicas = two launches per weekday (64 replica descriptors fit the kernel parameters), chained by programmatic
    dependent launch like the weekdays themselves: every replica equals its own oracle run."""
    R, n = 70, 1500
    pops = [mb.synth_population(n, 3, 4, 4, 6, seed=100 + r) for r in range(R)]
    tabs = [mb.synth_tables(3, 4, n, seed=100 + r) for r in range(R)]
    w = mb.weather_pattern(20, seed=9)
    with mb.Simulation(pops, tabs, mb.make_params()) as g:
        days, rows = g.run(w)
        ms, launches = g.last_run_stats()
        assert launches == 2 * (len(days) - 1)
        for r in (0, 1, 63, 64, 69):
            o = OracleSimulation(pops[r], tabs[r], mb.make_params())
            d2, r2 = o.run(w)
            assert np.array_equal(rows[r], r2), f"replica {r}"
            assert_state_equal(g.get_state(r), o.get_state(), f"replica {r}")
