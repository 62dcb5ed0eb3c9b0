"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bit-exact on everything: modes, norms, habit bits, histograms, congestion bits, statistics rows."""
import numpy as np
import pytest

import motivate_b200 as mb
from common import assert_state_equal, random_population, random_tables
from oracle.binding import OracleSimulation
from test_oracle_micro import CASES, PRM, check_case

pytestmark = pytest.mark.gpu


class GpuOne(mb.Simulation):
    """One-replica Simulation with the oracle's method shapes."""

    def __init__(self, pop, tab, prm):
        super().__init__(pop, tab, prm)

    def get_state(self):
        return super().get_state(0)

    def stats_row(self):
        return super().stats_row(0)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.__name__)
def test_micro_cases_gpu(case):
    check_case(case, GpuOne)


def lockstep(pop, tab, prm, weather, n_days, check_every=1, interventions=None):
    g = mb.Simulation(pop, tab, prm)
    o = OracleSimulation(pop, tab, prm)
    assert np.array_equal(g.stats_row(0), o.stats_row())
    prev = weather[0]
    k = 0
    for day in range(1, n_days):
        if interventions and day in interventions:
            iv = interventions[day]
            if iv.tables is not None:
                g.set_tables(0, iv.tables); o.set_tables(iv.tables)
            g.set_ownership(0, iv.owns_car, iv.owns_bike); o.set_ownership(iv.owns_car, iv.owns_bike)
        if day % 7 >= 5:
            continue
        g.step(weather[day], prev != weather[day]); o.step(weather[day], prev != weather[day])
        prev = weather[day]
        k += 1
        if k % check_every == 0 or day == n_days - 1:
            assert_state_equal(g.get_state(0), o.get_state(), f"day {day}")
            assert np.array_equal(g.stats_row(0), o.stats_row()), f"row day {day}"
    assert_state_equal(g.get_state(0), o.get_state(), "final")
    g.close()


@pytest.mark.parametrize("seed,per_agent", [(0, False), (1, True)])
def test_random_small_every_day(seed, per_agent):
    """Unstructured inputs: duplicate links, empty lists, sparse habits, quantised tables (exact ties), tight capacities."""
    n, S, H, n_days = 3000, 5, 7, 40
    pop = random_population(n, S, H, seed, per_agent=per_agent, max_deg=9)
    tab = random_tables(S, H, n, seed)
    w = mb.weather_pattern(n_days, seed=seed, p_good_given_good=0.6, p_good_given_bad=0.5)
    lockstep(pop, tab, mb.make_params(days_in_habit_average=3 + seed), w, n_days)


def test_synthetic_1k_two_years_every_day():
    pop = mb.synth_population(1000, 4, 10, 5, 10, seed=1)
    tab = mb.synth_tables(4, 10, 1000, seed=1)
    w = mb.weather_pattern(730, seed=1)
    lockstep(pop, tab, PRM, w, 730)


def test_synthetic_100k_run_matches_oracle_run():
    """mvb_run (whole loop on the device, with an intervention) against mvo_run: rows, days and final state."""
    n = 100_000
    pop = mb.synth_population(n, 4, 10, 5, 10, seed=2)
    tab = mb.synth_tables(4, 10, n, seed=2)
    n_days = 120
    w = mb.weather_pattern(n_days, seed=2)
    rng = np.random.default_rng(0)
    tab2 = tab.copy()
    tab2.supportiveness[::2, 2:] += np.float32(0.1); tab2.supportiveness[::2, 0] -= np.float32(0.1)
    tab2.capacity[::2, 0] = (tab2.capacity[::2, 0] * 0.5).astype(np.uint32)
    car2 = pop["owns_car"].copy(); car2[rng.choice(n, n // 20, replace=False)] ^= 1
    iv = mb.InterventionData(day=62, tables=tab2, owns_car=car2, owns_bike=None)      # 62 % 7 == 6: a Sunday
    g = mb.Simulation(pop, tab, PRM)
    days, rows = g.run(w, interventions=iv)
    o = OracleSimulation(pop, tab, PRM)
    d2, r2 = o.run(w, intervention=iv)
    assert np.array_equal(days, d2)
    assert np.array_equal(rows[0], r2)
    assert_state_equal(g.get_state(0), o.get_state(), "final")
    ms, launches = g.last_run_stats()
    assert launches == mb.run_rows(n_days) - 1 and ms > 0
    g.close()


def test_batch_replicas_equal_independent_runs():
    """R replicas stepped together == R independent runs (the rayon fan-out of main.rs:89-121); different sizes,
    different tables, one shared graph (scenario sweep)."""
    w = mb.weather_pattern(60, seed=4)
    pops = [mb.synth_population(5000, 4, 10, 5, 10, seed=10), mb.synth_population(3000, 3, 6, 4, 6, seed=11)]
    tabs = [mb.synth_tables(4, 10, 5000, seed=10), mb.synth_tables(3, 6, 3000, seed=11)]
    # third replica: same population object as the first (graph shared on the device), other tables + ownership
    pop3 = pops[0].copy()
    for k in ("social_rowptr", "social_col", "neighbour_rowptr", "neighbour_col"):
        pop3.arrays[k] = pops[0][k]
    pop3.arrays["owns_car"] = 1 - pops[0]["owns_car"]
    pops.append(pop3); tabs.append(mb.synth_tables(4, 10, 5000, seed=12))
    g = mb.Simulation(pops, tabs, PRM)
    days, rows = g.run(w)
    for r in range(3):
        o = OracleSimulation(pops[r], tabs[r], PRM)
        d2, r2 = o.run(w)
        assert np.array_equal(rows[r], r2), f"replica {r}"
        assert_state_equal(g.get_state(r), o.get_state(), f"replica {r}")
    g.close()


def test_run_in_pieces_equals_one_run():
    pop = mb.synth_population(4000, 4, 10, 5, 10, seed=5)
    tab = mb.synth_tables(4, 10, 4000, seed=5)
    w = mb.weather_pattern(100, seed=5)
    a = mb.Simulation(pop, tab, PRM)
    days, rows = a.run(w)
    b = mb.Simulation(pop, tab, PRM)
    b.run_async(w, 0, 37); b.run_async(w, 37, 38); b.run_async(w, 38, 100)
    d2, r2 = b.fetch_rows()
    assert np.array_equal(days, d2) and np.array_equal(rows[0], r2[0])
    assert_state_equal(a.get_state(0), b.get_state(0), "pieces")
    a.close(); b.close()


def test_full_size_properties_1m():
    """BASELINE config C2 size (1M agents, S=4, H=10, 5+10 link minimums): the oracle takes ~1 s/day here, so
    compare 3 weekdays exactly and check size-independent properties over a longer run."""
    n = 1_000_000
    pop = mb.synth_population(n, 4, 10, 5, 10, seed=0)
    tab = mb.synth_tables(4, 10, n, seed=0)
    w = mb.weather_pattern(60, seed=0)
    g = mb.Simulation(pop, tab, PRM)
    o = OracleSimulation(pop, tab, PRM)
    prev = w[0]
    for day in (1, 2, 3):
        g.step(w[day], prev != w[day]); o.step(w[day], prev != w[day]); prev = w[day]
    assert_state_equal(g.get_state(0), o.get_state(), "1M day 3")
    g.close()
    g = mb.Simulation(pop, tab, PRM)
    days, rows = g.run(w)
    st = g.get_state(0)
    rows = rows[0].astype(np.int64)
    res = np.bincount(pop["neighbourhood"], minlength=10)
    assert np.array_equal(st["hist"].sum(1), res)                       # histogram rows = residents per neighbourhood
    assert np.all(rows[:, 0] == rows[:, 4:7].sum(1))                    # active = sum over commute classes
    assert np.all(rows[:, 0] == rows[:, 7:11].sum(1)) and np.all(rows[:, 0] == rows[:, 11:].sum(1))
    assert np.all(rows[:, 2] <= rows[:, 0]) and np.all(rows[:, 3] <= rows[:, 1])
    assert np.all(rows[:, 0] - rows[:, 2] == rows[:, 1] - rows[:, 3])   # active&active_norm counted both ways
    assert rows[-1, 0] == np.sum(st["current_mode"] >= 2) and rows[-1, 1] == np.sum(st["norm"] >= 2)
    assert np.allclose(st["habit"].sum(1), 1.0, atol=1e-4)              # EMA of one-hot vectors stays a distribution
    assert not np.any((st["current_mode"] == 0) & (pop["owns_car"] == 0))
    assert not np.any((st["current_mode"] == 2) & (pop["owns_bike"] == 0))
    assert not np.any((st["current_mode"] >= 2) & (pop["commute"] == 2))  # Distant commute: Car / PT only
    g.close()


def test_more_replicas_than_one_launch_holds():
    """70 replicas in one handle: the day kernel takes 64 replica descriptors per launch (kernel parameters), so this
    steps with two launches per weekday; 6 subcultures exercise the shared-memory-atomic statistics path (S > 4)."""
    w = mb.weather_pattern(25, seed=8, p_good_given_good=0.6, p_good_given_bad=0.5)
    pops = [mb.synth_population(400 + 7 * r, 6, 3, 3, 4, seed=100 + r) for r in range(70)]
    tabs = [mb.synth_tables(6, 3, 400 + 7 * r, seed=100 + r) for r in range(70)]
    g = mb.Simulation(pops, tabs, PRM)
    days, rows = g.run(w)
    ms, launches = g.last_run_stats()
    assert launches == 2 * (mb.run_rows(25) - 1)
    for r in (0, 1, 33, 63, 64, 69):
        o = OracleSimulation(pops[r], tabs[r], PRM)
        d2, r2 = o.run(w)
        assert np.array_equal(rows[r], r2), f"replica {r}"
        assert_state_equal(g.get_state(r), o.get_state(), f"replica {r}")
    g.close()


def test_many_neighbourhoods_and_tiny_groups():
    """300 neighbourhoods over 3 000 agents: groups of ~10 agents (most slices are padding), neighbourhoods with a single
    resident or none (no neighbour links, empty lists), a 50-way table in shared memory."""
    n, S, H = 3000, 5, 300
    pop = mb.synth_population(n, S, H, 3, 2, seed=13)
    tab = random_tables(S, H, n, 13)
    w = mb.weather_pattern(30, seed=13, p_good_given_good=0.5, p_good_given_bad=0.5)
    lockstep(pop, tab, PRM, w, 30)


def test_population_above_shared_memory_capacity():
    """1.3M agents do not fit the staged vector (914k): a third of the agents are 'cold' and their modes are gathered
    from L2; 4 weekdays against the oracle, full state."""
    n = 1_300_000
    pop = mb.synth_population(n, 4, 10, 5, 10, seed=21)
    tab = mb.synth_tables(4, 10, n, seed=21)
    w = np.array([0, 1, 1, 0, 1], np.uint8)
    lockstep(pop, tab, PRM, w, 5, check_every=2)


@pytest.mark.gpu
def test_create_rejects_bad_indices():
    """What panics in the reference (indexing past the agent list, "Subculture not found") is an error return here; the
    link targets are range-checked on the device (count_hot_kernel), everything else by the host threads."""
    tab = mb.synth_tables(3, 4, 3000, seed=1)
    for what in ("social_col", "neighbour_col", "subculture", "neighbourhood", "current_mode"):
        pop = mb.synth_population(3000, 3, 4, seed=21)
        a = pop.arrays[what]
        a[len(a) // 2] = 3000 if what.endswith("col") else 200
        with pytest.raises(mb.MotivateError):
            mb.Simulation(pop, tab, mb.make_params())
    # and a good population next to a bad one: the error names the replica
    good, bad = mb.synth_population(3000, 3, 4, seed=22), mb.synth_population(3000, 3, 4, seed=23)
    bad.arrays["social_col"][7] = 1 << 31
    with pytest.raises(mb.MotivateError, match="replica 1"):
        mb.Simulation([good, bad], [tab, tab], mb.make_params())
    with mb.Simulation([good, good], [tab, tab], mb.make_params()) as sim:      # the handle machinery still works afterwards
        assert sim.n_replicas == 2


@pytest.mark.gpu
@pytest.mark.parametrize("group", ["2", "4"])
def test_both_claim_group_instantiations(group, monkeypatch):
    """The day kernel exists in two instantiations (work items claimed two or four at a time); the host picks by the
    number of slices a warp gets, which small test populations never make large.  MVB_CLAIM_GROUP pins the choice, so
    both are compared with the oracle here: several replicas, hubs, cold links, an intervention."""
    monkeypatch.setenv("MVB_CLAIM_GROUP", group)
    n, S, H, n_days = 40000, 4, 6, 20
    pops = [mb.synth_population(n, S, H, 5, 10, seed=40 + r) for r in range(3)]
    tabs = [mb.synth_tables(S, H, n, seed=40 + r) for r in range(3)]
    w = mb.weather_pattern(n_days, seed=21, p_good_given_good=0.6, p_good_given_bad=0.5)
    with mb.Simulation(pops, tabs, mb.make_params()) as sim:
        days, rows = sim.run(w, n_days)
        states = [sim.get_state(r) for r in range(3)]
    for r in range(3):
        o = OracleSimulation(pops[r], tabs[r], mb.make_params())
        odays, orows = o.run(w, n_days)
        assert np.array_equal(days, odays) and np.array_equal(rows[r], orows), (group, r)
        assert_state_equal(states[r], o.get_state(), f"group {group} replica {r}")


@pytest.mark.gpu
def test_gpu_equals_reference_shaped_model():
    """The CUDA path against oracle/reference_shaped.cpp directly (heap objects, a map per intermediate result: the
    reference's own shape), not only against the dense oracle: hubs, duplicates, per-agent weights, every day."""
    n, S, H, n_days = 4000, 4, 6, 24
    pop = random_population(n, S, H, 31, per_agent=True, max_deg=12)
    tab = random_tables(S, H, n, 31)
    prm = mb.make_params(days_in_habit_average=6)
    w = mb.weather_pattern(n_days, seed=31, p_good_given_good=0.6, p_good_given_bad=0.5)
    ref = OracleSimulation(pop, tab, prm, shape="reference")
    with mb.Simulation(pop, tab, prm) as sim:
        prev = w[0]
        for day in range(1, n_days):
            if day % 7 >= 5:
                continue
            sim.step(w[day], prev != w[day])
            ref.step(w[day], prev != w[day])
            prev = w[day]
            assert_state_equal(sim.get_state(), ref.get_state(), f"day {day}")
            assert np.array_equal(sim.stats_row(), ref.stats_row())
