"""oracle/reference_shaped.cpp — the reference's own shape (heap objects pointing at their friends, a node-based map
per intermediate result, union / intersection folds, group maps, stable sorts, max_by) compiled, so that the dense
oracle can be checked against it at sizes the Python literal model cannot reach, and so that bench.py has the cost
shape of the reference's CPU path to time.  Third restatement: hand-derived micro cases, literal model, dense oracle
must all agree with it bit for bit."""
import numpy as np
import pytest

import motivate_b200 as mb
from common import assert_state_equal, random_population, random_tables
from oracle.binding import OracleSimulation
from test_oracle_micro import CASES, check_case
from test_oracle_vs_literal import literal_states


class RefShaped(OracleSimulation):
    def __init__(self, pop, tab, params):
        super().__init__(pop, tab, params, shape="reference")


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.__name__)
def test_micro_cases(case):
    check_case(case, RefShaped)


@pytest.mark.parametrize("seed,per_agent", [(0, False), (1, True), (2, False), (3, True)])
def test_matches_literal_model_every_day(seed, per_agent):
    n, S, H, n_days = 90, 3, 4, 12
    pop = random_population(n, S, H, 100 + seed, per_agent=per_agent)
    tab = random_tables(S, H, n, 100 + seed)
    prm = mb.make_params(days_in_habit_average=3 + seed)
    weather = mb.weather_pattern(n_days, seed=seed, p_good_given_good=0.5, p_good_given_bad=0.5)
    days, rows, snaps = literal_states(pop, tab, prm, weather, n_days)
    o = RefShaped(pop, tab, prm)
    assert np.array_equal(o.stats_row(), rows[0])
    prev, r = weather[0], 1
    for day in range(1, n_days):
        if day % 7 >= 5:
            continue
        o.step(weather[day], prev != weather[day])
        prev = weather[day]
        st, ref = o.get_state(), snaps[day]
        ref["hist"] = st["hist"]
        assert_state_equal(st, ref, f"seed {seed} day {day}")
        assert np.array_equal(o.stats_row(), rows[r])
        r += 1


@pytest.mark.parametrize("seed,per_agent", [(5, False), (6, True)])
def test_dense_oracle_equals_reference_shape_random_5k(seed, per_agent):
    n, S, H, n_days = 5000, 5, 7, 30
    pop = random_population(n, S, H, seed, per_agent=per_agent, max_deg=9)
    tab = random_tables(S, H, n, seed)
    prm = mb.make_params(days_in_habit_average=7)
    weather = mb.weather_pattern(n_days, seed=seed, p_good_given_good=0.6, p_good_given_bad=0.5)
    a, b = OracleSimulation(pop, tab, prm), RefShaped(pop, tab, prm)
    prev = weather[0]
    for day in range(1, n_days):
        if day % 7 >= 5:
            continue
        for o in (a, b):
            o.step(weather[day], prev != weather[day])
        prev = weather[day]
        assert_state_equal(a.get_state(), b.get_state(), f"day {day}")
        assert np.array_equal(a.stats_row(), b.stats_row())


def test_dense_oracle_equals_reference_shape_synthetic_100k_with_intervention():
    """BASELINE-shaped population (BA links 5 + 10, hubs, duplicates), congested Car capacity, an intervention on a
    Saturday: rows of the whole run and the final state."""
    n, S, H, n_days = 100_000, 4, 10, 40
    pop = mb.synth_population(n, S, H, 5, 10, seed=77)
    tab = mb.synth_tables(S, H, n, seed=77)
    tab2 = mb.TablesData(np.clip(tab.supportiveness + np.float32(0.1), 0, 1).astype(np.float32),
                         (tab.capacity // 2).astype(np.uint32), tab.desirability.copy())
    rng = np.random.default_rng(5)
    car2 = rng.integers(0, 2, n).astype(np.uint8)
    iv = mb.InterventionData(day=20, tables=tab2, owns_car=car2, owns_bike=None)       # 20 % 7 == 6: a weekend day
    weather = mb.weather_pattern(n_days, seed=3)
    prm = mb.make_params()
    a, b = OracleSimulation(pop, tab, prm), RefShaped(pop, tab, prm)
    da, ra = a.run(weather, n_days, intervention=iv)
    db, rb = b.run(weather, n_days, intervention=iv)
    assert np.array_equal(da, db) and np.array_equal(ra, rb)
    assert_state_equal(a.get_state(), b.get_state(), "final state")
