"""Offline campaign (not collected by pytest): the C oracle against the literal, HashMap-structured model on many
random populations — every day, full state and statistics row, bit for bit.
    python tests/fuzz_oracle.py [n_seeds=2000] [first_seed=5000]
    python tests/fuzz_oracle.py --reference-shaped [n_seeds=300] [first_seed=20000]
Round 1: 2 000 seeds (5-120 agents, 1-6 subcultures, 1-8 neighbourhoods, uniform and per-agent weights, link lists
with duplicates and empty lists, all ownership / commute combinations, 8-15 days): 0 mismatches, 85 s.
With --reference-shaped the dense oracle runs against oracle/reference_shaped.cpp instead, on larger populations
(200-3 000 agents, up to 11 neighbourhoods, 8-21 days; rows of the run + final state): 300 seeds, 0 mismatches, 63 s."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import motivate_b200 as mb  # noqa: E402
from common import assert_state_equal, random_population, random_tables  # noqa: E402
from oracle.binding import OracleSimulation  # noqa: E402
from test_oracle_vs_literal import literal_states  # noqa: E402


def one(seed):
    rng = np.random.default_rng(seed)
    n, S, H = int(rng.integers(5, 120)), int(rng.integers(1, 7)), int(rng.integers(1, 9))
    n_days = int(rng.integers(8, 16))
    pop = random_population(n, S, H, seed, per_agent=bool(seed & 1), max_deg=int(rng.integers(1, 14)),
                            distant_share=float(rng.uniform(0, 0.5)))
    tab = random_tables(S, H, n, seed)
    prm = mb.make_params(days_in_habit_average=int(rng.integers(1, 40)))
    weather = mb.weather_pattern(n_days, seed=seed, p_good_given_good=float(rng.uniform(0.2, 0.9)),
                                 p_good_given_bad=float(rng.uniform(0.2, 0.9)))
    days, rows, snaps = literal_states(pop, tab, prm, weather, n_days)
    o = OracleSimulation(pop, tab, prm)
    assert np.array_equal(o.stats_row(), rows[0]), f"seed {seed}: day-0 row"
    prev, r = weather[0], 1
    for day in range(1, n_days):
        if day % 7 >= 5:
            continue
        o.step(weather[day], prev != weather[day])
        prev = weather[day]
        st, ref = o.get_state(), snaps[day]
        ref["hist"] = st["hist"]                          # the literal model keeps no histogram; it is checked via the rows
        assert_state_equal(st, ref, f"seed {seed} day {day}")
        assert np.array_equal(o.stats_row(), rows[r]), f"seed {seed} day {day}: row"
        r += 1


def one_reference_shaped(seed):
    rng = np.random.default_rng(seed)
    n, S, H = int(rng.integers(200, 3000)), int(rng.integers(1, 7)), int(rng.integers(1, 12))
    n_days = int(rng.integers(8, 22))
    pop = random_population(n, S, H, seed, per_agent=bool(seed & 1), max_deg=int(rng.integers(1, 20)),
                            distant_share=float(rng.uniform(0, 0.5)))
    tab = random_tables(S, H, n, seed)
    prm = mb.make_params(days_in_habit_average=int(rng.integers(1, 40)))
    w = mb.weather_pattern(n_days, seed=seed, p_good_given_good=float(rng.uniform(0.2, 0.9)),
                           p_good_given_bad=float(rng.uniform(0.2, 0.9)))
    a, b = OracleSimulation(pop, tab, prm), OracleSimulation(pop, tab, prm, shape="reference")
    (da, ra), (db, rb) = a.run(w, n_days), b.run(w, n_days)
    assert np.array_equal(da, db) and np.array_equal(ra, rb), f"seed {seed}: rows"
    assert_state_equal(a.get_state(), b.get_state(), f"seed {seed}")


if __name__ == "__main__":
    argv = [x for x in sys.argv[1:] if x != "--reference-shaped"]
    ref = len(argv) != len(sys.argv) - 1
    n_seeds = int(argv[0]) if argv else (300 if ref else 2000)
    first = int(argv[1]) if len(argv) > 1 else (20000 if ref else 5000)
    t0, bad = time.time(), 0
    for seed in range(first, first + n_seeds):
        try:
            (one_reference_shaped if ref else one)(seed)
        except AssertionError as e:
            bad += 1
            print("MISMATCH", str(e)[:300])
    print(f"{n_seeds} seeds, {bad} mismatches, {time.time() - t0:.1f} s")
    sys.exit(1 if bad else 0)
