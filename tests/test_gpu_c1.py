"""BASELINE.json configs[0] on the GPU: the reference's shipped configuration (S=3, H=20, 111 166 agents, the real
tables incl. Subculture C's exact Cycle = Walk tie, intervention on day 365) with the three documented fixes
(examples/c1/make_c1.py), generated with the host, two years of the reference's own rain sequence (HC-128),
against the CPU oracle EVERY weekday: full state (modes, norms, habit bits, histogram, congestion bits), the
statistics rows, and the CSV bytes."""
import os

import numpy as np
import pytest

import motivate_b200 as mb
from common import assert_state_equal
from motivate_b200 import c1, simulation
from motivate_b200.scenario import Parameters, write_csv
from motivate_b200.weather import make_pattern
from oracle.binding import OracleSimulation

pytestmark = pytest.mark.gpu


def test_c1_shipped_configuration_every_weekday(tmp_path):
    pp, sp = c1.materialise(str(tmp_path))
    prm = Parameters.from_file(pp)
    years = 2
    prep = simulation.prepare("1", True, str(tmp_path / "agents_1.yaml"), sp, prm.number_of_people, prm.social_connectivity,
                              prm.subculture_connectivity, prm.neighbourhood_connectivity, prm.number_of_neighbour_links,
                              prm.days_in_habit_average, prm.distributions,
                              number_of_social_network_links=prm.number_of_social_network_links)
    pop, tab, iv = prep.population, prep.tables, prep.intervention
    assert pop.n_agents == 111166 and tab.H == 20 and tab.S == 3 and iv.day == 365
    assert int(pop["owns_car"].sum()) == 64926 and int(pop["owns_bike"].sum()) == 64926      # agent_generation.rs:160 (sic)
    n_days = 365 * years
    w = make_pattern(n_days)                              # Weather::make_pattern with StdRng::from_seed([0; 32])
    # 1. the whole loop in one call (what simulation::run does), CSV written by the host
    g = mb.Simulation(pop, tab, prep.params)
    days, rows = g.run(w, n_days, interventions=iv)
    final_gpu = g.get_state(0)
    g.close()
    o = OracleSimulation(pop, tab, prep.params)
    d2, r2 = o.run(w, n_days, intervention=iv)
    assert np.array_equal(days, d2) and len(days) == mb.run_rows(n_days) == 522
    assert np.array_equal(rows[0], r2)
    assert_state_equal(final_gpu, o.get_state(), "final state after 2 years")
    write_csv(str(tmp_path / "gpu.csv"), prep.scenario, days, w, rows[0])
    write_csv(str(tmp_path / "cpu.csv"), prep.scenario, d2, w, r2)
    assert open(tmp_path / "gpu.csv", "rb").read() == open(tmp_path / "cpu.csv", "rb").read()
    k365 = int(np.searchsorted(days, 365))
    assert not np.array_equal(r2[k365 - 1, 7 + tab.S:], r2[k365 + 9, 7 + tab.S:])           # the intervention moved something
    # 2. day by day: full state after every weekday, across the intervention
    g = mb.Simulation(pop, tab, prep.params)
    o = OracleSimulation(pop, tab, prep.params)
    prev = w[0]
    for day in range(1, n_days):
        if day == iv.day:                                                    # before the weekday gate (simulation.rs:103-108)
            g.set_tables(0, iv.tables); o.set_tables(iv.tables)
            g.set_ownership(0, iv.owns_car, iv.owns_bike); o.set_ownership(iv.owns_car, iv.owns_bike)
        if day % 7 >= 5:
            continue
        g.step(w[day], prev != w[day]); o.step(w[day], prev != w[day])
        prev = w[day]
        assert_state_equal(g.get_state(0), o.get_state(), f"day {day}")
        assert np.array_equal(g.stats_row(0), o.stats_row()), f"row of day {day}"
    g.close()


def test_c1_eight_replicas_one_batch(tmp_path):
    """number_of_simulations = 8 replicas of the shipped configuration stepped together (main.rs:89-121), at a tenth of
    the population so that the oracle side stays short: rows of every replica equal its own oracle run."""
    pp, sp = c1.materialise(str(tmp_path), people=11117)
    prm = Parameters.from_file(pp)
    preps = [simulation.prepare(str(i), True, "", sp, prm.number_of_people, prm.social_connectivity, prm.subculture_connectivity,
                                prm.neighbourhood_connectivity, prm.number_of_neighbour_links, prm.days_in_habit_average,
                                prm.distributions, number_of_social_network_links=prm.number_of_social_network_links)
             for i in range(1, prm.number_of_simulations + 1)]
    w = make_pattern(365 * 2)
    days, rows = simulation.run_many(preps, [str(i) for i in range(1, 9)], 2, w, output_dir=str(tmp_path / "output"))
    for i, p in enumerate(preps):
        o = OracleSimulation(p.population, p.tables, p.params)
        d2, r2 = o.run(w, 730, intervention=p.intervention)
        assert np.array_equal(rows[i], r2), f"replica {i + 1}"
    assert len(open(tmp_path / "output" / "output_8.csv").read().splitlines()) == 1 + 522
