"""BASELINE.json configs[0]: the reference's shipped configuration with the three documented fixes
(examples/c1/make_c1.py).  CPU part: the committed values are the shipped ones; the hosts read what
`c1.materialise` writes.  GPU part: tests/test_gpu_c1.py."""
import os
import subprocess

import numpy as np
import pytest

from motivate_b200 import c1
from motivate_b200.scenario import Parameters, Scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "motivate_b200", "motivate_b200_run")


def test_c1_round_trips_through_the_reference_schema(tmp_path):
    pp, sp = c1.materialise(str(tmp_path))
    prm, sc = Parameters.from_file(pp), Scenario.from_file(sp)
    assert (prm.number_of_people, prm.number_of_simulations, prm.total_years) == (111166, 8, 5)
    assert (prm.number_of_social_network_links, prm.number_of_neighbour_links, prm.days_in_habit_average) == (5, 10, 10)
    assert len(prm.distributions) == 3 and abs(sum(d[2] for d in prm.distributions) - 1.0) < 1e-4
    assert len(sc.neighbourhood_ids) == 20 and len(sc.subculture_ids) == 3
    assert sc.number_of_cars == 64926 and sc.number_of_bikes == 55763
    assert sc.intervention.day == 365 and len(sc.intervention.neighbourhood_changes) == 5
    assert all(c.id in sc.neighbourhood_ids for c in sc.intervention.neighbourhood_changes)
    # Subculture C keeps the exact Cycle = Walk tie of the shipped file (scenario.yaml:280-285): the canonical
    # tie-break (SURVEY.md §8 a-T) is what the parity test exercises on it
    assert sc.tables.desirability[2, 2] == sc.tables.desirability[2, 3] == np.float32(0.9)
    # the C++ host reads the same files to the same bits
    out = subprocess.run([BIN, "--dump", "--dir", str(tmp_path)], capture_output=True, text=True, check=True).stdout
    nbs = [l.split("|") for l in out.splitlines() if l.startswith("neighbourhood|")]
    assert [n[1] for n in nbs] == sc.neighbourhood_ids
    assert np.array_equal(np.array([[float.fromhex(x) for x in n[2].split()] for n in nbs], np.float32), sc.tables.supportiveness)
    assert "parameters 5 111166 8" in out and "intervention 365 0 0 5 0" in out


@pytest.mark.skipif(not os.path.exists("/root/reference/config/scenario.yaml"), reason="reference tree not present")
def test_c1_is_the_shipped_configuration_plus_the_three_fixes(tmp_path):
    ref = Scenario.from_file("/root/reference/config/scenario.yaml")
    refp = Parameters.from_file("/root/reference/config/parameters.yaml")
    pp, sp = c1.materialise(str(tmp_path))
    prm, sc = Parameters.from_file(pp), Scenario.from_file(sp)
    for k in ("supportiveness", "capacity", "desirability"):
        assert np.array_equal(getattr(sc.tables, k), getattr(ref.tables, k)), k
    assert sc.neighbourhood_ids == ref.neighbourhood_ids and sc.subculture_ids == ref.subculture_ids
    assert (sc.number_of_cars, sc.number_of_bikes) == (ref.number_of_cars, ref.number_of_bikes)
    for a, b in zip(sc.intervention.neighbourhood_changes, ref.intervention.neighbourhood_changes):
        assert a.increase_in_supportiveness == b.increase_in_supportiveness and a.increase_in_capacity == b.increase_in_capacity
        assert a.id == b.id[:4] + "0" + b.id[4:]                                     # fix 3
    assert ref.intervention.day == 0 and sc.intervention.day == 365                  # fix 2
    assert refp.distributions == [] and len(prm.distributions) == 3                  # fix 1
    for k in ("days_in_habit_average", "neighbourhood_connectivity", "number_of_neighbour_links", "number_of_people",
              "number_of_simulations", "number_of_social_network_links", "social_connectivity", "subculture_connectivity",
              "total_years"):
        assert getattr(prm, k) == getattr(refp, k), k


def test_tie_free_variant_has_no_equal_table_entries(tmp_path):
    _, sp = c1.materialise(str(tmp_path), tie_free=True)
    sc = Scenario.from_file(sp)
    for t in (sc.tables.desirability, sc.tables.supportiveness):
        assert all(len(set(r.tolist())) == 4 for r in t)
