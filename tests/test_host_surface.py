"""The host-side mirror of the reference's configuration surface: parameters.yaml / scenario.yaml schema, intervention
resolution, network and agent files, CSV format (src/main.rs:185-228, scenario.rs, intervention.rs,
simulation.rs:194-268, 304-464).  Uses examples/config (same schema as the reference's config/, written for this repo)."""
import os

import numpy as np
import pytest

import motivate_b200 as mb
from motivate_b200 import files, simulation
from motivate_b200.scenario import (Parameters, Scenario, ScenarioError, apply_table_changes, generate_csv_header,
                                    intervene, write_csv)
from oracle.binding import OracleSimulation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "examples", "config")


def test_parameters_and_scenario_schema():
    p = Parameters.from_file(os.path.join(CFG, "parameters.yaml"))
    assert (p.total_years, p.number_of_people, p.number_of_simulations) == (2, 20000, 2)
    assert len(p.distributions) == 3 and p.days_in_habit_average == 10
    sc = Scenario.from_file(os.path.join(CFG, "scenario.yaml"))
    assert sc.neighbourhood_ids == ["north", "east", "south", "west"] and len(sc.subculture_ids) == 3
    # f32 fields go decimal -> f64 -> f32 like serde_yaml + `as f32`
    assert sc.tables.supportiveness[0, 0] == np.float32(0.4803388204468297)
    assert sc.tables.capacity[1].tolist() == [2084, 870, 150000, 150000]       # Car, PublicTransport, Cycle, Walk
    assert sc.tables.desirability[2].tolist() == [np.float32(0.4), np.float32(0.5), np.float32(0.9), np.float32(0.9)]
    iv = sc.intervention
    assert iv.day == 365 and iv.change_in_number_of_bikes == 500 and iv.change_in_number_of_cars == -300
    assert iv.neighbourhood_changes[1].increase_in_supportiveness == {}            # #[serde(default)]


@pytest.mark.skipif(not os.path.exists("/root/reference/config/scenario.yaml"), reason="reference tree not present")
def test_reads_the_reference_shipped_config():
    sc = Scenario.from_file("/root/reference/config/scenario.yaml")
    assert len(sc.neighbourhood_ids) == 20 and len(sc.subculture_ids) == 3 and sc.number_of_cars == 64926
    assert sc.intervention.day == 0                                  # never fires (SURVEY.md §5 defect 2)
    p = Parameters.from_file("/root/reference/config/parameters.yaml")
    assert p.number_of_people == 111166 and p.distributions == []
    pop = mb.synth_population(500, 3, 20, 3, 3, seed=0)
    with pytest.raises(ScenarioError, match="neighbourhood in your intervention was not found"):
        intervene(sc, pop)                                           # ids "E0500590" vs "E05000590" (defect 3)


def test_scenario_rejects_missing_modes_and_unknown_ids():
    sc = Scenario.from_file(os.path.join(CFG, "scenario.yaml"))
    import yaml
    d = yaml.safe_load(open(os.path.join(CFG, "scenario.yaml")))
    del d["subcultures"][0]["desirability"]["Walk"]
    with pytest.raises(ScenarioError, match="all four transport modes"):
        Scenario.from_dict(d)
    d = yaml.safe_load(open(os.path.join(CFG, "scenario.yaml")))
    d["intervention"]["subculture_changes"][0]["id"] = "nobody"
    pop = mb.synth_population(200, 3, 4, 3, 3, seed=0)
    with pytest.raises(ScenarioError, match="subculture in your intervention was not found"):
        intervene(Scenario.from_dict(d), pop)


def test_intervene_tables_and_ownership():
    sc = Scenario.from_file(os.path.join(CFG, "scenario.yaml"))
    pop = mb.synth_population(5000, 3, 4, 3, 3, seed=1)
    iv = intervene(sc, pop, seed=3)
    t0, t1 = sc.tables, iv.tables
    assert t1.supportiveness[0, 0] == np.float32(t0.supportiveness[0, 0]) + np.float32(-0.1)      # f32 add (:320-325)
    assert t1.supportiveness[0, 3] == np.float32(t0.supportiveness[0, 3]) + np.float32(0.1)
    assert t1.capacity[0, 0] == t0.capacity[0, 0] - 200 and t1.capacity[1, 1] == t0.capacity[1, 1] + 400
    assert np.array_equal(t1.capacity[2:], t0.capacity[2:]) and np.array_equal(t1.supportiveness[1:], t0.supportiveness[1:])
    assert t1.desirability[1, 0] == np.float32(t0.desirability[1, 0]) + np.float32(-0.1)
    # +500 bikes among non-owners, -300 cars among owners, uniformly sampled (:383-463)
    assert int(iv.owns_bike.sum()) == int(pop["owns_bike"].sum()) + 500 and np.all(iv.owns_bike >= pop["owns_bike"])
    assert int(iv.owns_car.sum()) == int(pop["owns_car"].sum()) - 300 and np.all(iv.owns_car <= pop["owns_car"])
    # capacity clamps to [0, u32::MAX] (:340-347)
    t2 = apply_table_changes(t0, [(0, {}, {0: -10**9}), (1, {}, {1: 2**40})], [])
    assert t2.capacity[0, 0] == 0 and t2.capacity[1, 1] == 0xFFFFFFFF
    # more cars taken away than exist: sample_slice_ref would panic
    sc.intervention.change_in_number_of_cars = -10**6
    with pytest.raises(ScenarioError, match="exceeds"):
        intervene(sc, pop)


def test_network_and_agent_files_round_trip(tmp_path):
    sc = Scenario.from_file(os.path.join(CFG, "scenario.yaml"))
    pop = mb.synth_population(300, 3, 4, 3, 3, seed=2)
    prm = mb.make_params()
    net, ag = str(tmp_path / "1.yaml"), str(tmp_path / "agents_1.yaml")
    files.write_network(net, pop["social_rowptr"], pop["social_col"])
    files.write_agents(ag, pop, sc, prm)
    txt = open(ag).read()
    assert "subculture_id: Subculture" in txt and "commute_length: " in txt and "owns_car: " in txt   # serde field names
    rp, col = files.read_network(net, 300)
    assert np.array_equal(rp, pop["social_rowptr"]) and np.array_equal(col, pop["social_col"])      # order + duplicates kept
    back = files.load_population(ag, net, sc, number_of_neighbour_links=3, seed=5)
    for k in ("subculture", "neighbourhood", "commute", "owns_car", "owns_bike", "weather_sensitivity", "habit",
              "current_mode", "norm", "social_rowptr", "social_col"):
        assert np.array_equal(back[k], pop[k]), k
    assert np.all(back["average_weight"] == np.float32(prm.average_weight)) and np.all(back["consistency"] == 1.0)
    # regenerated neighbour networks stay inside the neighbourhood (simulation.rs:165-173)
    src = np.repeat(np.arange(300), np.diff(back["neighbour_rowptr"]).astype(np.int64))
    assert np.array_equal(back["neighbourhood"][src], back["neighbourhood"][back["neighbour_col"]])
    with pytest.raises(ScenarioError):
        files.read_network(net, 100)                                # ids beyond the population: agents[id] panics


def test_csv_format_matches_reference(tmp_path):
    """Header and rows as simulation.rs:207-211, 254-267; Rain is the weather of the row's own day."""
    sc = Scenario.from_file(os.path.join(CFG, "scenario.yaml"))
    assert generate_csv_header(sc) == ("Day,Rain,ActiveMode,ActiveNorm,ActiveModeCounterToInactiveNorm,"
                                       "InactiveModeCounterToActiveNorm,LocalCommute,CityCommute,DistantCommute,"
                                       "Subculture A,Subculture B,Subculture C,north,east,south,west\n")
    prep = simulation.prepare("7", True, str(tmp_path / "agents.yaml"), os.path.join(CFG, "scenario.yaml"), 12500, 0.7, 0.5,
                              0.3, 3, 10, [], number_of_social_network_links=3, seed=7)
    w = mb.weather_pattern(20, seed=1)
    o = OracleSimulation(prep.population, prep.tables, prep.params)           # the checker, not the product path
    days, rows = o.run(w, 20)
    out = str(tmp_path / "output_7.csv")
    write_csv(out, prep.scenario, days, w, rows)
    lines = open(out).read().splitlines()
    assert len(lines) == 1 + len(days) and lines[0].count(",") == 15
    first = [int(x) for x in lines[1].split(",")]
    assert first[0] == 0 and first[1] == int(w[0]) and first[2] == first[3] and first[4] == first[5] == 0
    assert [int(l.split(",")[0]) for l in lines[1:]] == [d for d in range(20) if d == 0 or d % 7 < 5]
    assert os.path.exists(str(tmp_path / "agents.yaml"))                       # --generate saved the agents (:95)


@pytest.mark.gpu
def test_simulation_run_end_to_end_on_gpu(tmp_path):
    """simulation::run with generated agents, then again from the saved files: identical CSVs; rows equal the oracle's."""
    import yaml
    d = yaml.safe_load(open(os.path.join(CFG, "scenario.yaml")))
    d["number_of_cars"], d["number_of_bikes"] = 1200, 1000                     # scaled to 2 000 agents
    d["intervention"]["change_in_number_of_bikes"], d["intervention"]["change_in_number_of_cars"] = 100, -80
    for nb in d["neighbourhoods"]:
        nb["capacity"]["Car"] //= 10; nb["capacity"]["PublicTransport"] //= 10
    scen = str(tmp_path / "scenario.yaml")
    yaml.safe_dump(d, open(scen, "w"))
    n, years = 2000, 2
    w = mb.weather_pattern(365 * years, seed=4)
    pop0 = mb.synth_population(n, 3, 4, 3, 3, seed=9)
    files.write_network(str(tmp_path / "net.yaml"), pop0["social_rowptr"], pop0["social_col"])
    args = dict(total_years=years, number_of_people=n, social_connectivity=0.7, subculture_connectivity=0.5,
                neighbourhood_connectivity=0.3, number_of_neighbour_links=3, days_in_habit_average=10, distributions=[],
                weather_pattern=w, network=str(tmp_path / "net.yaml"), seed=11)
    d1, r1 = simulation.run("a", True, str(tmp_path / "agents.yaml"), scen, output_dir=str(tmp_path / "out1"), **args)
    d2, r2 = simulation.run("a", False, str(tmp_path / "agents.yaml"), scen, output_dir=str(tmp_path / "out2"), **args)
    assert open(tmp_path / "out1" / "output_a.csv").read() == open(tmp_path / "out2" / "output_a.csv").read()
    assert len(d1) == mb.run_rows(730) == 522
    prep = simulation.prepare("a", False, str(tmp_path / "agents.yaml"), scen, n, 0.7, 0.5, 0.3, 3, 10, [],
                              network=str(tmp_path / "net.yaml"), seed=11)
    o = OracleSimulation(prep.population, prep.tables, prep.params)
    do, ro = o.run(w, 730, intervention=prep.intervention)
    assert np.array_equal(r1, ro) and np.array_equal(d1, do)
    assert not np.array_equal(ro[259], ro[275])                                # the day-365 intervention moved something


def test_default_seed_is_the_same_in_every_process(tmp_path):
    """prepare() without a seed: the neighbour networks (never saved, rebuilt every run: simulation.rs:165-189) and the
    resolved intervention must not depend on the process (str hashes are salted per interpreter)."""
    import subprocess
    import sys
    code = (
        "import sys, zlib, numpy as np; sys.path.insert(0, %r)\n"
        "from motivate_b200 import simulation\n"
        "p = simulation.prepare(sys.argv[1], True, '', %r, 12500, 0.7, 0.5, 0.3, 3, 10, [], number_of_social_network_links=3)\n"
        "h = zlib.crc32(p.population['neighbour_col'].tobytes())\n"
        "h = zlib.crc32(p.population['social_col'].tobytes(), h)\n"
        "h = zlib.crc32(p.intervention.owns_car.tobytes(), h); h = zlib.crc32(p.intervention.owns_bike.tobytes(), h)\n"
        "print(h)\n" % (ROOT, os.path.join(CFG, "scenario.yaml")))
    outs = {}
    for ident in ("7", "run-a"):
        outs[ident] = [subprocess.run([sys.executable, "-c", code, ident], check=True, capture_output=True, text=True,
                                      env={**os.environ, "PYTHONHASHSEED": str(hs)}).stdout.strip() for hs in (1, 2)]
        assert outs[ident][0] == outs[ident][1], ident
    assert outs["7"][0] != outs["run-a"][0]                          # different ids still give different populations


def test_run_many_refuses_mixed_uniform_weights(tmp_path):
    prep = simulation.prepare("1", True, "", os.path.join(CFG, "scenario.yaml"), 12500, 0.7, 0.5, 0.3, 3, 10, [],
                              number_of_social_network_links=3, seed=1)
    other = simulation.prepare("2", True, "", os.path.join(CFG, "scenario.yaml"), 12500, 0.9, 0.5, 0.3, 3, 10, [],
                               number_of_social_network_links=3, seed=2)
    with pytest.raises(ValueError, match="same uniform weights"):
        simulation.run_many([prep, other], ["1", "2"], 1, mb.weather_pattern(365, seed=0), output_dir=str(tmp_path))
