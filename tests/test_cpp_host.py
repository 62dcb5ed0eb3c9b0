"""The C++ host layer (motivate_b200/csrc/host: yaml_lite + motivate_host + main, the reference's main.rs /
simulation::run surface over the C ABI), driven as the `motivate_b200_run` binary."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import yaml

import motivate_b200 as mb
from motivate_b200 import files
from motivate_b200.scenario import Parameters, Scenario, apply_table_changes
from oracle.binding import OracleSimulation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "motivate_b200", "motivate_b200_run")
hexf = lambda x: float.fromhex(x)


def small_config(tmp_path, people=1500, sims=2, bikes=0, cars=0):
    cfg = tmp_path / "config"
    cfg.mkdir()
    d = yaml.safe_load(open(os.path.join(ROOT, "examples", "config", "scenario.yaml")))
    d["number_of_cars"], d["number_of_bikes"] = int(0.58 * people), int(0.5 * people)
    d["intervention"]["change_in_number_of_bikes"], d["intervention"]["change_in_number_of_cars"] = bikes, cars
    for nb in d["neighbourhoods"]:
        nb["capacity"]["Car"] = people // 12
        nb["capacity"]["PublicTransport"] = people // 10
    yaml.safe_dump(d, open(cfg / "scenario.yaml", "w"))
    p = yaml.safe_load(open(os.path.join(ROOT, "examples", "config", "parameters.yaml")))
    p.update(number_of_people=people, number_of_simulations=sims, number_of_social_network_links=3, number_of_neighbour_links=4)
    yaml.safe_dump(p, open(cfg / "parameters.yaml", "w"))
    return str(tmp_path)


def test_yaml_reader_agrees_with_pyyaml_on_the_example_config():
    out = subprocess.run([BIN, "--dump", "--dir", os.path.join(ROOT, "examples")], capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    prm = Parameters.from_file(os.path.join(ROOT, "examples", "config", "parameters.yaml"))
    sc = Scenario.from_file(os.path.join(ROOT, "examples", "config", "scenario.yaml"))
    f = lines[0].split()
    assert [int(f[1]), int(f[2]), int(f[3])] == [prm.total_years, prm.number_of_people, prm.number_of_simulations]
    assert [np.float32(hexf(x)) for x in f[4:7]] == [np.float32(prm.social_connectivity), np.float32(prm.subculture_connectivity),
                                                       np.float32(prm.neighbourhood_connectivity)]
    assert [int(x) for x in f[7:10]] == [prm.number_of_social_network_links, prm.number_of_neighbour_links, prm.days_in_habit_average]
    dist = [[hexf(x) for x in l.split()[1:]] for l in lines if l.startswith("distribution")]
    assert dist == [list(map(float, d)) for d in prm.distributions]
    subs = [l.split("|") for l in lines if l.startswith("subculture|")]
    assert [s[1] for s in subs] == sc.subculture_ids
    assert np.array_equal(np.array([[hexf(x) for x in s[2].split()] for s in subs], np.float32), sc.tables.desirability)
    nbs = [l.split("|") for l in lines if l.startswith("neighbourhood|")]
    assert [n[1] for n in nbs] == sc.neighbourhood_ids
    assert np.array_equal(np.array([[hexf(x) for x in n[2].split()] for n in nbs], np.float32), sc.tables.supportiveness)
    assert np.array_equal(np.array([[int(x) for x in n[3].split()] for n in nbs], np.uint32), sc.tables.capacity)
    iv = [l for l in lines if l.startswith("intervention")][0].split()
    assert [int(x) for x in iv[1:]] == [sc.intervention.day, sc.intervention.change_in_number_of_bikes,
                                        sc.intervention.change_in_number_of_cars, 2, 1]
    nch = [l.split("|") for l in lines if l.startswith("nchange|")]
    assert nch[0][1] == "north" and "c0=-200" in nch[0][2] and "s0=" in nch[0][2] and nch[1][2].strip() == "c1=400"


@pytest.mark.skipif(not os.path.exists("/root/reference/config/scenario.yaml"), reason="reference tree not present")
def test_yaml_reader_reads_the_reference_shipped_config(tmp_path):
    cfg = tmp_path / "config"
    cfg.mkdir()
    for f in ("parameters.yaml", "scenario.yaml"):
        shutil.copy(os.path.join("/root/reference/config", f), cfg / f)          # read-only input of this test, not committed
    out = subprocess.run([BIN, "--dump", "--dir", str(tmp_path)], capture_output=True, text=True, check=True).stdout
    sc = Scenario.from_file("/root/reference/config/scenario.yaml")
    nbs = [l.split("|") for l in out.splitlines() if l.startswith("neighbourhood|")]
    assert [n[1] for n in nbs] == sc.neighbourhood_ids and len(nbs) == 20
    assert np.array_equal(np.array([[hexf(x) for x in n[2].split()] for n in nbs], np.float32), sc.tables.supportiveness)
    assert "parameters 5 111166 8" in out and "intervention 0 0 0 5 0" in out


def test_generate_writes_reference_format_files_and_fails_loudly_without_gpu(tmp_path):
    d = small_config(tmp_path)
    r = subprocess.run([BIN, "--generate", "--dir", d, "--seed", "3"], capture_output=True, text=True)
    if mb._ffi.load().mvb_device_count() == 0:
        assert r.returncode == 1 and "no CPU path" in r.stderr               # the product has no CPU fallback
    sc = Scenario.from_file(os.path.join(d, "config", "scenario.yaml"))
    # what it saved is what the reference's --generate saves: networks {id: [ids]} and agents with serde's field names
    pop = files.load_population(os.path.join(d, "config", "agents", "1.yaml"), os.path.join(d, "config", "networks", "1.yaml"),
                                sc, number_of_neighbour_links=4, seed=3 * 1000 + 1)
    assert pop.n_agents == 1500 and int(pop["owns_car"].sum()) == sc.number_of_cars
    assert np.all(pop["consistency"] == 1.0) and np.all(pop["average_weight"] == np.float32(2.0) / np.float32(11.0))
    assert abs(len(pop["social_col"]) / 1500 - 6) < 0.1                       # BA with 3 links per new node
    assert np.array_equal(pop["habit"].argmax(1), pop["current_mode"])


@pytest.mark.gpu
def test_cpp_host_run_equals_oracle(tmp_path):
    """Full `motivate_b200_run --generate` (2 replicas, 2 years, table intervention on day 365), then the same inputs
    rebuilt in Python from the files it saved and run on the oracle: the CSVs must be identical."""
    d = small_config(tmp_path)
    subprocess.run([BIN, "--generate", "--dir", d, "--seed", "3"], check=True, capture_output=True)
    sc = Scenario.from_file(os.path.join(d, "config", "scenario.yaml"))
    for sim_id in (1, 2):
        lines = open(os.path.join(d, "output", f"output_{sim_id}.csv")).read().splitlines()
        assert lines[0].startswith("Day,Rain,ActiveMode,") and len(lines) == 1 + mb.run_rows(730)
        body = np.array([[int(x) for x in l.split(",")] for l in lines[1:]])
        weather = np.zeros(730, np.uint8)
        weather[body[:, 0]] = body[:, 1]                                       # Rain column; weekends never matter
        pop = files.load_population(os.path.join(d, "config", "agents", f"{sim_id}.yaml"),
                                    os.path.join(d, "config", "networks", f"{sim_id}.yaml"), sc, 4, seed=3 * 1000 + sim_id)
        iv = sc.intervention
        ncs = [(sc.neighbourhood_ids.index(c.id), c.increase_in_supportiveness, c.increase_in_capacity) for c in iv.neighbourhood_changes]
        scs = [(sc.subculture_ids.index(c.id), c.increase_in_desirability) for c in iv.subculture_changes]
        ivd = mb.InterventionData(day=iv.day, tables=apply_table_changes(sc.tables, ncs, scs))
        o = OracleSimulation(pop, sc.tables, mb.make_params())
        days, rows = o.run(weather, 730, intervention=ivd)
        assert np.array_equal(days, body[:, 0]) and np.array_equal(rows, body[:, 2:]), f"replica {sim_id}"
    # a second run from the saved files (no --generate) reproduces the same CSVs
    first = open(os.path.join(d, "output", "output_1.csv")).read()
    subprocess.run([BIN, "--dir", d, "--seed", "3"], check=True, capture_output=True)
    assert open(os.path.join(d, "output", "output_1.csv")).read() == first
