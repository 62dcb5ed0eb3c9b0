"""BASELINE.json configs[0]: the reference's shipped configuration (config/parameters.yaml + config/scenario.yaml,
111 166 agents, 20 neighbourhoods, 3 subcultures) with the three fixes it needs to run at all — see
examples/c1/make_c1.py.  `materialise(dir)` writes DIR/config/{parameters,scenario}.yaml in the reference's schema."""
from __future__ import annotations

import json
import os

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
C1_JSON = os.path.join(os.path.dirname(HERE), "examples", "c1", "c1.json")


def dicts(path: str = C1_JSON, *, people=None, simulations=None, total_years=None, tie_free=False):
    """(parameters, scenario) as the dicts serde_yaml would read (src/main.rs:185-212, src/scenario.rs:11-27).
    people: scale number_of_people and, in proportion, cars / bikes / Car + PublicTransport capacities (tests).
    tie_free: perturb equal table entries by distinct multiples of 2^-12 so that no two modes tie anywhere
    (for comparisons with a real Rust run, whose HashMap iteration order is random: rust/patch/README.md)."""
    c = json.load(open(path))
    modes = c["mode_order"]
    prm = dict(c["parameters"])
    s = c["scenario"]
    des = [list(r) for r in s["desirability"]]
    sup = [list(r) for r in s["supportiveness"]]
    cap = [list(r) for r in s["capacity"]]
    n_cars, n_bikes = s["number_of_cars"], s["number_of_bikes"]
    if people is not None and people != prm["number_of_people"]:
        f = people / prm["number_of_people"]
        prm["number_of_people"] = int(people)
        n_cars, n_bikes = int(round(n_cars * f)), int(round(n_bikes * f))
        cap = [[max(1, int(round(v * f))) if k < 2 else v for k, v in enumerate(r)] for r in cap]
    if simulations is not None:
        prm["number_of_simulations"] = int(simulations)
    if total_years is not None:
        prm["total_years"] = int(total_years)
    if tie_free:
        eps = 2.0 ** -12
        des = [[v + eps * (k + 1) * (i + 1) for k, v in enumerate(r)] for i, r in enumerate(des)]
        sup = [[v + eps * (k + 1) for k, v in enumerate(r)] for r in sup]
    scen = {
        "id": s["id"], "number_of_bikes": n_bikes, "number_of_cars": n_cars,
        "subcultures": [{"id": i, "desirability": dict(zip(modes, r))} for i, r in zip(s["subculture_ids"], des)],
        "neighbourhoods": [{"id": i, "supportiveness": dict(zip(modes, a)), "capacity": dict(zip(modes, b))}
                           for i, a, b in zip(s["neighbourhood_ids"], sup, cap)],
        "intervention": json.loads(json.dumps(s["intervention"])),
    }
    if people is not None:
        f = prm["number_of_people"] / c["parameters"]["number_of_people"]
        for ch in scen["intervention"]["neighbourhood_changes"]:
            ch["increase_in_capacity"] = {k: int(round(v * f)) for k, v in ch["increase_in_capacity"].items()}
    return prm, scen


def materialise(directory: str, **kw):
    """Write DIR/config/parameters.yaml and DIR/config/scenario.yaml; returns their paths."""
    prm, scen = dicts(**kw)
    cfg = os.path.join(directory, "config")
    os.makedirs(cfg, exist_ok=True)
    pp, sp = os.path.join(cfg, "parameters.yaml"), os.path.join(cfg, "scenario.yaml")
    yaml.safe_dump(prm, open(pp, "w"), default_flow_style=False)
    yaml.safe_dump(scen, open(sp, "w"), default_flow_style=False)
    return pp, sp
