#!/usr/bin/env python
"""Writes examples/c1/c1.json: BASELINE.json configs[0] — the reference's SHIPPED configuration
(/root/reference/config/parameters.yaml + scenario.yaml) with the three fixes without which it cannot run
(none of them touches a table value):

  1. parameters.yaml:3 `distributions: []` makes --generate panic (gaussian.rs:77-82).  The intended commute-distance
     mixture is in data/Travel To Work.ipynb cell 14: (mean, sd = sqrt(variance), weight) x 3.
  2. scenario.yaml:6 `intervention.day: 0` never fires (the loop starts at day 1, simulation.rs:101-103).
     BASELINE.json says "intervention at day 365" (README.md:49).
  3. the intervention names neighbourhoods "E0500590" ... (8 characters, scenario.yaml:8-36) while the neighbourhoods
     are "E05000590" ... (9, :50): `intervene` would panic at simulation.rs:318.  Ids corrected.

The values are stored as compact arrays (mode order Car, PublicTransport, Cycle, Walk), not as a copy of the YAML;
`motivate_b200.c1.materialise(dir)` writes them back out in the reference's own schema (config/parameters.yaml,
config/scenario.yaml) for the hosts to read.  Run here, where the reference tree exists; the JSON travels."""
import json
import math
import os
import sys

import yaml

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
MODES = ["Car", "PublicTransport", "Cycle", "Walk"]
HERE = os.path.dirname(os.path.abspath(__file__))

prm = yaml.safe_load(open(os.path.join(REF, "config", "parameters.yaml")))
sc = yaml.safe_load(open(os.path.join(REF, "config", "scenario.yaml")))

fixes = []
assert prm["distributions"] == []
gmm = [(9845.09, 13629484.66, 0.66987), (2205.45, 2000457.67, 0.25230), (22606.84, 143161746.89, 0.07782)]   # notebook cell 14
prm["distributions"] = [[m, math.sqrt(v), w] for m, v, w in gmm]
fixes.append("parameters.distributions: [] -> the 3-component mixture of data/Travel To Work.ipynb cell 14 (mean, sqrt(variance), weight)")
assert sc["intervention"]["day"] == 0
sc["intervention"]["day"] = 365
fixes.append("scenario.intervention.day: 0 -> 365")
ids = [n["id"] for n in sc["neighbourhoods"]]
for ch in sc["intervention"]["neighbourhood_changes"]:
    assert ch["id"] not in ids and len(ch["id"]) == 8
    fixed = ch["id"][:4] + "0" + ch["id"][4:]
    assert fixed in ids, fixed
    ch["id"] = fixed
fixes.append("scenario.intervention.neighbourhood_changes[*].id: 8-character ids -> the 9-character neighbourhood ids")

out = {
    "source": "0xr0bert/Motivate config/parameters.yaml + config/scenario.yaml", "fixes": fixes, "mode_order": MODES,
    "parameters": prm,
    "scenario": {
        "id": sc["id"], "number_of_bikes": sc["number_of_bikes"], "number_of_cars": sc["number_of_cars"],
        "subculture_ids": [s["id"] for s in sc["subcultures"]],
        "desirability": [[s["desirability"][m] for m in MODES] for s in sc["subcultures"]],
        "neighbourhood_ids": ids,
        "supportiveness": [[n["supportiveness"][m] for m in MODES] for n in sc["neighbourhoods"]],
        "capacity": [[n["capacity"][m] for m in MODES] for n in sc["neighbourhoods"]],
        "intervention": {
            "day": sc["intervention"]["day"],
            "change_in_number_of_bikes": sc["intervention"]["change_in_number_of_bikes"],
            "change_in_number_of_cars": sc["intervention"]["change_in_number_of_cars"],
            "neighbourhood_changes": [
                {"id": c["id"], "increase_in_supportiveness": c.get("increase_in_supportiveness", {}),
                 "increase_in_capacity": c.get("increase_in_capacity", {})} for c in sc["intervention"]["neighbourhood_changes"]],
            "subculture_changes": sc["intervention"]["subculture_changes"],
        },
    },
}
json.dump(out, open(os.path.join(HERE, "c1.json"), "w"), indent=1)
print("wrote", os.path.join(HERE, "c1.json"), "with", len(ids), "neighbourhoods")
