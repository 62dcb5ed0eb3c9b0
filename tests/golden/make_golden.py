#!/usr/bin/env python
"""Generates tests/golden/*.npz: inputs and per-day outputs of the LITERAL model (oracle/literal_model.py — the
restatement that keeps the reference's HashMap / union_of / stable-sort structure), canonical map order.

The reference itself cannot run here (Rust, no cargo/rustc in the image) and ships no golden vectors, so these
fixtures are self-made (DESIGN.md "parity unpinned").  They pin the C oracle on the CPU and the CUDA path on the
GPU to the literal model without needing it at test time.

    python tests/golden/make_golden.py        # rewrites the .npz files (deterministic)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import motivate_b200 as mb                                   # noqa: E402
from common import random_population, random_tables, tab_dict   # noqa: E402
from oracle import literal_model as lm                       # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, pop, tab, prm, weather, n_days, intervention=None, iv_struct=None):
    snaps = {}

    def dump(day, agents, nbhds):
        h, c, l, n = lm.dense_state(agents)
        snaps[day] = (h, c, n)

    days, rows, agents, nbhds = lm.run(pop.arrays, tab_dict(tab), mb.params_dict(prm), weather, n_days,
                                       intervention=intervention, dump=dump)
    out = {f"pop_{k}": v for k, v in pop.arrays.items() if v is not None}
    out.update(tab_supportiveness=tab.supportiveness, tab_capacity=tab.capacity, tab_desirability=tab.desirability,
               weather=np.asarray(weather, np.uint8), n_days=np.int64(n_days), days=np.array(days, np.uint32),
               rows=np.array(rows, np.uint32),
               params=np.array([prm.consistency, prm.social_connectivity, prm.subculture_connectivity,
                                prm.neighbourhood_connectivity, prm.average_weight], np.float32),
               habit=np.stack([snaps[d][0] for d in days]), mode=np.stack([snaps[d][1] for d in days]),
               norm=np.stack([snaps[d][2] for d in days]))
    if iv_struct is not None:
        out.update(iv_day=np.int64(iv_struct["day"]), iv_supportiveness=iv_struct["tab"].supportiveness,
                   iv_capacity=iv_struct["tab"].capacity, iv_desirability=iv_struct["tab"].desirability,
                   iv_owns_car=iv_struct["car"], iv_owns_bike=iv_struct["bike"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "rows", len(days), "agents", pop.n_agents)


def main():
    # 1. unstructured population: duplicate links, empty lists, sparse habits, quantised tables, tight capacities
    n, S, H, n_days = 96, 3, 4, 23
    pop = random_population(n, S, H, 101)
    make("random_uniform_weights", pop, random_tables(S, H, n, 101), mb.make_params(days_in_habit_average=6),
         mb.weather_pattern(n_days, seed=7, p_good_given_good=0.5, p_good_given_bad=0.5), n_days)
    # 2. per-agent weights (hand-edited agents file, agent.rs:44-57)
    pop = random_population(n, S, H, 202, per_agent=True)
    make("random_per_agent_weights", pop, random_tables(S, H, n, 202), mb.make_params(),
         mb.weather_pattern(n_days, seed=8, p_good_given_good=0.5, p_good_given_bad=0.5), n_days)
    # 3. synthetic generator population (BA links, hubs) with an intervention on a weekend day
    n, S, H, n_days = 400, 4, 5, 30
    pop = mb.synth_population(n, S, H, 3, 4, seed=303)
    tab = mb.synth_tables(S, H, n, seed=303)
    rng = np.random.default_rng(303)
    car2 = pop["owns_car"].copy(); car2[rng.choice(n, 40, replace=False)] ^= 1
    bike2 = pop["owns_bike"].copy(); bike2[rng.choice(n, 40, replace=False)] ^= 1
    from motivate_b200.scenario import apply_table_changes
    ncs = [(0, {2: 0.1, 3: 0.1, 0: -0.1}, {0: -20}), (2, {1: 0.05}, {1: -10, 0: 5})]
    scs = [(1, {0: -0.1, 2: 0.1})]
    tab2 = apply_table_changes(tab, ncs, scs)
    iv = dict(day=13, neighbourhood_changes=ncs, subculture_changes=scs, owns_car=car2, owns_bike=bike2)
    make("synthetic_with_intervention", pop, tab, mb.make_params(), mb.weather_pattern(n_days, seed=9), n_days,
         intervention=iv, iv_struct=dict(day=13, tab=tab2, car=car2, bike=bike2))


if __name__ == "__main__":
    main()
