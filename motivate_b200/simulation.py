"""`simulation::run` of the reference (src/simulation.rs:40-149) on top of the C ABI: same arguments, same files in,
same CSV out; the weekday loop itself (:101-136) is one `mvb_run` call on the GPU.

    prepare(...)  setup part (:62-78): scenario, agents (generated or loaded), link graphs, resolved intervention
    run(...)      prepare + mvb_create + mvb_run + output/output_{id}.csv
`run_many` is the rayon fan-out of src/main.rs:89-121: all replicas of a GPU are created as one batch and stepped
together.
"""
from __future__ import annotations

import os
import zlib
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import files
from .api import (DEFAULT_GMM, InterventionData, PopulationData, Simulation, TablesData, make_params,
                  synth_population)
from .scenario import Scenario, intervene, write_csv


@dataclass
class Prepared:
    scenario: Scenario
    population: PopulationData
    tables: TablesData
    intervention: Optional[InterventionData]
    params: object


def prepare(id: str, generate: bool, agents_file: str, scenario_file: str, number_of_people: int,
            social_connectivity: float, subculture_connectivity: float, neighbourhood_connectivity: float,
            number_of_neighbour_links: int, days_in_habit_average: int, distributions: Sequence,
            network=None, number_of_social_network_links: int = 5, seed: Optional[int] = None) -> Prepared:
    """Everything `simulation::run` does before the day loop (src/simulation.rs:62-78).  `network` is the social
    network as (rowptr, col) or a path to config/networks/{id}.yaml; with generate=True and network=None a seeded
    Barabási–Albert network is made (what main.rs --generate does before calling run)."""
    scenario = Scenario.from_file(scenario_file)
    if seed is None:       # deterministic across processes (str hashes are salted): numeric ids as main.cpp's seed*1000+id
        seed = int(id) if str(id).isdigit() else zlib.crc32(str(id).encode())
    S, H = len(scenario.subculture_ids), len(scenario.neighbourhood_ids)
    if generate:
        gmm = [tuple(d) for d in distributions] or list(DEFAULT_GMM)
        pop = synth_population(number_of_people, S, H, number_of_social_network_links, number_of_neighbour_links,
                               car_ratio=scenario.number_of_cars / number_of_people,
                               bike_ratio=scenario.number_of_cars / number_of_people,   # agent_generation.rs:160 (sic)
                               seed=seed, gmm=gmm)
        if network is not None:
            rp, col = files.read_network(network, number_of_people) if isinstance(network, str) else network
            pop.arrays["social_rowptr"], pop.arrays["social_col"] = np.asarray(rp, np.uint64), np.asarray(col, np.uint32)
        # the neighbourhood networks are never saved: every run rebuilds them from the agents' neighbourhoods with a fixed
        # seed (simulation.rs:165-189, social_network.rs:15), so a --generate run and a later run from the files agree
        pop.arrays["neighbour_rowptr"], pop.arrays["neighbour_col"] = files.neighbour_networks(
            pop["neighbourhood"], H, number_of_neighbour_links, seed)
        pop._struct = None
        if agents_file:
            params = make_params(1.0, social_connectivity, subculture_connectivity, neighbourhood_connectivity,
                                 days_in_habit_average)
            files.write_agents(agents_file, pop, scenario, params)                      # save_agents, :60-63
    else:
        if network is None:
            raise ValueError("a social network is required when agents are loaded from a file")
        if not isinstance(network, str):
            rp, col = network
            tmp = files.read_agents(agents_file, scenario)
            tmp["social_rowptr"], tmp["social_col"] = np.asarray(rp, np.uint64), np.asarray(col, np.uint32)
            tmp["neighbour_rowptr"], tmp["neighbour_col"] = files.neighbour_networks(tmp["neighbourhood"], H,
                                                                                     number_of_neighbour_links, seed)
            pop = PopulationData(tmp)
        else:
            pop = files.load_population(agents_file, network, scenario, number_of_neighbour_links, seed)
    params = make_params(1.0, social_connectivity, subculture_connectivity, neighbourhood_connectivity,
                         days_in_habit_average)
    iv = intervene(scenario, pop, seed=seed) if scenario.intervention.day >= 1 else None   # day 0 never fires (:101-103)
    return Prepared(scenario, pop, scenario.tables, iv, params)


def run(id: str, generate: bool, agents_file: str, scenario_file: str, total_years: int, number_of_people: int,
        social_connectivity: float, subculture_connectivity: float, neighbourhood_connectivity: float,
        number_of_neighbour_links: int, days_in_habit_average: int, distributions: Sequence,
        weather_pattern, network=None, output_dir: str = "output", device: int = 0, **kw):
    """src/simulation.rs:40-149.  Returns (days, rows) and writes {output_dir}/output_{id}.csv."""
    prep = prepare(id, generate, agents_file, scenario_file, number_of_people, social_connectivity,
                   subculture_connectivity, neighbourhood_connectivity, number_of_neighbour_links,
                   days_in_habit_average, distributions, network, **kw)
    n_days = total_years * 365
    if len(weather_pattern) < n_days:
        raise ValueError("weather_pattern shorter than total_years * 365")
    with Simulation(prep.population, prep.tables, prep.params, device=device) as sim:
        days, rows = sim.run(weather_pattern, n_days, interventions=prep.intervention)
    os.makedirs(output_dir, exist_ok=True)
    write_csv(os.path.join(output_dir, f"output_{id}.csv"), prep.scenario, days, weather_pattern, rows[0])
    return days, rows[0]


def run_many(preps: Sequence[Prepared], ids: Sequence[str], total_years: int, weather_pattern,
             output_dir: str = "output", device: int = 0):
    """All replicas of one GPU in one batch (src/main.rs:89-121)."""
    n_days = total_years * 365
    # mvb_create_batch takes ONE mvb_params for the handle: replicas with other uniform weights must not be batched
    key = lambda q: (q.consistency, q.social_connectivity, q.subculture_connectivity, q.neighbourhood_connectivity,
                     q.average_weight, q.flags)
    if any(key(p.params) != key(preps[0].params) for p in preps):
        raise ValueError("run_many: all replicas of a batch must share the same uniform weights (mvb_params); "
                         "batch them separately or give per-agent weight arrays")
    with Simulation([p.population for p in preps], [p.tables for p in preps], preps[0].params, device=device) as sim:
        days, rows = sim.run(weather_pattern, n_days, interventions=[p.intervention for p in preps])
    os.makedirs(output_dir, exist_ok=True)
    for p, i, r in zip(preps, ids, rows):
        write_csv(os.path.join(output_dir, f"output_{i}.csv"), p.scenario, days, weather_pattern, r)
    return days, rows
