"""Host-side mirror of the reference's scenario / parameters / intervention surface for the hot path.

Same YAML schema, same names and the same error behaviour (what panics there raises here):
    Parameters           src/main.rs:185-228          config/parameters.yaml
    Scenario             src/scenario.rs:11-45        config/scenario.yaml
    Intervention, NeighbourhoodChange, SubcultureChange   src/intervention.rs:6-52
    intervene            src/simulation.rs:304-464    (resolved into tables + ownership bits for the device)
    generate_csv_header / generate_csv_output            src/simulation.rs:194-268
This module only prepares inputs for, and formats outputs of, the C ABI; it does no per-agent arithmetic.

f32 fields are parsed decimal -> f64 -> f32 exactly like serde_yaml + `as f32` (SURVEY.md §5 "Config").
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import yaml

from .api import InterventionData, PopulationData, TablesData

MODE_NAMES = ("Car", "PublicTransport", "Cycle", "Walk")           # src/transport_mode.rs:3-8
MODE_INDEX = {n: i for i, n in enumerate(MODE_NAMES)}


class ScenarioError(ValueError):
    """Raised where the reference would panic on a malformed scenario."""


def _mode_map(d, what, dtype, require_all=True) -> Dict[int, object]:
    out = {}
    for k, v in (d or {}).items():
        if k not in MODE_INDEX:
            raise ScenarioError(f"{what}: unknown transport mode {k!r}")
        out[MODE_INDEX[k]] = dtype(v)
    if require_all and len(out) != 4:
        # SURVEY.md §8 a-K: the build requires all four modes in every desirability / supportiveness / capacity map
        raise ScenarioError(f"{what}: all four transport modes are required, got {sorted(d or {})}")
    return out


@dataclass
class Parameters:
    """src/main.rs:185-212."""
    total_years: int
    number_of_people: int
    number_of_simulations: int
    social_connectivity: float
    subculture_connectivity: float
    neighbourhood_connectivity: float
    number_of_social_network_links: int
    number_of_neighbour_links: int
    days_in_habit_average: int
    distributions: list

    @classmethod
    def from_file(cls, path):
        with open(path) as f:
            d = yaml.safe_load(f)
        try:
            return cls(**{k: d[k] for k in cls.__dataclass_fields__})
        except KeyError as e:
            raise ScenarioError(f"There was an error parsing the file: missing field {e}") from None


@dataclass
class NeighbourhoodChange:
    id: str
    increase_in_supportiveness: Dict[int, np.float32] = field(default_factory=dict)   # #[serde(default)]
    increase_in_capacity: Dict[int, int] = field(default_factory=dict)


@dataclass
class SubcultureChange:
    id: str
    increase_in_desirability: Dict[int, np.float32] = field(default_factory=dict)


@dataclass
class Intervention:
    day: int
    neighbourhood_changes: List[NeighbourhoodChange]
    subculture_changes: List[SubcultureChange]
    change_in_number_of_bikes: int
    change_in_number_of_cars: int


@dataclass
class Scenario:
    """src/scenario.rs:11-27."""
    id: str
    subculture_ids: List[str]
    neighbourhood_ids: List[str]
    tables: TablesData
    number_of_bikes: int
    number_of_cars: int
    intervention: Intervention

    @classmethod
    def from_file(cls, path):
        with open(path) as f:
            return cls.from_dict(yaml.safe_load(f))

    @classmethod
    def from_dict(cls, d):
        subs, nbs = d["subcultures"], d["neighbourhoods"]
        des = np.zeros((len(subs), 4), np.float32)
        for i, s in enumerate(subs):
            for m, v in _mode_map(s["desirability"], f"subculture {s['id']}", np.float64).items():
                des[i, m] = np.float32(v)
        sup = np.zeros((len(nbs), 4), np.float32)
        cap = np.zeros((len(nbs), 4), np.uint32)
        for i, n in enumerate(nbs):
            for m, v in _mode_map(n["supportiveness"], f"neighbourhood {n['id']}", np.float64).items():
                sup[i, m] = np.float32(v)
            for m, v in _mode_map(n["capacity"], f"neighbourhood {n['id']}", int).items():
                if not 0 <= v <= 0xFFFFFFFF:
                    raise ScenarioError(f"neighbourhood {n['id']}: capacity {v} does not fit u32")
                cap[i, m] = v
        iv = d["intervention"]
        ncs = [NeighbourhoodChange(str(c["id"]),
                                   _mode_map(c.get("increase_in_supportiveness"), "increase_in_supportiveness",
                                             lambda x: np.float32(np.float64(x)), require_all=False),
                                   _mode_map(c.get("increase_in_capacity"), "increase_in_capacity", int, require_all=False))
               for c in (iv.get("neighbourhood_changes") or [])]
        scs = [SubcultureChange(str(c["id"]),
                                _mode_map(c["increase_in_desirability"], "increase_in_desirability",
                                          lambda x: np.float32(np.float64(x)), require_all=False))
               for c in (iv.get("subculture_changes") or [])]
        return cls(id=str(d["id"]), subculture_ids=[str(s["id"]) for s in subs],
                   neighbourhood_ids=[str(n["id"]) for n in nbs], tables=TablesData(sup, cap, des),
                   number_of_bikes=int(d["number_of_bikes"]), number_of_cars=int(d["number_of_cars"]),
                   intervention=Intervention(int(iv["day"]), ncs, scs, int(iv["change_in_number_of_bikes"]),
                                             int(iv["change_in_number_of_cars"])))


def apply_table_changes(tab: TablesData, neighbourhood_changes, subculture_changes) -> TablesData:
    """Table part of `intervene` (src/simulation.rs:307-381) on index-resolved changes:
    neighbourhood_changes = [(h, {mode: d_supportiveness}, {mode: d_capacity})], subculture_changes = [(s, {mode: d})].
    supportiveness/desirability: f32 add (:320-325, :370-379); capacity: i64 add clamped to [0, u32::MAX] (:328-351).
    Changes are applied in list order, so a neighbourhood listed twice accumulates."""
    out = tab.copy()
    for h, d_supp, d_cap in neighbourhood_changes:
        for m, v in d_supp.items():
            out.supportiveness[h, m] = np.float32(out.supportiveness[h, m]) + np.float32(v)
        for m, v in d_cap.items():
            nv = int(out.capacity[h, m]) + int(v)
            out.capacity[h, m] = 0 if nv < 0 else (0xFFFFFFFF if nv > 0xFFFFFFFF else nv)
    for s, d in subculture_changes:
        for m, v in d.items():
            out.desirability[s, m] = np.float32(out.desirability[s, m]) + np.float32(v)
    return out


def _resample_ownership(owned: np.ndarray, change: int, rng, what: str) -> Optional[np.ndarray]:
    """Car / bike part of `intervene` (src/simulation.rs:383-463): give to |change| uniformly sampled non-owners
    (change > 0) or take from |change| owners (change < 0).  sample_slice_ref panics if the pool is too small."""
    if change == 0:
        return None
    pool = np.flatnonzero(owned == (0 if change > 0 else 1))
    k = abs(change)
    if k > len(pool):
        raise ScenarioError(f"change_in_number_of_{what} = {change} exceeds the {len(pool)} eligible agents")
    new = owned.copy()
    new[rng.choice(pool, k, replace=False)] = 1 if change > 0 else 0
    return new


def intervene(scenario: Scenario, pop: PopulationData, tables: Optional[TablesData] = None, seed=0) -> InterventionData:
    """Resolve scenario.intervention into an `mvb_intervention` (tables and ownership that hold from `day` on).
    Unknown ids raise like the reference's `expect` (src/simulation.rs:318,368).  The reference samples with
    thread_rng; here the sample is seeded."""
    iv = scenario.intervention
    tab = tables if tables is not None else scenario.tables
    ncs = []
    for c in iv.neighbourhood_changes:
        if c.id not in scenario.neighbourhood_ids:
            raise ScenarioError("A neighbourhood in your intervention was not found")
        ncs.append((scenario.neighbourhood_ids.index(c.id), c.increase_in_supportiveness, c.increase_in_capacity))
    scs = []
    for c in iv.subculture_changes:
        if c.id not in scenario.subculture_ids:
            raise ScenarioError("A subculture in your intervention was not found")
        scs.append((scenario.subculture_ids.index(c.id), c.increase_in_desirability))
    rng = np.random.default_rng(seed)
    bikes = _resample_ownership(pop["owns_bike"], iv.change_in_number_of_bikes, rng, "bikes")
    cars = _resample_ownership(pop["owns_car"], iv.change_in_number_of_cars, rng, "cars")
    return InterventionData(day=iv.day, tables=apply_table_changes(tab, ncs, scs), owns_car=cars, owns_bike=bikes)


def generate_csv_header(scenario: Scenario) -> str:
    """src/simulation.rs:194-212."""
    return ("Day,Rain,ActiveMode,ActiveNorm,ActiveModeCounterToInactiveNorm,InactiveModeCounterToActiveNorm,"
            "LocalCommute,CityCommute,DistantCommute,{},{}\n").format(",".join(scenario.subculture_ids),
                                                                      ",".join(scenario.neighbourhood_ids))


def generate_csv_output(day: int, rain: int, row) -> str:
    """src/simulation.rs:220-268; `row` is a statistics row from mvb_run / mvb_stats_row."""
    return "{},{},{}\n".format(day, rain, ",".join(str(int(v)) for v in row))


def write_csv(path, scenario: Scenario, days, weather, rows):
    """output/output_{id}.csv of src/simulation.rs:88-98,134: Rain is the weather of the row's own day."""
    with open(path, "w") as f:
        f.write(generate_csv_header(scenario))
        for d, r in zip(days, rows):
            f.write(generate_csv_output(int(d), int(weather[int(d)]), r))
