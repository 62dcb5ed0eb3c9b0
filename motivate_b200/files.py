"""Readers / writers for the reference's on-disk formats, so a population written by the
reference's `--generate` run can be loaded as it is:

    config/networks/{id}.yaml   serde_yaml of HashMap<u32, Vec<u32>>            (src/main.rs:135-148, 173-179)
    config/agents/{id}.yaml     serde_yaml of Vec<Agent> minus #[serde(skip)]   (src/agent.rs:15-87,
                                                                                  src/agent_generation.rs:26-63)
and the neighbourhood networks the reference regenerates at every run (src/simulation.rs:165-189): one
Barabási–Albert network per neighbourhood over its residents in agent order — rebuilt here with the library's
seeded generator, because the reference never saves them.

These run once per replica at setup; the hot path never sees YAML.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import yaml

from . import _ffi
from .api import PopulationData, _check
from .scenario import MODE_INDEX, MODE_NAMES, Scenario, ScenarioError

JOURNEY_NAMES = ("LocalCommute", "CityCommute", "DistantCommute")          # src/journey_type.rs:5-9
JOURNEY_INDEX = {n: i for i, n in enumerate(JOURNEY_NAMES)}

try:                                                                        # libyaml if present: 10x faster on big files
    _Loader, _Dumper = yaml.CSafeLoader, yaml.CSafeDumper
except AttributeError:                                                      # pragma: no cover
    _Loader, _Dumper = yaml.SafeLoader, yaml.SafeDumper


def lists_to_csr(network: dict, n_agents: int):
    """{id: [friend ids]} -> CSR; ids missing from the map have no links (src/simulation.rs:274-291)."""
    rowptr = np.zeros(n_agents + 1, np.uint64)
    for k, v in network.items():
        k = int(k)
        if not 0 <= k < n_agents:
            raise ScenarioError(f"network id {k} outside 0..{n_agents - 1}")     # the reference indexes agents[k]: panic
        rowptr[k + 1] = len(v)
    np.cumsum(rowptr, out=rowptr)
    col = np.zeros(int(rowptr[-1]), np.uint32)
    for k, v in network.items():
        a = np.asarray(v, np.int64)
        if a.size and (a.min() < 0 or a.max() >= n_agents):
            raise ScenarioError(f"network of agent {k} names an agent outside 0..{n_agents - 1}")
        col[int(rowptr[int(k)]):int(rowptr[int(k) + 1])] = a
    return rowptr, col


def read_network(path, n_agents: int):
    """config/networks/{id}.yaml -> (rowptr u64[N+1], col u32[E]) with duplicates kept."""
    with open(path) as f:
        net = yaml.load(f, Loader=_Loader) or {}
    return lists_to_csr(net, n_agents)


def write_network(path, rowptr, col):
    net = {int(i): [int(x) for x in col[int(rowptr[i]):int(rowptr[i + 1])]] for i in range(len(rowptr) - 1)}
    with open(path, "w") as f:
        yaml.dump(net, f, Dumper=_Dumper, default_flow_style=False)


def neighbour_networks(neighbourhood: np.ndarray, n_neighbourhoods: int, min_links: int, seed: int = 0):
    """One BA network per neighbourhood over its residents in agent order (src/simulation.rs:165-189), as one CSR in
    global agent ids.  A neighbourhood with fewer than 2 residents has no links."""
    lib = _ffi.load()
    n = len(neighbourhood)
    lists = [[] for _ in range(n)]
    for h in range(n_neighbourhoods):
        members = np.flatnonzero(neighbourhood == h)
        if len(members) < 2:
            continue
        rp, cl = _ffi.u64p(), _ffi.u32p()
        _check(lib.mvb_synth_ba_network(min_links, len(members), seed * 1000003 + h, C.byref(rp), C.byref(cl)))
        rowptr = np.ctypeslib.as_array(rp, shape=(len(members) + 1,)).copy()
        col = np.ctypeslib.as_array(cl, shape=(int(rowptr[-1]),)).copy() if rowptr[-1] else np.zeros(0, np.uint32)
        lib.mvb_free(rp); lib.mvb_free(cl)
        for k, g in enumerate(members):
            lists[g] = members[col[int(rowptr[k]):int(rowptr[k + 1])]].tolist()
    return lists_to_csr({i: v for i, v in enumerate(lists)}, n)


def read_agents(path, scenario: Scenario):
    """config/agents/{id}.yaml -> dict of per-agent arrays (everything of mvb_population except the link graphs).
    Unknown subculture / neighbourhood ids raise like the reference's `expect` (agent_generation.rs:47-49)."""
    with open(path) as f:
        agents = yaml.load(f, Loader=_Loader) or []
    n = len(agents)
    sub_index = {s: i for i, s in enumerate(scenario.subculture_ids)}
    nb_index = {s: i for i, s in enumerate(scenario.neighbourhood_ids)}
    out = dict(subculture=np.zeros(n, np.uint8), neighbourhood=np.zeros(n, np.uint16), commute=np.zeros(n, np.uint8),
               owns_car=np.zeros(n, np.uint8), owns_bike=np.zeros(n, np.uint8), weather_sensitivity=np.zeros(n, np.float32),
               consistency=np.zeros(n, np.float32), social_connectivity=np.zeros(n, np.float32),
               subculture_connectivity=np.zeros(n, np.float32), neighbourhood_connectivity=np.zeros(n, np.float32),
               average_weight=np.zeros(n, np.float32), habit=np.zeros((n, 4), np.float32),
               current_mode=np.zeros(n, np.uint8), norm=np.zeros(n, np.uint8))
    f32 = lambda x: np.float32(np.float64(x))          # decimal -> f64 -> f32, like serde_yaml + `as f32`
    for i, a in enumerate(agents):
        if a["subculture_id"] not in sub_index:
            raise ScenarioError("Subculture not found")
        if a["neighbourhood_id"] not in nb_index:
            raise ScenarioError("Agent not found")                          # the reference's (sic) message
        out["subculture"][i] = sub_index[a["subculture_id"]]
        out["neighbourhood"][i] = nb_index[a["neighbourhood_id"]]
        out["commute"][i] = JOURNEY_INDEX[a["commute_length"]]
        out["owns_car"][i], out["owns_bike"][i] = bool(a["owns_car"]), bool(a["owns_bike"])
        for k in ("weather_sensitivity", "consistency", "social_connectivity", "subculture_connectivity",
                  "neighbourhood_connectivity", "average_weight"):
            out[k][i] = f32(a[k])
        for m, v in (a.get("habit") or {}).items():
            out["habit"][i, MODE_INDEX[m]] = f32(v)
        out["current_mode"][i] = MODE_INDEX[a["current_mode"]]
        out["norm"][i] = MODE_INDEX[a["norm"]]
    return out


def write_agents(path, pop: PopulationData, scenario: Scenario, params=None):
    """Inverse of read_agents (same field names as the reference's serde output)."""
    def w(name, i):
        arr = pop.get(name)
        return float(arr[i]) if arr is not None else float(getattr(params, name))
    agents = []
    for i in range(pop.n_agents):
        agents.append(dict(
            subculture_id=scenario.subculture_ids[int(pop["subculture"][i])],
            neighbourhood_id=scenario.neighbourhood_ids[int(pop["neighbourhood"][i])],
            commute_length=JOURNEY_NAMES[int(pop["commute"][i])], commute_length_continuous=0.0,
            weather_sensitivity=float(pop["weather_sensitivity"][i]), consistency=w("consistency", i),
            social_connectivity=w("social_connectivity", i), subculture_connectivity=w("subculture_connectivity", i),
            neighbourhood_connectivity=w("neighbourhood_connectivity", i), average_weight=w("average_weight", i),
            habit={MODE_NAMES[m]: float(pop["habit"][i, m]) for m in range(4) if pop["habit"][i, m] != 0},
            current_mode=MODE_NAMES[int(pop["current_mode"][i])], last_mode=MODE_NAMES[int(pop["current_mode"][i])],
            norm=MODE_NAMES[int(pop["norm"][i])], owns_bike=bool(pop["owns_bike"][i]), owns_car=bool(pop["owns_car"][i])))
    with open(path, "w") as f:
        yaml.dump(agents, f, Dumper=_Dumper, default_flow_style=False)


def load_population(agents_path, network_path, scenario: Scenario, number_of_neighbour_links: int, seed: int = 0) -> PopulationData:
    """What `load_unlinked_agents_from_file` + `link_agents` leave in memory (simulation.rs:64-78), as mvb_population."""
    arrays = read_agents(agents_path, scenario)
    n = len(arrays["subculture"])
    arrays["social_rowptr"], arrays["social_col"] = read_network(network_path, n)
    arrays["neighbour_rowptr"], arrays["neighbour_col"] = neighbour_networks(
        arrays["neighbourhood"], len(scenario.neighbourhood_ids), number_of_neighbour_links, seed)
    return PopulationData(arrays)
