"""Python face of the C ABI (harness plumbing for tests and bench; the product is the .so).

`PopulationData` / `TablesData` hold numpy arrays in the layout of `mvb_population` /
`mvb_tables`; `Simulation` wraps an `mvb_sim*` handle.  Names follow the reference:
`Simulation.run` is the weekday loop of src/simulation.rs:95-136, `step` one weekday
(:116-128), `stats_row` the CSV row of :220-268 without the Day and Rain columns.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _ffi

CAR, PUBLIC_TRANSPORT, CYCLE, WALK = 0, 1, 2, 3          # src/transport_mode.rs:3-8
LOCAL_COMMUTE, CITY_COMMUTE, DISTANT_COMMUTE = 0, 1, 2   # src/journey_type.rs:5-9
GOOD, BAD = 0, 1                                         # src/weather.rs:7-10
STATS_FIXED = 7

# commute-distance mixture (mean, sd, weight): data/Travel To Work.ipynb cell 14 — the values
# config/parameters.yaml's empty `distributions: []` should hold (SURVEY.md §5 defect 1)
DEFAULT_GMM = ((9845.09, 3691.81, 0.66987), (2205.45, 1414.38, 0.25230), (22606.84, 11965.02, 0.07782))


class MotivateError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise MotivateError(rc, (_ffi.load().mvb_last_error() or b"").decode())


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


_POP_FIELDS = (
    ("subculture", np.uint8, _ffi.u8p), ("neighbourhood", np.uint16, _ffi.u16p), ("commute", np.uint8, _ffi.u8p),
    ("owns_car", np.uint8, _ffi.u8p), ("owns_bike", np.uint8, _ffi.u8p),
    ("weather_sensitivity", np.float32, _ffi.f32p),
    ("consistency", np.float32, _ffi.f32p), ("social_connectivity", np.float32, _ffi.f32p),
    ("subculture_connectivity", np.float32, _ffi.f32p), ("neighbourhood_connectivity", np.float32, _ffi.f32p),
    ("average_weight", np.float32, _ffi.f32p),
    ("habit", np.float32, _ffi.f32p), ("current_mode", np.uint8, _ffi.u8p), ("norm", np.uint8, _ffi.u8p),
    ("social_rowptr", np.uint64, _ffi.u64p), ("social_col", np.uint32, _ffi.u32p),
    ("neighbour_rowptr", np.uint64, _ffi.u64p), ("neighbour_col", np.uint32, _ffi.u32p),
)
_OPTIONAL = {"consistency", "social_connectivity", "subculture_connectivity", "neighbourhood_connectivity",
             "average_weight"}


@dataclass
class PopulationData:
    """Structure-of-arrays population (mvb_population).  Arrays are owned numpy arrays."""
    arrays: dict
    _struct: Optional[_ffi.Population] = field(default=None, repr=False)

    def __post_init__(self):
        fixed = {}
        for name, dt, _ in _POP_FIELDS:
            v = self.arrays.get(name)
            if v is None:
                if name not in _OPTIONAL:
                    raise ValueError(f"population field {name} is required")
                fixed[name] = None
            else:
                fixed[name] = np.ascontiguousarray(v, dtype=dt)
        self.arrays = fixed

    @property
    def n_agents(self):
        return int(self.arrays["subculture"].shape[0])

    def __getitem__(self, k):
        return self.arrays[k]

    def get(self, k, default=None):
        v = self.arrays.get(k)
        return default if v is None else v

    def struct(self) -> _ffi.Population:
        if self._struct is None:
            s = _ffi.Population()
            s.n_agents = self.n_agents
            for name, _, typ in _POP_FIELDS:
                setattr(s, name, _ptr(self.arrays[name], typ))
            self._struct = s
        return self._struct

    def copy(self):
        return PopulationData({k: (None if v is None else v.copy()) for k, v in self.arrays.items()})


@dataclass
class TablesData:
    """Scenario tables (mvb_tables): supportiveness[H][4], capacity[H][4], desirability[S][4]."""
    supportiveness: np.ndarray
    capacity: np.ndarray
    desirability: np.ndarray
    _struct: Optional[_ffi.Tables] = field(default=None, repr=False)

    def __post_init__(self):
        self.supportiveness = np.ascontiguousarray(self.supportiveness, np.float32).reshape(-1, 4)
        self.capacity = np.ascontiguousarray(self.capacity, np.uint32).reshape(-1, 4)
        self.desirability = np.ascontiguousarray(self.desirability, np.float32).reshape(-1, 4)
        if self.supportiveness.shape != self.capacity.shape:
            raise ValueError("supportiveness and capacity must both be [H][4]")

    @property
    def H(self):
        return self.supportiveness.shape[0]

    @property
    def S(self):
        return self.desirability.shape[0]

    def __getitem__(self, k):
        return getattr(self, k)

    def struct(self) -> _ffi.Tables:
        if self._struct is None:
            t = _ffi.Tables()
            t.n_subcultures, t.n_neighbourhoods = self.S, self.H
            t.supportiveness = _ptr(self.supportiveness, _ffi.f32p)
            t.capacity = _ptr(self.capacity, _ffi.u32p)
            t.desirability = _ptr(self.desirability, _ffi.f32p)
            self._struct = t
        return self._struct

    def copy(self):
        return TablesData(self.supportiveness.copy(), self.capacity.copy(), self.desirability.copy())


def make_params(consistency=1.0, social_connectivity=0.7, subculture_connectivity=0.5,
                neighbourhood_connectivity=0.3, days_in_habit_average=10, average_weight=None) -> _ffi.Params:
    """Uniform weights (config/parameters.yaml:2,4,9,10; src/agent_generation.rs:226,246).
    average_weight = 2.0f32 / (days as f32 + 1.0f32) unless given."""
    p = _ffi.Params()
    p.consistency = consistency
    p.social_connectivity = social_connectivity
    p.subculture_connectivity = subculture_connectivity
    p.neighbourhood_connectivity = neighbourhood_connectivity
    p.average_weight = (float(np.float32(2.0) / (np.float32(days_in_habit_average) + np.float32(1.0)))
                        if average_weight is None else average_weight)
    p.flags = 0
    return p


def params_dict(p: _ffi.Params) -> dict:
    return {k: getattr(p, k) for k in ("consistency", "social_connectivity", "subculture_connectivity",
                                       "neighbourhood_connectivity", "average_weight")}


@dataclass
class InterventionData:
    """Resolved intervention (mvb_intervention): the tables / ownership that hold from `day` on."""
    day: int
    tables: Optional[TablesData] = None
    owns_car: Optional[np.ndarray] = None
    owns_bike: Optional[np.ndarray] = None
    _struct: Optional[_ffi.Intervention] = field(default=None, repr=False)

    def struct(self) -> _ffi.Intervention:
        if self._struct is None:
            iv = _ffi.Intervention()
            iv.day = self.day
            iv.tables = C.pointer(self.tables.struct()) if self.tables is not None else None
            if self.owns_car is not None:
                self.owns_car = np.ascontiguousarray(self.owns_car, np.uint8)
            if self.owns_bike is not None:
                self.owns_bike = np.ascontiguousarray(self.owns_bike, np.uint8)
            iv.owns_car = _ptr(self.owns_car, _ffi.u8p)
            iv.owns_bike = _ptr(self.owns_bike, _ffi.u8p)
            self._struct = iv
        return self._struct


def synth_population(n_agents, n_subcultures=4, n_neighbourhoods=10, social_links_min=5,
                     neighbour_links_min=10, car_ratio=64926 / 111166, bike_ratio=None, seed=0,
                     gmm=DEFAULT_GMM) -> PopulationData:
    """Seeded synthetic population (mvb_synth_create).  Bikes default to the number of
    cars, reproducing src/agent_generation.rs:160."""
    lib = _ffi.load()
    sp = _synth_spec(n_agents, n_subcultures, n_neighbourhoods, social_links_min, neighbour_links_min, car_ratio, bike_ratio, seed, gmm)
    h = C.c_void_p()
    _check(lib.mvb_synth_create(C.byref(sp), C.byref(h)))
    try:
        return _population_from_struct(lib.mvb_synth_population(h).contents)
    finally:
        lib.mvb_synth_destroy(h)


def _synth_spec(n_agents, n_subcultures, n_neighbourhoods, social_links_min, neighbour_links_min, car_ratio, bike_ratio, seed, gmm):
    sp = _ffi.SynthSpec()
    sp.n_agents, sp.n_subcultures, sp.n_neighbourhoods = n_agents, n_subcultures, n_neighbourhoods
    sp.social_links_min, sp.neighbour_links_min = social_links_min, neighbour_links_min
    sp.n_cars = int(round(car_ratio * n_agents))
    sp.n_bikes = sp.n_cars if bike_ratio is None else int(round(bike_ratio * n_agents))
    sp.seed = seed
    sp.n_gmm = len(gmm)
    for g, row in enumerate(gmm):
        for k in range(3):
            sp.gmm[g][k] = row[k]
    return sp


def _population_from_struct(p) -> "PopulationData":
    N = p.n_agents

    def arr(ptr, n, dt):
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)

    s_rowptr = arr(p.social_rowptr, N + 1, np.uint64)
    n_rowptr = arr(p.neighbour_rowptr, N + 1, np.uint64)
    return PopulationData(dict(
        subculture=arr(p.subculture, N, np.uint8), neighbourhood=arr(p.neighbourhood, N, np.uint16),
        commute=arr(p.commute, N, np.uint8), owns_car=arr(p.owns_car, N, np.uint8),
        owns_bike=arr(p.owns_bike, N, np.uint8), weather_sensitivity=arr(p.weather_sensitivity, N, np.float32),
        habit=arr(p.habit, 4 * N, np.float32).reshape(N, 4), current_mode=arr(p.current_mode, N, np.uint8),
        norm=arr(p.norm, N, np.uint8), social_rowptr=s_rowptr,
        social_col=arr(p.social_col, int(s_rowptr[N]), np.uint32), neighbour_rowptr=n_rowptr,
        neighbour_col=arr(p.neighbour_col, int(n_rowptr[N]), np.uint32)))


class DevicePopulation:
    """A synthetic population generated on the GPU and living in its memory (mvb_synth_create_device): pass it to
    Simulation / Simulation.partitioned like a PopulationData (nothing crosses PCIe); to_host() copies it out."""

    def __init__(self, n_agents, n_subcultures=4, n_neighbourhoods=10, social_links_min=5, neighbour_links_min=10,
                 car_ratio=64926 / 111166, bike_ratio=None, seed=0, gmm=DEFAULT_GMM, device=0):
        self._lib = _ffi.load()
        sp = _synth_spec(n_agents, n_subcultures, n_neighbourhoods, social_links_min, neighbour_links_min, car_ratio,
                         bike_ratio, seed, gmm)
        self._h = C.c_void_p()
        _check(self._lib.mvb_synth_create_device(C.byref(sp), device, C.byref(self._h)))
        self.device = device

    @property
    def n_agents(self):
        return int(self._lib.mvb_synth_population(self._h).contents.n_agents)

    def struct(self) -> _ffi.Population:
        return self._lib.mvb_synth_population(self._h).contents          # device pointers

    def to_host(self) -> "PopulationData":
        h = C.c_void_p()
        _check(self._lib.mvb_synth_copy_to_host(self._h, C.byref(h)))
        try:
            return _population_from_struct(self._lib.mvb_synth_population(h).contents)
        finally:
            self._lib.mvb_synth_destroy(h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mvb_synth_destroy(self._h)
            self._h = None

    __del__ = close


def synth_tables(n_subcultures=4, n_neighbourhoods=10, n_agents=1_000_000, seed=0) -> TablesData:
    """Synthetic scenario tables of SURVEY.md §8(d): desirability ~U[0.4,0.9], supportiveness ~U[0.25,0.75],
    Car/PT capacity = shipped magnitudes (config/scenario.yaml:45-49, ~3725 / ~5548 for 5294 residents)
    scaled to the residents per neighbourhood, Cycle/Walk = 1.5*N (never congested)."""
    rng = np.random.default_rng(seed + 7919)
    des = rng.uniform(0.4, 0.9, (n_subcultures, 4)).astype(np.float32)
    sup = rng.uniform(0.25, 0.75, (n_neighbourhoods, 4)).astype(np.float32)
    scale = (n_agents / n_neighbourhoods) / 5294.0
    cap = np.zeros((n_neighbourhoods, 4), np.uint32)
    cap[:, CAR] = np.maximum(1, rng.uniform(2000, 6600, n_neighbourhoods) * scale).astype(np.uint32)
    cap[:, PUBLIC_TRANSPORT] = np.maximum(1, rng.uniform(90, 6900, n_neighbourhoods) * scale).astype(np.uint32)
    cap[:, CYCLE] = cap[:, WALK] = np.uint32(min(int(1.5 * n_agents), 0xFFFFFFFF))
    return TablesData(sup, cap, des)


def weather_pattern(n_days, seed=0, p_rain_day0=0.14, p_good_given_good=0.886, p_good_given_bad=0.699):
    """Two-state Markov chain with the matrix of src/main.rs:73-85 (src/weather.rs:19-50) from numpy's seeded PRNG:
    benchmarks and tests that want several different sequences.  The rain sequence of a reference run (StdRng =
    HC-128, zero seed) is `motivate_b200.weather.make_pattern`."""
    rng = np.random.default_rng(seed + 104729)
    u = rng.random(n_days)
    w = np.zeros(n_days, np.uint8)
    w[0] = GOOD if u[0] > p_rain_day0 else BAD
    for i in range(1, n_days):
        pg = p_good_given_good if w[i - 1] == GOOD else p_good_given_bad
        w[i] = GOOD if u[i] < pg else BAD
    return w


def run_rows(n_days):
    return 1 + sum(1 for d in range(1, n_days) if d % 7 < 5)


def comm_unique_id() -> bytes:
    """128-byte NCCL id for the agent-partitioned mode: rank 0 makes it, every rank passes it to Simulation.partitioned."""
    buf = C.create_string_buffer(128)
    _check(_ffi.load().mvb_comm_unique_id(buf))
    return buf.raw


def partition_plan(pop: PopulationData, n_subcultures: int, n_neighbourhoods: int, world: int) -> np.ndarray:
    """Which rank would update each agent if `world` B200s shared this population (no device needed)."""
    out = np.zeros(pop.n_agents, np.uint32)
    _check(_ffi.load().mvb_partition_plan(C.byref(pop.struct()), n_subcultures, n_neighbourhoods, world, _ptr(out, _ffi.u32p)))
    return out


class Simulation:
    """R replicas resident on one B200 (mvb_sim).  `pops`/`tabs` are sequences (or single objects)."""

    def __init__(self, pops, tabs, params: Optional[_ffi.Params] = None, device: int = 0):
        self._lib = _ffi.load()
        if isinstance(pops, (PopulationData, DevicePopulation)):
            pops = [pops]
        if isinstance(tabs, TablesData):
            tabs = [tabs] * len(pops)
        if len(pops) != len(tabs):
            raise ValueError("one tables object per population")
        self.pops, self.tabs = list(pops), list(tabs)
        self.params = params if params is not None else make_params()
        R = len(self.pops)
        pp = (_ffi.PopP * R)(*[C.pointer(p.struct()) for p in self.pops])
        tp = (_ffi.TabP * R)(*[C.pointer(t.struct()) for t in self.tabs])
        self._h = _ffi.SimP()
        _check(self._lib.mvb_create_batch(R, pp, tp, C.byref(self.params), device, C.byref(self._h)))
        self.n_replicas = R
        self.widths = [int(self._lib.mvb_stats_width(self._h, r)) for r in range(R)]

    @classmethod
    def partitioned(cls, pop: PopulationData, tab: TablesData, params, device: int, rank: int, world: int, unique_id: bytes):
        """One population shared by `world` GPUs (mvb_create_partitioned): one process per GPU, same inputs on every rank."""
        self = cls.__new__(cls)
        self._lib = _ffi.load()
        self.pops, self.tabs = [pop], [tab]
        self.params = params if params is not None else make_params()
        self._h = _ffi.SimP()
        idbuf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _check(self._lib.mvb_create_partitioned(C.byref(pop.struct()), C.byref(tab.struct()), C.byref(self.params), device,
                                                rank, world, idbuf, C.byref(self._h)))
        self.n_replicas = 1
        self.widths = [int(self._lib.mvb_stats_width(self._h, 0))]
        return self

    def exchange_mode(self) -> str:
        """How the ranks of an agent-partitioned handle exchange modes and counters after a weekday (mvb_exchange_mode):
        "p2p" = fused into the weekday kernel over peer memory, "nccl" = all-gather + all-reduce, "none" = not partitioned."""
        return {0: "none", 1: "nccl", 2: "p2p"}[int(self._lib.mvb_exchange_mode(self._h))]

    def partition_owner(self):
        out = np.zeros(self.pops[0].n_agents, np.uint32)
        _check(self._lib.mvb_partition_owner(self._h, _ptr(out, _ffi.u32p)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mvb_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def step(self, weather_today, weather_changed):
        _check(self._lib.mvb_step(self._h, int(weather_today), int(bool(weather_changed))))

    def _ivs(self, interventions):
        if interventions is None:
            return None, None
        if isinstance(interventions, InterventionData):
            interventions = [interventions] * self.n_replicas
        keep = [iv.struct() if iv is not None else None for iv in interventions]
        arr = (_ffi.IvP * self.n_replicas)(*[C.pointer(k) if k is not None else None for k in keep])
        return arr, keep

    def run(self, weather, n_days=None, interventions=None):
        """Weekday loop of src/simulation.rs:95-136.  Returns (days[rows], [rows_r[rows][width_r] per replica])."""
        weather = np.ascontiguousarray(weather, np.uint8)
        n_days = len(weather) if n_days is None else n_days
        rows = run_rows(n_days)
        days = np.zeros(rows, np.uint32)
        out = np.zeros(rows * sum(self.widths), np.uint32)
        ivs, _keep = self._ivs(interventions)
        _check(self._lib.mvb_run(self._h, _ptr(weather, _ffi.u8p), n_days, ivs, _ptr(days, _ffi.u32p),
                                 _ptr(out, _ffi.u32p)))
        per, o = [], 0
        for w in self.widths:
            per.append(out[o:o + rows * w].reshape(rows, w).copy())
            o += rows * w
        return days, per

    def run_async(self, weather, first_day, end_day, interventions=None):
        weather = np.ascontiguousarray(weather, np.uint8)
        ivs, _keep = self._ivs(interventions)
        _check(self._lib.mvb_run_async(self._h, _ptr(weather, _ffi.u8p), first_day, end_day, ivs))

    def fetch_rows(self):
        n = C.c_uint32()
        _check(self._lib.mvb_fetch_rows(self._h, C.byref(n), None, None, 0))
        rows = n.value
        days = np.zeros(rows, np.uint32)
        out = np.zeros(rows * sum(self.widths), np.uint32)
        _check(self._lib.mvb_fetch_rows(self._h, C.byref(n), _ptr(days, _ffi.u32p), _ptr(out, _ffi.u32p), out.size))
        per, o = [], 0
        for w in self.widths:
            per.append(out[o:o + rows * w].reshape(rows, w).copy())
            o += rows * w
        return days, per

    def reset_rows(self):
        _check(self._lib.mvb_reset_rows(self._h))

    def sync(self):
        _check(self._lib.mvb_sync(self._h))

    @property
    def stream(self):
        return self._lib.mvb_stream(self._h)

    def last_run_stats(self):
        ms, n = C.c_float(), C.c_uint64()
        _check(self._lib.mvb_last_run_stats(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_tables(self, replica, tab: TablesData):
        _check(self._lib.mvb_set_tables(self._h, replica, C.byref(tab.struct())))
        self.tabs[replica] = tab

    def set_ownership(self, replica, owns_car=None, owns_bike=None):
        car = None if owns_car is None else np.ascontiguousarray(owns_car, np.uint8)
        bike = None if owns_bike is None else np.ascontiguousarray(owns_bike, np.uint8)
        _check(self._lib.mvb_set_ownership(self._h, replica, _ptr(car, _ffi.u8p), _ptr(bike, _ffi.u8p)))

    def stats_row(self, replica=0):
        row = np.zeros(self.widths[replica], np.uint32)
        _check(self._lib.mvb_stats_row(self._h, replica, _ptr(row, _ffi.u32p)))
        return row

    def get_state(self, replica=0):
        N, H = self.pops[replica].n_agents, self.tabs[replica].H
        st = dict(habit=np.zeros((N, 4), np.float32), current_mode=np.zeros(N, np.uint8),
                  last_mode=np.zeros(N, np.uint8), norm=np.zeros(N, np.uint8),
                  hist=np.zeros((H, 4), np.uint32), congestion=np.zeros((H, 4), np.float32))
        _check(self._lib.mvb_get_state(self._h, replica, _ptr(st["habit"], _ffi.f32p),
                                       _ptr(st["current_mode"], _ffi.u8p), _ptr(st["last_mode"], _ffi.u8p),
                                       _ptr(st["norm"], _ffi.u8p), _ptr(st["hist"], _ffi.u32p),
                                       _ptr(st["congestion"], _ffi.f32p)))
        return st

    def algorithmic_bytes_per_day(self, replica=0):
        return int(self._lib.mvb_algorithmic_bytes_per_day(self._h, replica))
