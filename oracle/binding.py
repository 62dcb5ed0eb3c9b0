"""ctypes binding of the CPU oracle (oracle/motivate_oracle.c) and of its reference-shaped twin
(oracle/reference_shaped.cpp: same entry points over heap objects and per-call maps).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from motivate_b200 import _ffi
from motivate_b200.api import InterventionData, PopulationData, TablesData, _ptr, run_rows

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmotivate_oracle.so")
REFSHAPED_PATH = os.path.join(_HERE, "libmotivate_refshaped.so")
_libs = {}


def build(native=False, out=None, shape="dense"):
    """Compile the oracle with gcc (shape="reference": reference_shaped.cpp with g++).  native=True adds -O3
    -march=native (the README's intent for the reference's release build, README.md:89-92) for CPU-baseline timing
    on the box it runs on."""
    out = out or (LIB_PATH if shape == "dense" else REFSHAPED_PATH)
    flags = ["-O3", "-march=native"] if native else ["-O2", "-march=x86-64-v2"]
    if shape == "dense":
        cmd = ["gcc", *flags, "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=c11", "-shared",
               "-o", out, os.path.join(_HERE, "motivate_oracle.c")]
    else:
        cmd = ["g++", *flags, "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=c++17", "-shared",
               "-o", out, os.path.join(_HERE, "reference_shaped.cpp")]
    subprocess.run(cmd, check=True)
    return out


def load(native=False, shape="dense"):
    key = ("native" if native else "portable", shape)
    if key in _libs:
        return _libs[key]
    path = LIB_PATH if shape == "dense" else REFSHAPED_PATH
    if native:
        path = os.path.join(tempfile.gettempdir(), f"libmotivate_{shape}_native_{os.getpid()}.so")
        try:
            build(native=True, out=path, shape=shape)
        except Exception:
            return load(False, shape)
    elif not os.path.exists(path):
        build(shape=shape)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    P = C.c_void_p
    sig = {
        "mvo_create": (C.c_int, [_ffi.PopP, _ffi.TabP, _ffi.PrmP, C.POINTER(P)]),
        "mvo_destroy": (None, [P]),
        "mvo_step": (C.c_int, [P, C.c_uint8, C.c_uint8]),
        "mvo_run": (C.c_int, [P, _ffi.u8p, C.c_uint32, _ffi.IvP, _ffi.u32p, _ffi.u32p]),
        "mvo_run_rows": (C.c_uint32, [C.c_uint32]),
        "mvo_set_tables": (C.c_int, [P, _ffi.TabP]),
        "mvo_set_ownership": (C.c_int, [P, _ffi.u8p, _ffi.u8p]),
        "mvo_stats_row": (C.c_int, [P, _ffi.u32p]),
        "mvo_get_state": (C.c_int, [P, _ffi.f32p, _ffi.u8p, _ffi.u8p, _ffi.u8p, _ffi.u32p, _ffi.f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _libs[key] = lib
    return lib


SYNTH_PATH = os.path.join(_HERE, "libmotivate_synthhost.so")


def synth_population(n_agents, n_subcultures=4, n_neighbourhoods=10, social_links_min=5, neighbour_links_min=10,
                     car_ratio=64926 / 111166, bike_ratio=None, seed=0, gmm=None) -> PopulationData:
    """motivate_b200.synth_population (same seeded host generator, same output) WITHOUT loading the product library:
    oracle/libmotivate_synthhost.so is motivate_b200/csrc/synth.cpp compiled on its own (oracle/Makefile).  For
    bench.py's CPU arm, whose process must not depend on libmotivate_b200.so."""
    from motivate_b200.api import DEFAULT_GMM, _population_from_struct, _synth_spec
    if "synth" not in _libs:
        if not os.path.exists(SYNTH_PATH):
            subprocess.run(["make", "-C", _HERE, "libmotivate_synthhost.so"], check=True)
        lib = C.CDLL(SYNTH_PATH, mode=C.RTLD_LOCAL)
        lib.mvb_synth_create.restype, lib.mvb_synth_create.argtypes = C.c_int, [C.POINTER(_ffi.SynthSpec), C.POINTER(C.c_void_p)]
        lib.mvb_synth_population.restype, lib.mvb_synth_population.argtypes = _ffi.PopP, [C.c_void_p]
        lib.mvb_synth_destroy.restype, lib.mvb_synth_destroy.argtypes = None, [C.c_void_p]
        lib.mvb_last_error.restype = C.c_char_p
        _libs["synth"] = lib
    lib = _libs["synth"]
    sp = _synth_spec(n_agents, n_subcultures, n_neighbourhoods, social_links_min, neighbour_links_min, car_ratio, bike_ratio,
                     seed, gmm or DEFAULT_GMM)
    h = C.c_void_p()
    if lib.mvb_synth_create(C.byref(sp), C.byref(h)):
        raise RuntimeError((lib.mvb_last_error() or b"").decode())
    try:
        return _population_from_struct(lib.mvb_synth_population(h).contents)
    finally:
        lib.mvb_synth_destroy(h)


class OracleSimulation:
    """One replica on the CPU oracle; same method names as motivate_b200.Simulation.  shape="reference" runs the
    reference-shaped twin (heap objects, per-call maps) instead of the dense oracle."""

    def __init__(self, pop: PopulationData, tab: TablesData, params, native=False, shape="dense"):
        self._lib = load(native, shape)
        self.pop, self.tab, self.params = pop, tab, params
        self._h = C.c_void_p()
        rc = self._lib.mvo_create(C.byref(pop.struct()), C.byref(tab.struct()), C.byref(params), C.byref(self._h))
        if rc:
            raise RuntimeError(f"mvo_create failed: {rc}")
        self.width = 7 + tab.S + tab.H

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mvo_destroy(self._h)
            self._h = None

    __del__ = close

    def step(self, weather_today, weather_changed):
        self._lib.mvo_step(self._h, int(weather_today), int(bool(weather_changed)))

    def run(self, weather, n_days=None, intervention: InterventionData | None = None):
        weather = np.ascontiguousarray(weather, np.uint8)
        n_days = len(weather) if n_days is None else n_days
        rows = run_rows(n_days)
        days = np.zeros(rows, np.uint32)
        out = np.zeros((rows, self.width), np.uint32)
        iv = C.byref(intervention.struct()) if intervention is not None else None
        rc = self._lib.mvo_run(self._h, _ptr(weather, _ffi.u8p), n_days, iv, _ptr(days, _ffi.u32p), _ptr(out, _ffi.u32p))
        if rc:
            raise RuntimeError(f"mvo_run failed: {rc}")
        return days, out

    def set_tables(self, tab: TablesData):
        self._lib.mvo_set_tables(self._h, C.byref(tab.struct()))
        self.tab = tab

    def set_ownership(self, owns_car=None, owns_bike=None):
        car = None if owns_car is None else np.ascontiguousarray(owns_car, np.uint8)
        bike = None if owns_bike is None else np.ascontiguousarray(owns_bike, np.uint8)
        self._lib.mvo_set_ownership(self._h, _ptr(car, _ffi.u8p), _ptr(bike, _ffi.u8p))

    def stats_row(self):
        row = np.zeros(self.width, np.uint32)
        self._lib.mvo_stats_row(self._h, _ptr(row, _ffi.u32p))
        return row

    def get_state(self):
        N, H = self.pop.n_agents, self.tab.H
        st = dict(habit=np.zeros((N, 4), np.float32), current_mode=np.zeros(N, np.uint8),
                  last_mode=np.zeros(N, np.uint8), norm=np.zeros(N, np.uint8),
                  hist=np.zeros((H, 4), np.uint32), congestion=np.zeros((H, 4), np.float32))
        self._lib.mvo_get_state(self._h, _ptr(st["habit"], _ffi.f32p), _ptr(st["current_mode"], _ffi.u8p),
                                _ptr(st["last_mode"], _ffi.u8p), _ptr(st["norm"], _ffi.u8p),
                                _ptr(st["hist"], _ffi.u32p), _ptr(st["congestion"], _ffi.f32p))
        return st
