"""ctypes declarations for include/motivate_b200.h and the loader of libmotivate_b200.so.

The library is built in-tree by `__graft_entry__.build()` / `make -C motivate_b200/csrc`.
There is no fallback: if the shared object is missing or does not load, importing the
product API raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MVB_LIB_PATH: a debug build of the same library (tools/launch_gaps.py uses libmotivate_b200_stamps.so)
LIB_PATH = os.environ.get("MVB_LIB_PATH") or os.path.join(_HERE, "libmotivate_b200.so")

u8p, u16p, u32p, u64p, f32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_uint32),
                               C.POINTER(C.c_uint64), C.POINTER(C.c_float))


class Population(C.Structure):
    _fields_ = [
        ("n_agents", C.c_uint32),
        ("subculture", u8p), ("neighbourhood", u16p), ("commute", u8p),
        ("owns_car", u8p), ("owns_bike", u8p), ("weather_sensitivity", f32p),
        ("consistency", f32p), ("social_connectivity", f32p), ("subculture_connectivity", f32p),
        ("neighbourhood_connectivity", f32p), ("average_weight", f32p),
        ("habit", f32p), ("current_mode", u8p), ("norm", u8p),
        ("social_rowptr", u64p), ("social_col", u32p),
        ("neighbour_rowptr", u64p), ("neighbour_col", u32p),
    ]


class Tables(C.Structure):
    _fields_ = [("n_subcultures", C.c_uint32), ("n_neighbourhoods", C.c_uint32),
                ("supportiveness", f32p), ("capacity", u32p), ("desirability", f32p)]


class Params(C.Structure):
    _fields_ = [("consistency", C.c_float), ("social_connectivity", C.c_float),
                ("subculture_connectivity", C.c_float), ("neighbourhood_connectivity", C.c_float),
                ("average_weight", C.c_float), ("flags", C.c_uint32)]


class Intervention(C.Structure):
    _fields_ = [("day", C.c_uint32), ("tables", C.POINTER(Tables)), ("owns_car", u8p), ("owns_bike", u8p)]


class SynthSpec(C.Structure):
    _fields_ = [("n_agents", C.c_uint32), ("n_subcultures", C.c_uint32), ("n_neighbourhoods", C.c_uint32),
                ("social_links_min", C.c_uint32), ("neighbour_links_min", C.c_uint32),
                ("n_cars", C.c_uint32), ("n_bikes", C.c_uint32), ("seed", C.c_uint64),
                ("n_gmm", C.c_uint32), ("gmm", (C.c_double * 3) * 8)]


PopP, TabP, PrmP, IvP = C.POINTER(Population), C.POINTER(Tables), C.POINTER(Params), C.POINTER(Intervention)
SimP = C.c_void_p

# name -> (restype, argtypes): every symbol include/motivate_b200.h declares
SIGNATURES = {
    "mvb_abi_version": (C.c_int, []),
    "mvb_device_count": (C.c_int, []),
    "mvb_last_error": (C.c_char_p, []),
    "mvb_create": (C.c_int, [PopP, TabP, PrmP, C.c_int, C.POINTER(SimP)]),
    "mvb_create_batch": (C.c_int, [C.c_uint32, C.POINTER(PopP), C.POINTER(TabP), PrmP, C.c_int, C.POINTER(SimP)]),
    "mvb_destroy": (None, [SimP]),
    "mvb_n_replicas": (C.c_uint32, [SimP]),
    "mvb_stats_width": (C.c_uint32, [SimP, C.c_uint32]),
    "mvb_step": (C.c_int, [SimP, C.c_uint8, C.c_uint8]),
    "mvb_run": (C.c_int, [SimP, u8p, C.c_uint32, C.POINTER(IvP), u32p, u32p]),
    "mvb_run_rows": (C.c_uint32, [C.c_uint32]),
    "mvb_run_async": (C.c_int, [SimP, u8p, C.c_uint32, C.c_uint32, C.POINTER(IvP)]),
    "mvb_fetch_rows": (C.c_int, [SimP, u32p, u32p, u32p, C.c_size_t]),
    "mvb_reset_rows": (C.c_int, [SimP]),
    "mvb_sync": (C.c_int, [SimP]),
    "mvb_stream": (C.c_void_p, [SimP]),
    "mvb_last_run_stats": (C.c_int, [SimP, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "mvb_set_tables": (C.c_int, [SimP, C.c_uint32, TabP]),
    "mvb_set_ownership": (C.c_int, [SimP, C.c_uint32, u8p, u8p]),
    "mvb_stats_row": (C.c_int, [SimP, C.c_uint32, u32p]),
    "mvb_get_state": (C.c_int, [SimP, C.c_uint32, f32p, u8p, u8p, u8p, u32p, f32p]),
    "mvb_algorithmic_bytes_per_day": (C.c_uint64, [SimP, C.c_uint32]),
    "mvb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mvb_create_partitioned": (C.c_int, [PopP, TabP, PrmP, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(SimP)]),
    "mvb_partition_owner": (C.c_int, [SimP, u32p]),
    "mvb_partition_plan": (C.c_int, [PopP, C.c_uint32, C.c_uint32, C.c_uint32, u32p]),
    "mvb_exchange_mode": (C.c_int, [SimP]),
    "mvb_synth_create": (C.c_int, [C.POINTER(SynthSpec), C.POINTER(C.c_void_p)]),
    "mvb_synth_create_device": (C.c_int, [C.POINTER(SynthSpec), C.c_int, C.POINTER(C.c_void_p)]),
    "mvb_synth_copy_to_host": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mvb_synth_destroy": (None, [C.c_void_p]),
    "mvb_synth_population": (PopP, [C.c_void_p]),
    "mvb_synth_ba_network": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(u64p), C.POINTER(u32p)]),
    "mvb_free": (None, [C.c_void_p]),
}

_lib = None


def load():
    """Load libmotivate_b200.so (once) and type its entry points.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built.  Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C motivate_b200/csrc`.  There is no CPU fallback for the product path.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib
