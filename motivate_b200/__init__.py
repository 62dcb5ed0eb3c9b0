"""motivate_b200 — B200-native per-day agent update of Motivate behind a C ABI.

The product is `libmotivate_b200.so` (hand-written sm_100a CUDA, include/motivate_b200.h).
This package is the thin ctypes face of that ABI used by the tests and bench; importing
`motivate_b200.api` loads the library and raises if it has not been built.
"""
from .api import (BAD, CAR, CITY_COMMUTE, CYCLE, DEFAULT_GMM, DISTANT_COMMUTE, GOOD, LOCAL_COMMUTE,
                  PUBLIC_TRANSPORT, STATS_FIXED, WALK, DevicePopulation, InterventionData, MotivateError, PopulationData,
                  Simulation, TablesData, comm_unique_id, make_params, partition_plan, params_dict, run_rows, synth_population,
                  synth_tables, weather_pattern)
from .weather import make_pattern

__all__ = [n for n in dir() if not n.startswith("_")]
