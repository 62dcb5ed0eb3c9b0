"""Pinning the oracle on a real run of the reference.  tests/golden/ref_*.npz come from tools/make_ref_fixtures.sh
(a Rust toolchain is needed; none exists in the build image, so normally there are none and the comparison tests
skip).  The dump format, the converter and the comparison are exercised on a dump written from the oracle itself."""
import glob
import os
import struct
import sys

import numpy as np
import pytest

import motivate_b200 as mb
from oracle.binding import OracleSimulation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ref_dump_to_npz import read_dump  # noqa: E402

FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_*.npz")))
NO_FIXTURES = "no tests/golden/ref_*.npz (needs a Rust toolchain: tools/make_ref_fixtures.sh)"
POP_KEYS = ("subculture", "neighbourhood", "commute", "owns_car", "owns_bike", "weather_sensitivity", "consistency",
            "social_connectivity", "subculture_connectivity", "neighbourhood_connectivity", "average_weight", "habit",
            "current_mode", "norm", "social_rowptr", "social_col", "neighbour_rowptr", "neighbour_col")


class OracleOne(OracleSimulation):
    def state(self):
        return self.get_state()

    def apply(self, tables, car, bike):
        self.set_tables(tables)
        self.set_ownership(car, bike)


class GpuOne(mb.Simulation):
    def state(self):
        return self.get_state(0)

    def apply(self, tables, car, bike):
        self.set_tables(0, tables)
        self.set_ownership(0, car, bike)


def replay(fx, make_sim):
    """Run the recorded inputs through `make_sim` and compare with the recorded states, every weekday.  Returns the
    number of agent-days compared."""
    pop = mb.PopulationData({k: fx[k] for k in POP_KEYS})
    tab = mb.TablesData(fx["supportiveness"], fx["capacity"], fx["desirability"])
    sim = make_sim(pop, tab, mb.make_params())             # per-agent weights come with the population
    w = fx["weather"]
    days = [int(d) for d in fx["days"]]
    assert days[0] == 0
    st = sim.state()
    assert np.array_equal(st["current_mode"], fx["day_current_mode"][0]) and np.array_equal(st["norm"], fx["day_norm"][0])
    prev, k, compared = w[0], 1, 0
    for day in range(1, days[-1] + 1):
        if "iv_day" in fx and day == int(fx["iv_day"]):
            sim.apply(mb.TablesData(fx["iv_supportiveness"], fx["iv_capacity"], fx["iv_desirability"]),
                      fx["iv_owns_car"], fx["iv_owns_bike"])
        if day % 7 >= 5:
            continue
        sim.step(w[day], prev != w[day])
        prev = w[day]
        assert days[k] == day
        st = sim.state()
        bad = np.flatnonzero(st["current_mode"] != fx["day_current_mode"][k])
        assert bad.size == 0, f"day {day}: {bad.size} agents chose another mode than the reference run (first: {bad[:5]})"
        assert np.array_equal(st["norm"], fx["day_norm"][k]), f"day {day}: norm"
        assert np.array_equal(st["habit"].view(np.uint32), fx["day_habit"][k].view(np.uint32)), f"day {day}: habit bits"
        compared += len(st["current_mode"])
        k += 1
    sim.close()
    return compared


@pytest.mark.skipif(not FIXTURES, reason=NO_FIXTURES)
@pytest.mark.parametrize("path", FIXTURES or ["none"], ids=os.path.basename)
def test_oracle_equals_reference_run(path):
    assert replay(dict(np.load(path)), OracleOne) > 0


@pytest.mark.gpu
@pytest.mark.skipif(not FIXTURES, reason=NO_FIXTURES)
@pytest.mark.parametrize("path", FIXTURES or ["none"], ids=os.path.basename)
def test_cuda_path_equals_reference_run(path):
    assert replay(dict(np.load(path)), GpuOne) > 0


def write_dump_like_dump_state_rs(d, pop, tab, w, iv, states):
    """The byte layout of rust/patch/dump_state.rs, written from Python."""
    os.makedirs(d, exist_ok=True)
    N = pop.n_agents
    prm = mb.make_params()
    with open(os.path.join(d, "inputs.bin"), "wb") as f:
        f.write(b"MVBD" + struct.pack("<5I", 1, N, tab.S, tab.H, len(w)))
        for k, dt in (("subculture", "u1"), ("neighbourhood", "<u2"), ("commute", "u1"), ("owns_car", "u1"), ("owns_bike", "u1"),
                      ("weather_sensitivity", "<f4")):
            f.write(np.asarray(pop[k], dt).tobytes())
        for k in ("consistency", "social_connectivity", "subculture_connectivity", "neighbourhood_connectivity", "average_weight"):
            f.write(np.full(N, getattr(prm, k), "<f4").tobytes())
        f.write(np.asarray(pop["habit"], "<f4").tobytes())
        f.write(np.asarray(pop["current_mode"], "u1").tobytes()); f.write(np.asarray(pop["norm"], "u1").tobytes())
        for g in ("social", "neighbour"):
            f.write(np.asarray(pop[g + "_rowptr"], "<u8").tobytes()); f.write(np.asarray(pop[g + "_col"], "<u4").tobytes())
        f.write(tab.supportiveness.astype("<f4").tobytes()); f.write(tab.capacity.astype("<u4").tobytes())
        f.write(tab.desirability.astype("<f4").tobytes()); f.write(np.asarray(w, "u1").tobytes())
    with open(os.path.join(d, "intervention.bin"), "wb") as f:
        f.write(struct.pack("<I", iv.day))
        f.write(iv.tables.supportiveness.astype("<f4").tobytes()); f.write(iv.tables.capacity.astype("<u4").tobytes())
        f.write(iv.tables.desirability.astype("<f4").tobytes())
        f.write(np.asarray(iv.owns_car, "u1").tobytes()); f.write(np.asarray(iv.owns_bike, "u1").tobytes())
    with open(os.path.join(d, "days.bin"), "wb") as f:
        for day, st in states:
            f.write(struct.pack("<I", day)); f.write(st["current_mode"].tobytes()); f.write(st["norm"].tobytes())
            f.write(st["habit"].astype("<f4").tobytes())


def test_dump_format_round_trip(tmp_path):
    """A dump in dump_state.rs's layout (written here from the oracle's own states) goes through the converter and
    the replay: proves the recipe's plumbing; it does NOT pin anything (oracle against itself)."""
    n = 600
    pop = mb.synth_population(n, 3, 5, 3, 4, seed=3)
    tab = mb.synth_tables(3, 5, n, seed=3)
    w = mb.weather_pattern(40, seed=3)
    tab2 = tab.copy()
    tab2.capacity[:, 0] //= 2
    iv = mb.InterventionData(day=20, tables=tab2, owns_car=1 - pop["owns_car"], owns_bike=pop["owns_bike"].copy())
    o = OracleSimulation(pop, tab, mb.make_params())
    states, prev = [(0, o.get_state())], w[0]
    for day in range(1, 40):
        if day == iv.day:
            o.set_tables(iv.tables); o.set_ownership(iv.owns_car, iv.owns_bike)
        if day % 7 < 5:
            o.step(w[day], prev != w[day]); prev = w[day]
            states.append((day, o.get_state()))
    write_dump_like_dump_state_rs(str(tmp_path / "1"), pop, tab, w, iv, states)
    fx = read_dump(str(tmp_path / "1"))
    assert int(fx["iv_day"]) == 20 and len(fx["days"]) == len(states)
    assert replay(fx, OracleOne) == n * (len(states) - 1)
    fx["day_current_mode"][5, 7] ^= 1                       # a corrupted record must be caught
    with pytest.raises(AssertionError, match="chose another mode"):
        replay(fx, OracleOne)
