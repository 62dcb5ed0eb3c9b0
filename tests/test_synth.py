"""The seeded synthetic generator follows the reference's setup distributions (agent_generation.rs, social_network.rs)."""
import ctypes as C

import numpy as np
import pytest

import motivate_b200 as mb
from motivate_b200 import _ffi


def ba(min_links, n, seed=0):
    lib = _ffi.load()
    rp, cl = _ffi.u64p(), _ffi.u32p()
    rc = lib.mvb_synth_ba_network(min_links, n, seed, C.byref(rp), C.byref(cl))
    assert rc == 0, lib.mvb_last_error()
    rowptr = np.ctypeslib.as_array(rp, shape=(n + 1,)).copy()
    col = np.ctypeslib.as_array(cl, shape=(int(rowptr[n]),)).copy() if rowptr[n] else np.zeros(0, np.uint32)
    lib.mvb_free(rp); lib.mvb_free(cl)
    return rowptr, col


def test_ba_network_shape():
    m, n = 5, 20000
    rowptr, col = ba(m, n, 7)
    deg = np.diff(rowptr).astype(np.int64)
    # clique on the first m nodes, then m draws per node: m(m-1)/2 + (n-m)*m links, each stored twice
    assert rowptr[n] == 2 * (m * (m - 1) // 2 + (n - m) * m)
    assert deg.min() >= m - 1 and deg[m:].min() >= m          # every later node has at least its own m links
    assert abs(deg.mean() - 2 * m) < 0.01                     # "mean number of links / 2" (main.rs:199-204)
    assert deg.max() > 20 * m                                 # heavy tail
    # symmetric as a multiset: every (a,b) has a matching (b,a)
    src = np.repeat(np.arange(n, dtype=np.int64), deg)
    a = np.sort(src * n + col.astype(np.int64)); b = np.sort(col.astype(np.int64) * n + src)
    assert np.array_equal(a, b)
    # duplicates are kept (draws with replacement, social_network.rs:37-41)
    assert len(np.unique(a)) < len(a)
    # deterministic in the seed
    r2, c2 = ba(m, n, 7)
    assert np.array_equal(rowptr, r2) and np.array_equal(col, c2)
    assert not np.array_equal(col, ba(m, n, 8)[1])


def test_ba_rejects_min_1():
    lib = _ffi.load()
    rp, cl = _ffi.u64p(), _ffi.u32p()
    assert lib.mvb_synth_ba_network(1, 100, 0, C.byref(rp), C.byref(cl)) == -5      # MVB_ERR_DOMAIN


def test_population_distributions():
    n, S, H = 60000, 4, 10
    pop = mb.synth_population(n, S, H, 5, 10, seed=3)
    assert pop.n_agents == n
    cars, bikes = int(pop["owns_car"].sum()), int(pop["owns_bike"].sum())
    assert cars == round(64926 / 111166 * n) and bikes == cars         # agent_generation.rs:152,160 (bikes use number_of_cars)
    assert np.bincount(pop["subculture"], minlength=S).min() > 0.9 * n / S
    assert np.bincount(pop["neighbourhood"], minlength=H).min() > 0.9 * n / H
    cm = np.bincount(pop["commute"], minlength=3) / n
    assert 0.15 < cm[0] < 0.35 and 0.55 < cm[1] < 0.8 and 0.02 < cm[2] < 0.12
    # initial mode respects ownership (agent_generation.rs:286-325); habit one-hot, norm = mode
    m = pop["current_mode"]
    assert not np.any((m == 0) & (pop["owns_car"] == 0)) and not np.any((m == 2) & (pop["owns_bike"] == 0))
    both = (pop["owns_car"] == 1) & (pop["owns_bike"] == 1)
    assert abs(np.mean(m[both] == 0) - 0.40) < 0.02 and abs(np.mean(m[both] == 2) - 0.30) < 0.02
    assert np.array_equal(pop["norm"], m)
    assert np.array_equal(pop["habit"].argmax(1), m) and np.all(pop["habit"].sum(1) == 1.0)
    # neighbour links stay inside the neighbourhood (simulation.rs:165-173)
    nb = pop["neighbourhood"]
    src = np.repeat(np.arange(n), np.diff(pop["neighbour_rowptr"]).astype(np.int64))
    assert np.array_equal(nb[src], nb[pop["neighbour_col"]])
    assert abs(len(pop["social_col"]) / n - 10) < 0.01 and abs(len(pop["neighbour_col"]) / n - 20) < 0.05
    assert 0 <= pop["weather_sensitivity"].min() and pop["weather_sensitivity"].max() < 1
