"""Shared helpers for the tests: tiny hand-built populations and comparison utilities."""
from __future__ import annotations

import numpy as np

import motivate_b200 as mb

CAR, PT, CYCLE, WALK = 0, 1, 2, 3
LOCAL, CITY, DISTANT = 0, 1, 2
GOOD, BAD = 0, 1


def csr(lists, n):
    rowptr = np.zeros(n + 1, np.uint64)
    col = []
    for i in range(n):
        col.extend(lists[i] if i < len(lists) else [])
        rowptr[i + 1] = len(col)
    return rowptr, np.array(col, np.uint32)


def tiny_population(modes, social, neighbours, *, sub=None, nbhd=None, commute=None, car=None, bike=None,
                    ws=None, habit=None, norm=None, **per_agent):
    """Population from python lists; habit defaults to {mode: 1.0} (agent_generation.rs:198-200)."""
    n = len(modes)
    z = lambda v, d, dt: np.array(v if v is not None else [d] * n, dt)
    modes = np.array(modes, np.uint8)
    if habit is None:
        habit = np.zeros((n, 4), np.float32)
        habit[np.arange(n), modes] = 1.0
    s_rp, s_col = csr(social, n)
    n_rp, n_col = csr(neighbours, n)
    arrays = dict(subculture=z(sub, 0, np.uint8), neighbourhood=z(nbhd, 0, np.uint16), commute=z(commute, LOCAL, np.uint8),
                  owns_car=z(car, 1, np.uint8), owns_bike=z(bike, 1, np.uint8),
                  weather_sensitivity=z(ws, 0.0, np.float32), habit=np.array(habit, np.float32),
                  current_mode=modes, norm=np.array(norm if norm is not None else modes, np.uint8),
                  social_rowptr=s_rp, social_col=s_col, neighbour_rowptr=n_rp, neighbour_col=n_col)
    for k, v in per_agent.items():
        arrays[k] = np.array(v, np.float32)
    return mb.PopulationData(arrays)


def tables(supp, cap, des):
    return mb.TablesData(np.array(supp, np.float32), np.array(cap, np.uint32), np.array(des, np.float32))


def tab_dict(t):
    return {"supportiveness": t.supportiveness, "capacity": t.capacity, "desirability": t.desirability}


def random_population(n, S, H, seed, *, per_agent=False, max_deg=6, distant_share=0.15, dup_links=True):
    """Unstructured random population for oracle-vs-model tests: random link lists WITH duplicates and
    empty lists, random habits summing to <= 1, all commute classes, all ownership combinations."""
    rng = np.random.default_rng(seed)
    sub = rng.integers(0, S, n)
    nb = rng.integers(0, H, n)
    commute = rng.choice([LOCAL, CITY, DISTANT], n, p=[0.45, 0.55 - distant_share, distant_share])
    car, bike = rng.integers(0, 2, n), rng.integers(0, 2, n)
    modes = rng.integers(0, 4, n)
    social = [list(rng.integers(0, n, rng.integers(0, max_deg + 1))) for _ in range(n)]
    neigh = []
    for i in range(n):
        same = np.flatnonzero(nb == nb[i])
        k = rng.integers(0, max_deg + 1)
        neigh.append(list(rng.choice(same, k, replace=dup_links)) if len(same) and k else [])
    habit = rng.dirichlet(np.ones(4), n).astype(np.float32) * rng.uniform(0.2, 1.0, (n, 1)).astype(np.float32)
    habit[rng.random((n, 4)) < 0.3] = 0.0           # sparse keys like the reference's HashMap
    extra = {}
    if per_agent:
        extra = dict(consistency=rng.uniform(0.5, 1.0, n), social_connectivity=rng.uniform(0.1, 0.9, n),
                     subculture_connectivity=rng.uniform(0.1, 0.9, n),
                     neighbourhood_connectivity=rng.uniform(0.1, 0.9, n), average_weight=rng.uniform(0.05, 0.5, n))
    return tiny_population(modes, social, neigh, sub=sub, nbhd=nb, commute=commute, car=car, bike=bike,
                           ws=rng.random(n), habit=habit, norm=rng.integers(0, 4, n), **extra)


def random_tables(S, H, n, seed, tight=True):
    """Tables with ties on purpose (quantised values) and capacities small enough to congest."""
    rng = np.random.default_rng(seed + 1)
    des = rng.integers(2, 10, (S, 4)) / np.float32(10)
    sup = rng.integers(2, 9, (H, 4)) / np.float32(10)
    per = max(1, n // H)
    cap = rng.integers(0 if tight else per, max(2, per // (2 if tight else 1)) + per * (0 if tight else 2), (H, 4))
    return tables(sup, cap, des)


def assert_state_equal(a, b, what=""):
    for k in ("current_mode", "last_mode", "norm", "hist"):
        assert np.array_equal(a[k], b[k]), f"{what}: {k} differs ({np.sum(a[k] != b[k])} entries)"
    assert np.array_equal(a["habit"].view(np.uint32), b["habit"].view(np.uint32)), f"{what}: habit bits differ"
    assert np.array_equal(a["congestion"].view(np.uint32), b["congestion"].view(np.uint32)), f"{what}: congestion bits differ"
