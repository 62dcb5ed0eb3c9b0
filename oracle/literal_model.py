"""literal_model.py — a second, structure-preserving restatement of the reference's per-day update.

TEST INFRASTRUCTURE ONLY (see oracle/motivate_oracle.c).  Where the C oracle is a *dense*
restatement (4-lane arrays, absent key == 0.0), this file keeps the reference's own shape:
sparse dict "HashMaps", `union_of` / `intersection_of` folds, `into_group_map`, stable sorts
and `max_by`, agent objects holding lists of friend objects.  It exists to check that the
dense oracle and the sparse reference code are the same function, and to generate the golden
vectors under tests/golden/ (tests/golden/make_golden.py).  All reals are numpy.float32 scalars
so every operation rounds exactly like Rust's f32.

The only freedom the reference leaves is HashMap iteration order (std RandomState).  `ORDER`
below is the iteration order used for every map; the canonical order is the enum order
Car, PublicTransport, Cycle, Walk (SURVEY.md §8 a-T).  Tests also run other orders to show
that results differ only where exact f32 ties exist.

Citations are into /root/reference/src.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
CAR, PT, CYCLE, WALK = 0, 1, 2, 3          # transport_mode.rs:3-8
LOCAL, CITY, DISTANT = 0, 1, 2             # journey_type.rs:5-9
GOOD, BAD = 0, 1                           # weather.rs:7-10

CANONICAL_ORDER = (CAR, PT, CYCLE, WALK)


class MapOrder:
    """Iteration order applied to every 'HashMap' (keys absent from `order` go last)."""

    def __init__(self, order=CANONICAL_ORDER):
        self.order = tuple(order)
        self.pos = {k: i for i, k in enumerate(self.order)}

    def items(self, d):
        return sorted(d.items(), key=lambda kv: self.pos.get(kv[0], 99))


def union_of(a, b, f):
    """hashmap_union 0.2.0 `union_of`: all keys of a∪b; f(va, vb) where both present."""
    out = {}
    for k, v in a.items():
        out[k] = f(v, b[k]) if k in b else v
    for k, v in b.items():
        if k not in a:
            out[k] = v
    return out


def intersection_of(a, b, f):
    """hashmap_union 0.2.0 `intersection_of`: keys in both."""
    return {k: f(v, b[k]) for k, v in a.items() if k in b}


def stable_sorted(items, key_lt):
    """`sorted_by` (itertools 0.7.8 = collect + slice::sort_by, stable): insertion sort keeps
    the incoming order among elements for which neither is strictly before the other."""
    out = []
    for it in items:
        j = len(out)
        while j > 0 and key_lt(it, out[j - 1]):
            j -= 1
        out.insert(j, it)
    return out


def journey_cost(commute):
    """JourneyType::cost, journey_type.rs:16-35."""
    if commute == LOCAL:
        return {CAR: f32(0.2), PT: f32(0.2), CYCLE: f32(0.2), WALK: f32(0.2)}
    if commute == CITY:
        return {CAR: f32(0.3), PT: f32(0.3), CYCLE: f32(0.6), WALK: f32(0.9)}
    return {CAR: f32(0.1), PT: f32(0.1)}


class Neighbourhood:
    """neighbourhood.rs:13-31."""

    def __init__(self, supportiveness, capacity):
        self.supportiveness = dict(supportiveness)      # mode -> f32
        self.capacity = dict(capacity)                  # mode -> int
        self.congestion_modifier = {m: f32(1.0) for m in CANONICAL_ORDER}   # :34-41
        self.residents = []

    def update_congestion_modifier(self):
        """neighbourhood.rs:57-89."""
        groups = {}
        for a in self.residents:
            groups.setdefault(a.last_mode, []).append(a)
        new = {}
        for mode, grouped in groups.items():
            count = len(grouped)
            cap = self.capacity[mode]
            if count <= cap:
                new[mode] = f32(1.0)
            else:
                maximum_excess = len(self.residents) - cap
                actual_excess = count - cap
                new[mode] = f32(1.0) - (f32(actual_excess) / f32(maximum_excess))
        self.congestion_modifier = new


class Subculture:
    """subculture.rs:9-14."""

    def __init__(self, desirability):
        self.desirability = dict(desirability)


class Agent:
    """agent.rs:16-87."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.social_network = []
        self.neighbours = []
        self.last_mode = self.current_mode

    # agent.rs:322-332
    @staticmethod
    def count_in_subgroup(agents, weight):
        agents_size = f32(len(agents))
        groups = {}
        for x in agents:
            groups.setdefault(x.last_mode, []).append(1)
        return {k: f32(len(v)) * weight / agents_size for k, v in groups.items()}

    # agent.rs:93-177
    def calculate_mode_budget(self, mo: MapOrder):
        social_vals = self.count_in_subgroup(self.social_network, self.social_connectivity)
        neighbour_vals = self.count_in_subgroup(self.neighbours, self.neighbourhood_connectivity)
        subculture_vals = {k: v * self.subculture_connectivity
                           for k, v in self.subculture.desirability.items()}
        acc = {}
        for x in (social_vals, neighbour_vals, subculture_vals):
            acc = union_of(acc, x, lambda v1, v2: v1 + v2)
        # max_by: keeps the current maximum only if it compares Greater; ties/NaN -> later one
        best = None
        for k, v in mo.items(acc):
            if best is None or not (best[1] > v):
                best = (k, v)
        self.norm = best[0]

        habit_vals = {k: v * self.consistency for k, v in self.habit.items()}
        acc = {}
        values_to_average = (social_vals, neighbour_vals, subculture_vals, habit_vals)
        for x in values_to_average:
            acc = union_of(acc, x, lambda v1, v2: v1 + v2)
        average = {k: v / f32(len(values_to_average)) for k, v in acc.items()}

        ownership_modifier = {CAR: f32(1.0) if self.owns_car else f32(0.0),
                              CYCLE: f32(1.0) if self.owns_bike else f32(0.0)}
        congestion = self.neighbourhood.congestion_modifier
        acc = {}
        for x in (average, ownership_modifier, congestion):
            acc = union_of(acc, x, lambda v1, v2: v1 * v2)
        ordered = stable_sorted(mo.items(acc), lambda a, b: a[1] > b[1])     # descending
        return {mode: i for i, (mode, _) in enumerate(ordered)}

    # agent.rs:183-240
    def calculate_cost(self, weather, change_in_weather, mo: MapOrder):
        neighbourhood_vals = {k: f32(1.0) - v for k, v in self.neighbourhood.supportiveness.items()}
        commute_cost = journey_cost(self.commute_length)
        average = intersection_of(commute_cost, neighbourhood_vals, lambda v1, v2: (v1 + v2) / f32(2.0))
        if weather == BAD:
            if (not change_in_weather) and self.last_mode in (WALK, CYCLE):
                resolve = f32(-0.1)
            elif not change_in_weather:
                resolve = f32(0.1)
            else:
                resolve = f32(0.0)
            weather_modifier = {PT: f32(1.0), CAR: f32(1.0),
                                CYCLE: f32(1.0) + self.weather_sensitivity + resolve,
                                WALK: f32(1.0) + self.weather_sensitivity + resolve}
            average = intersection_of(dict(average), weather_modifier, lambda v1, v2: v1 * v2)
        ordered = stable_sorted(mo.items(average), lambda a, b: a[1] < b[1])  # ascending
        return {mode: i for i, (mode, _) in enumerate(ordered)}

    # agent.rs:244-264
    def update_habit(self):
        self.last_mode = self.current_mode
        last_mode_map = {self.last_mode: self.average_weight}
        old_habit = {k: v * (f32(1.0) - self.average_weight) for k, v in self.habit.items()}
        acc = {}
        for x in (old_habit, last_mode_map):
            acc = union_of(acc, x, lambda v1, v2: v1 + v2)
        self.habit = acc

    # agent.rs:269-316
    def choose(self, weather, change_in_weather, mo: MapOrder):
        budget = self.calculate_mode_budget(mo)
        cost = self.calculate_cost(weather, change_in_weather, mo)
        total = intersection_of(budget, cost, lambda v1, v2: v1 + v2)
        if not self.owns_car:
            total.pop(CAR, None)
        if not self.owns_bike:
            total.pop(CYCLE, None)
        groups = {}
        for k, v in total.items():
            groups.setdefault(v, []).append(k)
        minimum_values = groups[min(groups)]
        if len(minimum_values) == 1:
            self.current_mode = minimum_values[0]
        else:
            self.current_mode = min(minimum_values, key=lambda m: budget[m])


def build(pop, tab, prm):
    """Objects from the SoA dicts used by the tests (same fields as mvb_population/mvb_tables)."""
    S, H = tab["desirability"].shape[0], tab["supportiveness"].shape[0]
    subs = [Subculture({m: f32(tab["desirability"][s, m]) for m in range(4)}) for s in range(S)]
    nbhds = [Neighbourhood({m: f32(tab["supportiveness"][h, m]) for m in range(4)},
                           {m: int(tab["capacity"][h, m]) for m in range(4)}) for h in range(H)]
    N = len(pop["subculture"])

    def w(name, i):
        arr = pop.get(name)
        return f32(arr[i]) if arr is not None else f32(prm[name])

    agents = []
    for i in range(N):
        a = Agent(
            subculture=subs[int(pop["subculture"][i])], sub_id=int(pop["subculture"][i]),
            neighbourhood=nbhds[int(pop["neighbourhood"][i])], nbhd_id=int(pop["neighbourhood"][i]),
            commute_length=int(pop["commute"][i]),
            weather_sensitivity=f32(pop["weather_sensitivity"][i]),
            consistency=w("consistency", i), social_connectivity=w("social_connectivity", i),
            subculture_connectivity=w("subculture_connectivity", i),
            neighbourhood_connectivity=w("neighbourhood_connectivity", i),
            average_weight=w("average_weight", i),
            habit={m: f32(pop["habit"][i, m]) for m in range(4) if pop["habit"][i, m] != 0},
            current_mode=int(pop["current_mode"][i]), norm=int(pop["norm"][i]),
            owns_car=bool(pop["owns_car"][i]), owns_bike=bool(pop["owns_bike"][i]))
        agents.append(a)
        a.neighbourhood.residents.append(a)
    for i, a in enumerate(agents):
        a.social_network = [agents[j] for j in
                            pop["social_col"][pop["social_rowptr"][i]:pop["social_rowptr"][i + 1]]]
        a.neighbours = [agents[j] for j in
                        pop["neighbour_col"][pop["neighbour_rowptr"][i]:pop["neighbour_rowptr"][i + 1]]]
    return agents, subs, nbhds


def stats_row(agents, S, H):
    """statistics.rs:14-93 in the column order of simulation.rs:254-267 (without Day, Rain)."""
    act = lambda m: m in (WALK, CYCLE)
    row = [0] * (7 + S + H)
    for a in agents:
        A, AN = act(a.current_mode), act(a.norm)
        row[0] += A
        row[1] += AN
        row[2] += A and not AN
        row[3] += (not A) and AN
        if A:
            row[4 + a.commute_length] += 1
            row[7 + a.sub_id] += 1
            row[7 + S + a.nbhd_id] += 1
    return row


def apply_intervention_tables(subs, nbhds, iv):
    """Table part of `intervene`, simulation.rs:307-381.  iv: dict with
    neighbourhood_changes=[(index, {mode: d_supp}, {mode: d_cap})], subculture_changes=[(index, {mode: d})]."""
    for idx, d_supp, d_cap in iv.get("neighbourhood_changes", []):
        n = nbhds[idx]
        n.supportiveness = union_of(n.supportiveness, {k: f32(v) for k, v in d_supp.items()},
                                    lambda v1, v2: v1 + v2)
        signed = {k: int(v) for k, v in n.capacity.items()}
        new = union_of(signed, {k: int(v) for k, v in d_cap.items()}, lambda v1, v2: v1 + v2)
        n.capacity = {k: (0 if v < 0 else (0xFFFFFFFF if v > 0xFFFFFFFF else v)) for k, v in new.items()}
    for idx, d in iv.get("subculture_changes", []):
        s = subs[idx]
        new = union_of(s.desirability, {k: f32(v) for k, v in d.items()}, lambda v1, v2: v1 + v2)
        for k in list(s.desirability):          # iter_mut over EXISTING keys only (:375-379)
            s.desirability[k] = new[k]


def run(pop, tab, prm, weather, n_days, order=CANONICAL_ORDER, intervention=None, dump=None):
    """simulation.rs:95-136.  `intervention` = dict(day=..., neighbourhood_changes=..., subculture_changes=...,
    owns_car=array|None, owns_bike=array|None).  `dump(day, agents, nbhds)` is called after each row."""
    mo = MapOrder(order)
    agents, subs, nbhds = build(pop, tab, prm)
    S, H = len(subs), len(nbhds)
    rows, days = [stats_row(agents, S, H)], [0]
    prev = int(weather[0])
    if dump:
        dump(0, agents, nbhds)
    for day in range(1, n_days):
        if intervention is not None and day == intervention["day"]:
            apply_intervention_tables(subs, nbhds, intervention)
            for name, attr in (("owns_car", "owns_car"), ("owns_bike", "owns_bike")):
                if intervention.get(name) is not None:
                    for a, v in zip(agents, intervention[name]):
                        setattr(a, attr, bool(v))
        if day % 7 < 5:
            new_weather = int(weather[day])
            for a in agents:
                a.update_habit()
            for n in nbhds:
                n.update_congestion_modifier()
            for a in agents:
                a.choose(new_weather, prev != new_weather, mo)
            prev = new_weather
            rows.append(stats_row(agents, S, H))
            days.append(day)
            if dump:
                dump(day, agents, nbhds)
    return days, rows, agents, nbhds


def dense_state(agents):
    N = len(agents)
    habit = np.zeros((N, 4), np.float32)
    for i, a in enumerate(agents):
        for m, v in a.habit.items():
            habit[i, m] = v
    cur = np.array([a.current_mode for a in agents], np.uint8)
    last = np.array([a.last_mode for a in agents], np.uint8)
    norm = np.array([a.norm for a in agents], np.uint8)
    return habit, cur, last, norm
