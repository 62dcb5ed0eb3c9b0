"""Load a tests/golden/*.npz fixture back into population / tables / params objects."""
import os

import numpy as np

import motivate_b200 as mb

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["random_uniform_weights", "random_per_agent_weights", "synthetic_with_intervention"]


def load(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    pop = mb.PopulationData({k[4:]: z[k] for k in z.files if k.startswith("pop_")})
    tab = mb.TablesData(z["tab_supportiveness"], z["tab_capacity"], z["tab_desirability"])
    p = z["params"]
    prm = mb.make_params(consistency=float(p[0]), social_connectivity=float(p[1]), subculture_connectivity=float(p[2]),
                         neighbourhood_connectivity=float(p[3]), average_weight=float(p[4]))
    iv = None
    if "iv_day" in z.files:
        iv = mb.InterventionData(day=int(z["iv_day"]),
                                 tables=mb.TablesData(z["iv_supportiveness"], z["iv_capacity"], z["iv_desirability"]),
                                 owns_car=z["iv_owns_car"], owns_bike=z["iv_owns_bike"])
    return z, pop, tab, prm, iv


def replay(z, sim_step, sim_state, sim_row, apply_iv, iv):
    """Drive a simulation day by day along the fixture's calendar and compare state + row after every weekday."""
    weather, n_days = z["weather"], int(z["n_days"])
    assert np.array_equal(sim_row(), z["rows"][0])
    prev, r = weather[0], 1
    for day in range(1, n_days):
        if iv is not None and day == iv.day:
            apply_iv(iv)
        if day % 7 >= 5:
            continue
        sim_step(weather[day], prev != weather[day])
        prev = weather[day]
        st = sim_state()
        assert int(z["days"][r]) == day
        assert np.array_equal(st["current_mode"], z["mode"][r]), f"mode, day {day}"
        assert np.array_equal(st["norm"], z["norm"][r]), f"norm, day {day}"
        assert np.array_equal(st["habit"].view(np.uint32), z["habit"][r].view(np.uint32)), f"habit bits, day {day}"
        assert np.array_equal(sim_row(), z["rows"][r]), f"row, day {day}"
        r += 1
    assert r == len(z["days"])
