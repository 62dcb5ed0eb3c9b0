#!/usr/bin/env python
"""Convert a dump written by rust/patch/dump_state.rs (an UNMODIFIED CPU run of the reference with three logging
calls added) into tests/golden/ref_<name>.npz, the fixture tests/test_reference_fixtures.py consumes.

    python tools/ref_dump_to_npz.py DUMP_DIR/<id> tests/golden/ref_<name>.npz [--max-rows K]

Layout of the dump: see the header of rust/patch/dump_state.rs (little-endian, no padding)."""
import os
import sys

import numpy as np


def read_dump(d, max_rows=None):
    b = open(os.path.join(d, "inputs.bin"), "rb").read()
    assert b[:4] == b"MVBD", "not a dump_state.rs file"
    ver, N, S, H, n_days = (int(x) for x in np.frombuffer(b, "<u4", 5, 4))
    assert ver == 1
    at = [24]

    def take(dt, *shape):
        n = int(np.prod(shape))
        a = np.frombuffer(b, dt, n, at[0]).reshape(shape).copy()
        at[0] += a.nbytes
        return a

    out = dict(n_days=np.uint32(n_days))
    out["subculture"] = take("u1", N); out["neighbourhood"] = take("<u2", N); out["commute"] = take("u1", N)
    out["owns_car"] = take("u1", N); out["owns_bike"] = take("u1", N); out["weather_sensitivity"] = take("<f4", N)
    for k in ("consistency", "social_connectivity", "subculture_connectivity", "neighbourhood_connectivity", "average_weight"):
        out[k] = take("<f4", N)
    out["habit"] = take("<f4", N, 4); out["current_mode"] = take("u1", N); out["norm"] = take("u1", N)
    for g in ("social", "neighbour"):
        rp = take("<u8", N + 1)
        out[g + "_rowptr"] = rp
        out[g + "_col"] = take("<u4", int(rp[-1]))
    out["supportiveness"] = take("<f4", H, 4); out["capacity"] = take("<u4", H, 4); out["desirability"] = take("<f4", S, 4)
    out["weather"] = take("u1", n_days)
    assert at[0] == len(b), "trailing bytes in inputs.bin"
    ip = os.path.join(d, "intervention.bin")
    if os.path.exists(ip):
        v = open(ip, "rb").read()
        out["iv_day"] = np.frombuffer(v, "<u4", 1, 0)[0]
        o = 4
        out["iv_supportiveness"] = np.frombuffer(v, "<f4", H * 4, o).reshape(H, 4).copy(); o += 16 * H
        out["iv_capacity"] = np.frombuffer(v, "<u4", H * 4, o).reshape(H, 4).copy(); o += 16 * H
        out["iv_desirability"] = np.frombuffer(v, "<f4", S * 4, o).reshape(S, 4).copy(); o += 16 * S
        out["iv_owns_car"] = np.frombuffer(v, "u1", N, o).copy(); o += N
        out["iv_owns_bike"] = np.frombuffer(v, "u1", N, o).copy(); o += N
        assert o == len(v)
    rec = 4 + N + N + 16 * N
    raw = open(os.path.join(d, "days.bin"), "rb").read()
    rows = len(raw) // rec
    assert rows * rec == len(raw), "days.bin is not a whole number of records"
    if max_rows:
        rows = min(rows, max_rows)
    days = np.zeros(rows, np.uint32); modes = np.zeros((rows, N), np.uint8); norms = np.zeros((rows, N), np.uint8)
    habits = np.zeros((rows, N, 4), np.float32)
    for k in range(rows):
        o = k * rec
        days[k] = np.frombuffer(raw, "<u4", 1, o)[0]
        modes[k] = np.frombuffer(raw, "u1", N, o + 4)
        norms[k] = np.frombuffer(raw, "u1", N, o + 4 + N)
        habits[k] = np.frombuffer(raw, "<f4", 4 * N, o + 4 + 2 * N).reshape(N, 4)
    out.update(days=days, day_current_mode=modes, day_norm=norms, day_habit=habits)
    return out


if __name__ == "__main__":
    argv = sys.argv[1:]
    mx = None
    if "--max-rows" in argv:
        i = argv.index("--max-rows")
        mx = int(argv[i + 1])
        del argv[i:i + 2]
    data = read_dump(argv[0], mx)
    np.savez_compressed(argv[1], **data)
    print(f"wrote {argv[1]}: {len(data['subculture'])} agents, {len(data['days'])} rows")
