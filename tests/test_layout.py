"""CPU check of the device layout builder (motivate_b200/csrc/layout.cpp): the invariants the kernels rely on —
permutation + padding, hub boundary, neighbourhood-uniform slices, staged ranges within the shared-memory capacity,
link codes, slice lengths and offsets, hub units that cover every hub list exactly, balanced hub slices, partition
bounds.  The checker is tests/c/layout_invariants.cpp, linked against the built library."""
import os
import shutil
import subprocess

import pytest

from motivate_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("layout") / "layout_invariants")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "motivate_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "layout_invariants.cpp"),
                    "-o", exe, "-L", libdir, "-lmotivate_b200", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


# (agents, subcultures, neighbourhoods, shared-memory capacity in agents, seed)
@pytest.mark.parametrize("n,s,h,cap,seed", [
    (3000, 3, 4, 914880, 1),          # everything fits
    (40000, 4, 10, 8192, 2),          # most agents cold
    (40000, 2, 300, 914880, 3),       # many small neighbourhoods (regions padded to 64)
    (200000, 4, 10, 64000, 4),        # hubs + partial hot set
    (5000, 1, 1, 64, 5),              # smallest capacity the builder accepts: more hubs than fit
])
@pytest.mark.parametrize("mode", ["global", "local"])
def test_layout_invariants(checker, n, s, h, cap, seed, mode):
    out = subprocess.run([checker, str(n), str(s), str(h), str(cap), str(seed)], capture_output=True, text=True,
                         env=dict(os.environ, MVB_LAYOUT=mode))
    assert out.returncode == 0 and out.stdout.startswith("ok "), out.stdout + out.stderr
    fields = out.stdout.split()
    assert fields[fields.index("mode") + 1] == (mode if n > cap else "global")       # a population that fits is staged whole
    padding = float(fields[fields.index("sell_padding") + 1])
    if n // h >= 2000:                                # (tiny neighbourhoods are mostly the ragged last slice)
        assert padding < 1.25, out.stdout             # degree / hot-length sorted slices keep the padding small
    if (n, h, cap) == (40000, 10, 8192):
        # staging the block's own neighbourhood makes far more links shared-memory gathers than one global set does
        share = float(fields[fields.index("hot_link_share") + 1])
        assert share > 0.6 if mode == "local" else share < 0.5, out.stdout
    balance = float(fields[fields.index("hub_slice_balance") + 1])
    assert balance < 2.0, out.stdout                  # hubs are dealt round the hub slices
    assert fields[fields.index("push") + 1] == "0"    # small populations keep the pull path


@pytest.mark.parametrize("mode", ["global", "local"])
def test_layout_invariants_push_mode(checker, mode):
    """PUSH mode (layout.h), forced on a small population: the SELL blocks hold hot rows only (offsets and slice records
    without the cold rows), sell_K still carries the cold lengths for the work model, everything else is unchanged."""
    out = subprocess.run([checker, "40000", "4", "10", "8192", "2"], capture_output=True, text=True,
                         env=dict(os.environ, MVB_LAYOUT=mode, MVB_PUSH="1"))
    assert out.returncode == 0 and out.stdout.startswith("ok "), out.stdout + out.stderr
    fields = out.stdout.split()
    assert fields[fields.index("push") + 1] == "1"
