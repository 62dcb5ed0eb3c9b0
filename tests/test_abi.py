"""The C-ABI library loads and exports every symbol include/motivate_b200.h declares; error behaviour that needs
no GPU (argument checks, 'no CPU path')."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import motivate_b200 as mb
from motivate_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "motivate_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 25
    lib = C.CDLL(_ffi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_ffi.SIGNATURES), set(names) ^ set(_ffi.SIGNATURES)
    assert _ffi.load().mvb_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes on LP64: catches accidental field drift between the header and the ctypes mirror
    assert C.sizeof(_ffi.Population) == 8 + 18 * 8
    assert C.sizeof(_ffi.Tables) == 8 + 3 * 8
    assert C.sizeof(_ffi.Params) == 24
    assert C.sizeof(_ffi.Intervention) == 8 + 3 * 8
    assert C.sizeof(_ffi.SynthSpec) == 7 * 4 + 4 + 8 + 4 + 4 + 8 * 3 * 8


def test_run_rows():
    lib = _ffi.load()
    assert lib.mvb_run_rows(1) == 1
    assert lib.mvb_run_rows(730) == 522           # 2 years: day 0 + 521 weekdays
    assert lib.mvb_run_rows(5 * 365) == 1305      # the 1 304 data rows + day 0 of output/*/output_N.csv


def test_bad_arguments_report_errors_without_a_device():
    lib = _ffi.load()
    h = _ffi.SimP()
    rc = lib.mvb_create(None, None, None, 0, C.byref(h))
    assert rc == -1 and b"null" in lib.mvb_last_error()
    assert lib.mvb_step(None, 0, 0) == -1
    assert lib.mvb_n_replicas(None) == 0
    assert lib.mvb_algorithmic_bytes_per_day(None, 0) == 0
    lib.mvb_destroy(None)


@pytest.mark.skipif(_ffi.load().mvb_device_count() > 0, reason="this checks the no-device behaviour")
def test_no_cpu_fallback():
    pop = mb.synth_population(64, 2, 2, 2, 2, seed=0)
    tab = mb.synth_tables(2, 2, 64)
    with pytest.raises(mb.MotivateError) as e:
        mb.Simulation(pop, tab)
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under motivate_b200/ may import, link or dlopen it."""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|libmotivate_oracle|\bmvo_|motivate_oracle\.h)")
    for root, _, files in os.walk(os.path.join(ROOT, "motivate_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                assert not pat.search(open(os.path.join(root, f)).read()), (root, f)


def test_plain_c_consumer(tmp_path):
    """include/motivate_b200.h compiles as C99 and a C program linked against the library runs (host-only calls)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_consumer")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_consumer.c"), "-o", exe, "-L", libdir, "-lmotivate_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok 5000 ")
