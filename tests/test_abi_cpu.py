"""CPU tests of the drop-in boundary: libalga_gpu.so loads, exports every symbol include/alga_gpu.h declares, and
its compute entry points fail loudly (ALGA_E_CUDA) instead of falling back when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from alga_b200 import _lib, readset
from alga_b200.graph_creator import GraphCreatorPrefSuf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "alga_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(alga_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/alga_gpu.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature in alga_b200/_lib.py"
    assert set(_lib.SYMBOLS) == set(names)


def test_struct_layouts_match_header():
    # sizes the C compiler gives the ABI structs (LP64): guards the ctypes mirrors against drift
    assert C.sizeof(_lib.Reads) == 56
    assert C.sizeof(_lib.PsParams) == 28
    assert C.sizeof(_lib.Csr) == 48
    assert C.sizeof(_lib.Timing) == 48 + 64
    assert C.sizeof(_lib.VerifyParams) == 32  # + lcs_rate_pct, lcs_band
    assert C.sizeof(_lib.InputParams) == 24
    assert C.sizeof(_lib.ReadSetOut) == 88
    assert C.sizeof(_lib.DriverParams) == 40
    assert C.sizeof(_lib.OverlapGraphOut) == 192


def test_version_and_error_strings():
    lib = _lib.load()
    assert b"sm_100a" in lib.alga_gpu_version()
    assert isinstance(lib.alga_gpu_last_error(), bytes)


def test_no_cpu_fallback():
    lib = _lib.load()
    if lib.alga_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    rs = readset.from_code_list([np.zeros(40, np.uint8), np.ones(40, np.uint8)])
    gc = GraphCreatorPrefSuf(rs, 20, 30)
    with pytest.raises(_lib.AlgaGpuError) as e:
        gc.startAlignmentGraphCreation()
    assert e.value.code == -2  # ALGA_E_CUDA
    h = C.c_void_p()
    assert lib.alga_ps_plan_create(C.byref(h), C.byref(_lib.PsParams(20, 30, 0, 500, 0, 0))) == -2


def test_invalid_arguments_are_rejected():
    lib = _lib.load()
    assert lib.alga_gpu_prefsuf_build(None, None, None, None) == -1
    assert lib.alga_ps_plan_run(None, None) == -1
    assert b"null" in lib.alga_gpu_last_error()


def test_header_is_plain_c(tmp_path):
    """include/alga_gpu.h is the boundary: it must compile as C99 and as C++14 on its own (no CUDA, no torch types)."""
    import shutil
    import subprocess

    if not shutil.which("gcc") or not shutil.which("g++"):
        pytest.skip("no host compiler")
    src = tmp_path / "hdr.c"
    src.write_text('#include "alga_gpu.h"\nint main(void) { alga_reads r; alga_csr g; alga_driver_params p; alga_overlap_graph o;'
                   ' (void) r; (void) g; (void) p; (void) o; return 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)], check=True)
    subprocess.run(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)], check=True)
