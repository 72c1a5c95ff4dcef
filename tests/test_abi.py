"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/p3m_b200.h declares, and its host-only entry points behave like the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import conftest
import refapi
from particlesimulation_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "p3m_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p3m_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    names = declared_symbols()
    assert names == sorted(capi.SYMBOLS), set(names) ^ set(capi.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), f"{n} missing from libp3m_b200.so"


def test_params_struct_layout_and_defaults():
    p = capi.default_params()
    assert C.sizeof(p) == 120  # 30 four-byte fields, no padding
    assert p.sr_particle_diameter == 0.0
    assert (p.assignment, p.fd_scheme, p.greens_function) == (capi.TSC, capi.TWO_POINT, capi.S1_OPTIMAL)
    assert p.use_sr_table == 1 and p.DT == 1.0 and p.device == -1 and p.unit_roundtrip == 1


def test_chaining_neighbors_match_reference_order():
    """ChainingMesh::getNeighborsAndSelf: 9 cells of row y-1, 3 of (y, z-1), (x-1,y,z), self; -1 outside."""
    dims = np.array([5, 4, 3], np.int32)
    o = refapi.Oracle("f32")
    for cell in range(int(dims.prod())):
        got = capi.chaining_neighbors(dims, cell)
        assert np.array_equal(got, o.chaining_neighbors(dims, cell))
        assert got[13] == cell
    # corner cell 0: everything but self is outside
    assert np.array_equal(capi.chaining_neighbors(dims, 0), [-1] * 13 + [0])
    if refapi.have_ref():
        p = refapi.make_params(1, (16, 16, 16), (5.0, 4.0, 3.0), H=1.0, cutoff=1.0)
        ref = refapi.Ref()
        for cell in (0, 7, 26, 33, 59):
            assert np.array_equal(capi.chaining_neighbors(dims, cell), ref.chaining_neighbors(p, cell))


def test_invalid_arguments_fail_loudly():
    with pytest.raises(capi.P3MError):
        capi.chaining_neighbors(np.array([2, 2, 2], np.int32), 99)
    lib = capi.lib()
    assert lib.p3m_create(None, None) == -1
    assert b"null" in lib.p3m_last_error()


@pytest.mark.skipif(conftest._have_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device():
    p = capi.default_params()
    p.nx = p.ny = p.nz = 16
    p.box[:] = [8.0, 8.0, 8.0]
    with pytest.raises(capi.P3MError) as e:
        capi.Context(p)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "particlesimulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) and "build" not in dirpath:
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "p3m_oracle" not in text and "libp3m_ref" not in text, os.path.join(dirpath, f)
