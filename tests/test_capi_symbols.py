"""CPU check of the drop-in boundary: libpsfem.so loads and exports every symbol that
include/psfem.h declares, and the ctypes table binds exactly that set (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "psfem.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psfem_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(psfem):
    lib = ctypes.CDLL(psfem._capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/psfem.h but not exported"


def test_ctypes_table_matches_header(psfem):
    bound = set(psfem._capi.SIGNATURES) | {"psfem_last_error"}
    assert bound == set(declared_symbols())


def test_version_and_npe(psfem):
    lib = psfem._capi.lib
    assert lib.psfem_version() >= 100
    assert [lib.psfem_nodes_per_element(e) for e in range(4)] == [3, 4, 6, 9]
    assert lib.psfem_nodes_per_element(7) == -1


def test_bad_argument_reports_text(psfem):
    import ctypes as C
    sz = C.c_size_t(0)
    rc = psfem._capi.lib.psfem_pattern_workspace(10, 17, 10, C.byref(sz))     # npe = 17: invalid, no GPU touched
    assert rc == psfem._capi.EINVAL
    assert "bad sizes" in psfem._capi.last_error()


def test_symmetric_lattice_storage_size(psfem):
    """psfem_spmv_sl_size is host arithmetic: 19 doubles per node of every owned lattice column (+ the column before the
    first owned one when the rows start inside the x vector), node rows padded to a multiple of 4."""
    import ctypes as C
    lib = psfem._capi.lib
    sz = C.c_size_t(0)
    assert lib.psfem_spmv_sl_size(4096 * 4097, 0, 4097, C.byref(sz)) == 0
    assert sz.value == 4096 * 19 * 4100 * 8                                   # the 33.6 M-DOF plate, clamped: 2.55 GB
    assert lib.psfem_spmv_sl_size(10 * 30, 30, 30, C.byref(sz)) == 0            # a strip behind a left halo column
    assert sz.value == 11 * 19 * 32 * 8
    for bad in ((10, 0, 3), (12, 2, 3), (12, 0, 1), (0, 0, 4)):                # rows / offset not whole columns, m < 2, empty
        assert lib.psfem_spmv_sl_size(*bad, C.byref(sz)) == psfem._capi.EINVAL
    assert "spmv_sl_size" in psfem._capi.last_error()


def test_no_cpu_fallback(psfem):
    """Without a CUDA device every compute entry of the facade must raise, not fall back."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        psfem.Polygones.mesh(1, 1, 2)
    import numpy as np
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        psfem.Polygones.global_stiffness_matrix(np.zeros((3, 2)), {0: [0, 1, 2]}, 1.0, 0.3, 1.0)
