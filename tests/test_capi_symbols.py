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
