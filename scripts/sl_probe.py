"""Diagnostics of the symmetric-lattice SpMV storage (psfem_spmv_sl_*): candidates, build verdicts, symmetry of K."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import psfem_b200  # noqa: E402
from psfem_b200 import Polygones as P  # noqa: E402
from psfem_b200._capi import call, ptr, stream_ptr  # noqa: E402

E, NU, T = 210e9, 0.3, 0.1
for et in ("quad", "tri"):
    nodes, elements = P.mesh(3.0, 2.0, 200, 150, element_type=et)
    K = P.global_stiffness_matrix(nodes, elements, E, NU, T, et)
    D = (K - K.T).tocsr()
    print(et, "max|K-K^T|/max|K| =", abs(D).max() / abs(K).max(), "entries that differ:", D.count_nonzero(), "of", K.nnz)
    dev = P.device_copy(K)
    s = dev.struct()
    cand, cnt = (C.c_int64 * 8)(), C.c_int64(0)
    call("psfem_spmv_sl_candidates", C.byref(s), 0, K.shape[0] // 2, cand, C.byref(cnt), stream_ptr())
    print(" candidates:", list(cand[: cnt.value]), "nb plan:", dev.plan()[3] is not None)
    for tol in (0.0, 1e-14, 1e-10):
        dev._sl = None
        print("  tol", tol, "->", dev.sym_lattice(tol=tol))
