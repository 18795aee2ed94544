"""psfem_b200 -- B200-native plane-stress FEM hot path (assembly -> static / Newmark-beta / modal).

The directory name carries hyphens, so the package is imported through the tiny loader
`psfem_b200.py` at the repository root:

    import psfem_b200
    from psfem_b200 import Polygones          # drop-in for the reference's module
    nodes, elements = Polygones.mesh(L=10, l=1, N=40)
    K = Polygones.global_stiffness_matrix(nodes, elements, 210e9, 0.3, 0.1)   # scipy CSR
    U, U_red, cd, info = psfem_b200.static_solve(K, F, bc)

Importing the package loads libpsfem.so (hand-written sm_100a kernels, C ABI in include/psfem.h) and
fails loudly if it is missing.  There is no CPU fallback.
"""
from . import _capi  # noqa: F401  (loads the shared library)
from . import Polygones  # noqa: F401
from ._capi import ConvergenceError  # noqa: F401
from .device import (Assembler, DeviceCSR, ElementTable, FreeMap, device_mesh, pcg_jacobi)  # noqa: F401
from .drivers import explicit_euler, modal, newmark, ramp_load, static_solve  # noqa: F401
from .matio import load_matrix, save_matrix  # noqa: F401
from .diagnostics import (check_matrix_properties, critical_time_step, estimate_max_eigenvalue,  # noqa: F401
                          extreme_eigenvalues, rigid_body_check, symmetry_defect)

__version__ = "0.1.0"
