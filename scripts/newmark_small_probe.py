"""Kernel time of the reference's own Newmark cases (main.py: 10 001 steps) in the persistent banded-Cholesky kernel."""
import sys
sys.path.insert(0, '.')
import numpy as np
import psfem_b200
from psfem_b200 import Polygones as P
E, NU, RHO = 210e9, 0.3, 7850.0
for elt, met in (("quad", "fine"), ("tri", "coarse")):
    L, l, N, G, t = 20, 2, 20, 2, 0.1
    nodes, elements = P.mesh(L, l, N, G, element_type=elt, mesh_type=met, offsetY=l / 2)
    K = P.global_stiffness_matrix(nodes, elements, E, NU, t, elt, met)
    M = P.global_mass_matrix(nodes, elements, RHO, t, elt, met)
    bc = [(int(G / 2), [0, 0]), (int((len(nodes) - 1) - G / 2), [None, 0])]
    shape, scale, times = psfem_b200.ramp_load(len(nodes), L)
    for rep in range(2):
        out = psfem_b200.newmark(K, M, bc, shape, scale, dt=1e-3)
    print(f"{elt}/{met}: {K.shape[0] - 3} DOF, {len(scale)} steps: psfem_newmark_run {out['info']['run_ms']:.1f} ms "
          f"= {1e3 * out['info']['run_ms'] / (len(scale) - 1):.1f} us per step, pcg iterations {out['info']['pcg_iterations']}")
