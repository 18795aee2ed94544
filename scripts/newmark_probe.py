"""Newmark-beta step at benchmark scale: per-step time and PCG iterations of psfem_newmark_run on the clamped
Q4 plate (main.py's recurrence, dt = 1e-3, ramp load).  usage: python scripts/newmark_probe.py [cells] [steps]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import psfem_b200
from psfem_b200 import Polygones as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 11
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 2e-7    # ~ h / wave speed: K_eff = K + M/(beta dt^2) is well conditioned
h = 1e-3
nodes, elements = P.mesh(N * h, N * h, N, N, element_type='quad')
K = P.global_stiffness_matrix(nodes, elements, 210e9, 0.3, 0.1, 'quad')
M = P.global_mass_matrix(nodes, elements, 7850.0, 0.1, 'quad')
bc = P.boundary('left', [0, 0], N, N)
shape, scale, times = psfem_b200.ramp_load(len(nodes), N * h, P=1e6, T=10000 * dt, dt=dt)
for energies in (False, True):
    for rep in range(1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = psfem_b200.newmark(K, M, bc, shape, scale[:steps], dt=dt, store_every=steps, energies=energies, rtol=1e-8, maxit=300)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    its = out['info']['pcg_iterations']
    print(f"dt={dt} N={N} steps={steps - 1} energies={energies}: driver wall {t1 - t0:.3f} s, PCG iterations {its} "
          f"({its / (steps - 1):.1f}/step), max relres {out['info']['max_relres']:.2e}, kernel time {out['info'].get('run_ms', float('nan')):.2f} ms "
          f"= {out['info'].get('run_ms', float('nan')) / (steps - 1):.3f} ms/step")
