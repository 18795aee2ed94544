import sys, time, torch, os
sys.path.insert(0, '.')
import psfem_b200
from psfem_b200 import _capi
N = 4096
nodes, conn = psfem_b200.device_mesh(N * 1e-3, N * 1e-3, N, N, element_type='quad')
for R in (64, 128):
    asm = psfem_b200.Assembler(conn, (N + 1) ** 2, _capi.Q4)
    asm.tile_nodes_R = R
    asm.use_lattice = False            # force the general tile-fused path on this lattice mesh
    K, _ = asm.assemble(nodes, E=210e9, nu=.3, t=.1, what=1)
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        asm.assemble(nodes, E=210e9, nu=.3, t=.1, what=1, K_out=K)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print('MINB', os.environ.get('PSFEM_TILE_MINB'), 'R', R, 'ms', min(ts))
    del asm
