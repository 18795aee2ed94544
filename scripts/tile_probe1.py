import sys, time, torch
sys.path.insert(0, '.')
import psfem_b200
from psfem_b200 import _capi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R = int(sys.argv[2]) if len(sys.argv) > 2 else 128
nodes, conn = psfem_b200.device_mesh(N * 1e-3, N * 1e-3, N, N, element_type='quad')
asm = psfem_b200.Assembler(conn, (N + 1) ** 2, _capi.Q4)
asm.tile_nodes_R = R
K, _ = asm.assemble(nodes, E=210e9, nu=.3, t=.1, what=1)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    asm.assemble(nodes, E=210e9, nu=.3, t=.1, what=1, K_out=K)
    torch.cuda.synchronize(); print('ms', (time.perf_counter() - t0) * 1e3)
