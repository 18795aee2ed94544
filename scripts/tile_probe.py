import sys, time, torch
sys.path.insert(0,'.')
import psfem_b200
from psfem_b200 import _capi
N=4096
nodes, conn = psfem_b200.device_mesh(4.096,4.096,N,N,element_type='quad')
for R in (64,128,256):
    asm = psfem_b200.Assembler(conn, (N+1)**2, _capi.Q4)
    asm.tile_nodes_R = R
    K,_ = asm.assemble(nodes, E=210e9, nu=.3, t=.1, what=1)
    tp = asm._tiles
    print('R',R,'info',list(tp['info'])[:5], 'tile_elems/nel', tp['info'][0]/asm.nel)
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.perf_counter()
        asm.assemble(nodes, E=210e9, nu=.3, t=.1, what=1, K_out=K)
        torch.cuda.synchronize(); print('  ms', (time.perf_counter()-t0)*1e3)
    del asm
