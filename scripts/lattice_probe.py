import sys, time, torch, os
sys.path.insert(0, '.')
import psfem_b200
from psfem_b200 import _capi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
jit = len(sys.argv) > 2 and sys.argv[2] == 'jitter'
et = sys.argv[3] if len(sys.argv) > 3 else 'quad'
nodes, conn = psfem_b200.device_mesh(N * 1e-3, N * 1e-3, N, N, element_type=et)
if jit:
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    nodes = nodes + (torch.rand(nodes.shape, generator=g, device='cuda', dtype=torch.float64) - 0.5) * 0.4e-3
asm = psfem_b200.Assembler(conn, (N + 1) ** 2, _capi.Q4 if et == 'quad' else _capi.T3)
assert asm.use_lattice
K, _ = asm.assemble(nodes, E=210e9, nu=.3, t=.1, what=1)
ev = lambda: torch.cuda.Event(enable_timing=True)
for what in (1, 2):
    out = torch.empty_like(K)
    ts = []
    for rep in range(5):
        a, b = ev(), ev()
        a.record()
        asm.assemble(nodes, E=210e9, nu=.3, rho=7850., t=.1, what=what, K_out=out, M_out=out)
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(et, 'N', N, 'jitter', jit, 'what', what, 'SEG', os.environ.get('PSFEM_LATTICE_SEG'), 'OCC', os.environ.get('PSFEM_LATTICE_OCC'),
          'ms min %.3f med %.3f' % (min(ts), sorted(ts)[2]), 'GB/s(algorithmic) %.0f' % ((320 if et == 'quad' else 2 * 132) * N * N / min(ts) / 1e6))
