"""Assembly of the quadratic element kinds (T6 / Q9, mesh_type='fine') at scale: time, rate and fraction of the HBM roofline
(coordinates + CSR values, written once).  usage: python scripts/fine_probe.py [cells]"""
import sys, time
sys.path.insert(0, '.')
import torch
import psfem_b200
from psfem_b200 import _capi
from psfem_b200.device import Assembler, device_mesh
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ev = lambda: torch.cuda.Event(enable_timing=True)
for et, mt in (("quad", "fine"), ("tri", "fine")):
    nodes, conn = device_mesh(cells * 1e-3, cells * 1e-3, cells, cells, 0, 0, et, mt)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    asm = Assembler(conn, nodes.shape[0], _capi.ETYPE[(et, mt)])
    torch.cuda.synchronize(); t_sym = time.perf_counter() - t0
    K = torch.empty(asm.nnz, dtype=torch.float64, device="cuda")
    ts = []
    for what, kw in ((1, dict(E=210e9, nu=0.3, t=0.1)), (2, dict(rho=7850.0, t=0.1))):
        for rep in range(4):
            a, b = ev(), ev()
            a.record()
            if what == 1: asm.assemble(nodes, what=1, K_out=K, **kw)
            else: asm.assemble(nodes, what=2, M_out=K, **kw)
            b.record(); torch.cuda.synchronize()
            ts.append((what, a.elapsed_time(b)))
    for what in (1, 2):
        ms = min(t for w, t in ts[1:] if w == what)
        nbytes = 8 * asm.nnz + 16 * nodes.shape[0]
        print(f"{et}/{mt} {cells}x{cells}: {asm.nel} elements, {nodes.shape[0]} nodes, nnz {asm.nnz}; symbolic {1e3 * t_sym:.1f} ms; "
              f"{'K' if what == 1 else 'M'} {ms:.3f} ms = {asm.nel / ms / 1e6:.2f} G el/s, {nbytes / asm.nel:.0f} B/element algorithmic -> "
              f"{nbytes / (ms * 1e-3) / 1e9:.0f} GB/s = {nbytes / (ms * 1e-3) / 1e9 / 6546.6:.1%} of the HBM roofline; path: "
              f"{'lattice' if asm.use_lattice else 'tiles' if asm.use_tiles else 'gauss + segmented reduction'}")
    del asm, K, nodes, conn
    torch.cuda.empty_cache()
