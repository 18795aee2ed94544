"""Classic three-kernel PCG iteration against the fused one-kernel iteration (psfem_cg_sl.cuh) at benchmark scale.
usage: python scripts/cg_probe.py [cells] [cols]"""
import os, sys, time
sys.path.insert(0, '.')
import torch
import psfem_b200
from psfem_b200.dist import StripProblem
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cols = int(sys.argv[2]) if len(sys.argv) > 2 else cells
ev = lambda: torch.cuda.Event(enable_timing=True)
for mode in ("classic", "fused"):
    os.environ["PSFEM_PCG"] = mode
    prob = StripProblem(L=cols * 1e-3, l=cells * 1e-3, N=cols, M=cells, rank=0, world=1).setup()
    ts = []
    for rep in range(5):
        a, b = ev(), ev()
        a.record(); prob.cg_iterations(50); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 50)
    r, bn = prob.residual()
    n = prob.n_owned_free
    nb = prob.iteration_bytes((19 * 8 + 32) * (n // 2))
    print(f"{mode:8s} {cols}x{cells}: {min(ts[1:]):.4f} ms/iteration (all: {['%.4f' % t for t in ts]}), {n / min(ts[1:]) / 1e6:.2f} G DOF*it/s, "
          f"{nb / n:.0f} B/DOF -> {nb / (min(ts[1:]) * 1e-3) / 1e9:.0f} GB/s; relres after 250 it {r / bn:.4e}; kernels/iter {prob.kernels_per_iteration}")
    if mode == "fused":
        t0 = time.perf_counter(); info = prob.full_solve(1e-8); print("full solve:", info)
    prob.release(); torch.cuda.empty_cache()
