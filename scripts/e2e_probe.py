"""Where the end-to-end static_solve(host CSR, host F, bc) call spends its time at benchmark scale."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, scipy.sparse as sparse, torch
import psfem_b200
from psfem_b200 import Polygones as P
from psfem_b200.device import DeviceCSR, FreeMap, constrained_dofs_of, pcg_jacobi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
h = 1e-3
nodes, elements = P.mesh(N * h, N * h, N, N, element_type='quad')
F = P.edgeForces(np.zeros(2 * len(nodes)), [0, -1e6], 'right', N * h, N, N)
bc = P.boundary('left', [0, 0], N, N)
def T(label, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print(f"  {label:34s} {1e3 * (t1 - t0):9.1f} ms", flush=True); return t1
for rep in range(2):
    print("rep", rep)
    t = time.perf_counter()
    K = P.global_stiffness_matrix(nodes, elements, 210e9, 0.3, 0.1, 'quad'); t = T("global_stiffness_matrix (H2D nodes, assemble, D2H CSR)", t)
    Kh = sparse.csr_matrix((K.data, K.indices, K.indptr), shape=K.shape, copy=False); del K; t = T("rewrap host CSR", t)
    print("  pinned:", torch.from_numpy(Kh.data).is_pinned(), torch.from_numpy(Kh.indices).is_pinned())
    for name, arr in (("data", Kh.data), ("indices", Kh.indices), ("indptr", Kh.indptr)):
        tt = time.perf_counter(); d_ = torch.from_numpy(arr).cuda(); torch.cuda.synchronize()
        print(f"    H2D {name}: {arr.nbytes / 1e9:.2f} GB at {arr.nbytes / 1e9 / (time.perf_counter() - tt):.1f} GB/s"); del d_
    hp = torch.empty(Kh.data.shape[0], dtype=torch.float64, pin_memory=True)
    tt = time.perf_counter(); d_ = hp.cuda(); torch.cuda.synchronize()
    print(f"    H2D fresh pinned tensor: {hp.numel() * 8 / 1e9 / (time.perf_counter() - tt):.1f} GB/s"); del d_, hp
    t = time.perf_counter()
    Kd = DeviceCSR.from_scipy(Kh); t = T("DeviceCSR.from_scipy (H2D + device check)", t)
    fm = FreeMap(Kd.n_rows, constrained_dofs_of(bc)); t = T("FreeMap", t)
    K_red = fm.reduce_csr(Kd); t = T("reduce_csr (+ SpMV plans)", t)
    F_red = fm.reduce_vec(torch.from_numpy(F).cuda()); t = T("reduce_vec", t)
    try:
        U_red, info = pcg_jacobi(K_red, F_red, rtol=1e-8, maxit=200)
    except psfem_b200.ConvergenceError:
        pass
    t = T("pcg 200 iterations", t)
    U = fm.expand(K_red.matvec(F_red)); Uh = U.cpu().numpy(); t = T("expand + D2H U", t)
    del Kd, K_red, fm, Kh
