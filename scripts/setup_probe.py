"""Device-timed set-up kernels at the 4096 x 4096 plate: the closed-form lattice pattern (psfem_lattice_pattern) and the
boundary-condition reduction of the pattern (psfem_reduce_fill).  usage: python scripts/setup_probe.py [cells]"""
import ctypes as C
import sys
sys.path.insert(0, '.')
import torch
import psfem_b200  # noqa: F401
from psfem_b200 import _capi
from psfem_b200._capi import call, ptr, stream_ptr
from psfem_b200.device import FreeMap

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = m = cells + 1
i32 = dict(dtype=torch.int32, device="cuda")
npairs = C.c_int64(0)
call("psfem_lattice_pattern", n, m, _capi.Q4, None, None, None, C.byref(npairs), stream_ptr())
nnz = 4 * npairs.value
row_ptr, col_idx, node_ptr = torch.empty(2 * n * m + 1, **i32), torch.empty(nnz, **i32), torch.empty(n * m + 1, **i32)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        a, b = ev(), ev()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts[1:])


t = timed(lambda: call("psfem_lattice_pattern", n, m, _capi.Q4, ptr(row_ptr), ptr(col_idx), ptr(node_ptr), C.byref(npairs), stream_ptr()))
bytes_pat = 4 * (nnz + 3 * n * m)
print(f"lattice_pattern {cells}^2: {t:.3f} ms, {bytes_pat / t / 1e6:.0f} GB/s written ({bytes_pat / 1e9:.2f} GB)")
fm = FreeMap(2 * n * m, list(range(2 * m)))
row_ptr_red = torch.empty(fm.n_free + 1, **i32)
nz = C.c_int64(0)
call("psfem_reduce_count", fm.n_total, ptr(row_ptr), ptr(col_idx), ptr(fm.new_id), fm.n_free, ptr(row_ptr_red), C.byref(nz),
     ptr(fm._ws), fm._ws_bytes, stream_ptr())
col_red, keep = torch.empty(nz.value, **i32), torch.empty(nz.value, **i32)
t = timed(lambda: call("psfem_reduce_fill", fm.n_total, ptr(row_ptr), ptr(col_idx), ptr(fm.new_id), ptr(row_ptr_red), ptr(col_red), ptr(keep),
                       stream_ptr()))
bytes_red = 4 * (nnz + 2 * nz.value + 2 * fm.n_total + fm.n_free)
print(f"reduce_fill {cells}^2: {t:.3f} ms, {bytes_red / t / 1e6:.0f} GB/s ({bytes_red / 1e9:.2f} GB: col_idx read, new_id, col_idx_red + keep written)")
# spot check against a direct evaluation on the device
k = torch.arange(0, nz.value, max(nz.value // 1000003, 1), device="cuda")
src = keep[k].long()
assert torch.equal(fm.new_id[col_idx[src].long()], col_red[k]), "reduce_fill: keep / col_idx_red mismatch"
print("ok")
