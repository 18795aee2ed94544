"""ctypes binding of libpsfem.so (include/psfem.h) -- the only bridge to the CUDA kernels.

There is no CPU fallback: if the shared library is missing the import fails loudly, and every
compute call needs a CUDA device.  Torch is used for device memory and the current stream only.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsfem.so")

OK, EINVAL, ECUDA, EJACOBIAN, ENOCONV = 0, 1, 2, 3, 4
T3, Q4, T6, Q9 = 0, 1, 2, 3

ETYPE = {("tri", "coarse"): T3, ("quad", "coarse"): Q4, ("tri", "fine"): T6, ("quad", "fine"): Q9}
NPE = {T3: 3, Q4: 4, T6: 6, Q9: 9}


class ConvergenceError(RuntimeError):
    """Jacobi-PCG reached maxit before the requested relative residual."""


class psfem_csr(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("nnz", C.c_int64), ("row_ptr", C.c_void_p), ("col_idx", C.c_void_p),
                ("vals", C.c_void_p), ("rowblk", C.c_void_p), ("n_blocks", C.c_int64), ("max_row_nnz", C.c_int64),
                ("nb_node_ptr", C.c_void_p), ("nb_pcol", C.c_void_p), ("nb_pcol16", C.c_void_p), ("nb_nodeblk", C.c_void_p),
                ("nb_n_blocks", C.c_int64), ("nb_max_pairs", C.c_int64),
                ("sl_vals", C.c_void_p), ("sl_src_vals", C.c_void_p), ("sl_m", C.c_int64), ("sl_roff", C.c_int64),
                ("sl_x_nodes", C.c_int64)]


MAX_RANKS = 8


class psfem_dist_ctx(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("has_left", C.c_int32), ("has_right", C.c_int32),
                ("board", C.c_void_p * MAX_RANKS), ("p_ext", C.c_void_p), ("left_p_ext", C.c_void_p), ("right_p_ext", C.c_void_p),
                ("own_lo", C.c_int64), ("n_own", C.c_int64), ("send_left_n", C.c_int64), ("left_dst_off", C.c_int64),
                ("send_right_n", C.c_int64), ("right_dst_off", C.c_int64),
                ("x", C.c_void_p), ("r", C.c_void_p), ("q", C.c_void_p), ("dinv", C.c_void_p), ("out3", C.c_void_p),
                ("partials", C.c_void_p), ("counter", C.c_void_p), ("masked", C.c_int32)]


class psfem_dist_cg_ctx(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("has_left", C.c_int32), ("has_right", C.c_int32),
                ("board", C.c_void_p * MAX_RANKS), ("ll_mine", C.c_void_p), ("ll_left", C.c_void_p), ("ll_right", C.c_void_p),
                ("m", C.c_int64), ("x", C.c_void_p), ("counter", C.c_void_p), ("masked", C.c_int32)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()' "
        "or make -C plane-stress-fem-with-python_b200/csrc). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_i64, _i32, _dbl, _ptr, _sz = C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_size_t
_pi64 = C.POINTER(C.c_int64)
_pdbl = C.POINTER(C.c_double)
_psz = C.POINTER(C.c_size_t)
_pcsr = C.POINTER(psfem_csr)
_pctx = C.POINTER(psfem_dist_ctx)
_u64 = C.c_uint64
_pcgctx = C.POINTER(psfem_dist_cg_ctx)

# name -> argtypes; must list every symbol include/psfem.h declares (tests/test_capi_symbols.py checks)
SIGNATURES = {
    "psfem_version": [],
    "psfem_nodes_per_element": [_i32],
    "psfem_set_device": [_i32],
    "psfem_mesh_rect": [_dbl, _dbl, _i64, _i64, _dbl, _dbl, _i32, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr],
    "psfem_pattern_workspace": [_i64, _i32, _i64, _psz],
    "psfem_pattern_count": [_ptr, _i64, _i32, _i64, _ptr, _ptr, _sz, _pi64, _ptr],
    "psfem_pattern_fill": [_i64, _i32, _i64, _i64, _ptr, _sz, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "psfem_assemble_workspace": [_i64, _i32, _i32, _psz],
    "psfem_assemble": [_ptr, _ptr, _i64, _i32, _i64, _dbl, _dbl, _dbl, _dbl, _i32, _i64, _ptr, _ptr, _ptr, _ptr,
                       _ptr, _ptr, _ptr, _sz, _pi64, _ptr],
    "psfem_tileplan_workspace": [_i64, _i32, _i64, _i32, _psz],
    "psfem_tileplan_build": [_ptr, _ptr, _i64, _i32, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _pi64,
                             _ptr, _sz, _ptr],
    "psfem_assemble_tiled": [_ptr, _i64, _i32, _i64, _dbl, _dbl, _dbl, _dbl, _i32, _i32, _pi64, _ptr, _ptr, _ptr, _ptr, _ptr,
                             _ptr, _ptr, _ptr, _sz, _pi64, _ptr],
    "psfem_lattice_detect": [_ptr, _i64, _i32, _i64, _pi64, _ptr, _sz, _ptr],
    "psfem_lattice_pattern": [_i64, _i64, _i32, _ptr, _ptr, _ptr, _pi64, _ptr],
    "psfem_assemble_lattice": [_ptr, _i64, _i64, _i32, _dbl, _dbl, _dbl, _dbl, _i32, _ptr, _ptr, _ptr, _sz, _pi64, _ptr],
    "psfem_assemble_lattice_sl": [_ptr, _i64, _i64, _i32, _dbl, _dbl, _dbl, _dbl, _i32, _i64, _i64, _ptr, _sz, _ptr, _sz, _pi64, _ptr],
    "psfem_element_matrices": [_ptr, _ptr, _i64, _i32, _dbl, _dbl, _dbl, _dbl, _i32, _ptr, _ptr, _ptr, _sz, _pi64, _ptr],
    "psfem_b_matrix_at_point": [_ptr, _ptr, _i32, _dbl, _dbl, _ptr, _pdbl, _ptr],
    "psfem_scan_workspace": [_i64, _psz],
    "psfem_free_map": [_i64, _ptr, _i64, _ptr, _ptr, _pi64, _ptr, _sz, _ptr],
    "psfem_reduce_count": [_i64, _ptr, _ptr, _ptr, _i64, _ptr, _pi64, _ptr, _sz, _ptr],
    "psfem_reduce_fill": [_i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "psfem_csr_check": [_i64, _i64, _ptr, _ptr, _i64, _pi64, _ptr, _sz, _ptr],
    "psfem_gather": [_ptr, _ptr, _i64, _ptr, _ptr],
    "psfem_scatter_free": [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr],
    "psfem_spmv_plan_size": [_i64, _i64, _pi64],
    "psfem_spmv_plan": [_i64, _ptr, _i64, _ptr, _i64, _pi64, _ptr],
    "psfem_spmv_nb_plan_size": [_i64, _i64, _pi64],
    "psfem_spmv_nb_plan": [_i64, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _pi64, _ptr],
    "psfem_spmv_nb_compress": [_i64, _ptr, _ptr, _i64, _ptr, _ptr],
    "psfem_spmv_sl_candidates": [_pcsr, _i64, _i64, _pi64, _pi64, _ptr],
    "psfem_spmv_sl_size": [_i64, _i64, _i64, _psz],
    "psfem_spmv_sl_build": [_pcsr, _i64, _i64, _i64, _dbl, _ptr, _sz, _pi64, _ptr],
    "psfem_spmv": [_pcsr, _ptr, _dbl, _dbl, _ptr, _ptr, _ptr],
    "psfem_spmv2": [_pcsr, _ptr, _ptr, _ptr, _dbl, _dbl, _dbl, _ptr, _ptr, _ptr],
    "psfem_extract_diag": [_pcsr, _i64, _i32, _ptr, _ptr],
    "psfem_pcg_workspace": [_i64, _i64, _psz],
    "psfem_pcg_jacobi": [_pcsr, _ptr, _ptr, _dbl, _i64, _i64, _pdbl, _ptr, _sz, _ptr],
    "psfem_pcg_jacobi_masked": [_pcsr, _ptr, _ptr, _ptr, _dbl, _i64, _i64, _pdbl, _ptr, _sz, _ptr],
    "psfem_band_width": [_pcsr, _pi64, _ptr, _sz, _ptr],
    "psfem_direct_limits": [_pi64, _pi64],
    "psfem_direct_workspace": [_i64, _i64, _psz],
    "psfem_direct_solve": [_pcsr, _ptr, _ptr, _i64, _i32, _pdbl, _ptr, _sz, _ptr],
    "psfem_direct_factor": [_pcsr, _i64, _ptr, _sz, _ptr],
    "psfem_direct_substitute": [_pcsr, _ptr, _ptr, _i64, _i32, _pdbl, _ptr, _sz, _ptr],
    "psfem_fp64_peak": [_pdbl, _ptr, _sz, _ptr],
    "psfem_residual_dd": [_pcsr, _ptr, _ptr, _ptr, _ptr],
    "psfem_csr_symmetry": [_pcsr, _dbl, _dbl, _pi64, _pdbl, _ptr, _sz, _ptr],
    "psfem_cg_sl_usable": [_pcsr, C.POINTER(C.c_int)],
    "psfem_cg_sl_workspace": [_i64, _psz],
    "psfem_cg_sl_start": [_pcsr, _ptr, _ptr, _ptr, _dbl, _ptr, _sz, _ptr],
    "psfem_cg_sl_iterations": [_pcsr, _ptr, _i32, _i64, _i64, _ptr, _sz, _ptr],
    "psfem_cg_sl_state": [_ptr, _i64, _pdbl, _ptr],
    "psfem_cg_sl_finish": [_ptr, _i64, _ptr, _ptr],
    "psfem_cg_sl_solve": [_pcsr, _ptr, _ptr, _ptr, _dbl, _i64, _i64, _pdbl, _ptr, _sz, _ptr],
    "psfem_cg_sl_solve_checked": [_pcsr, _ptr, _ptr, _ptr, _dbl, _i64, _i64, _i64, _pdbl, _ptr, _sz, _ptr],
    "psfem_pcg_init": [_i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "psfem_pcg_spmv_dot": [_pcsr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr],
    "psfem_pcg_update": [_i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "psfem_pcg_update_masked": [_i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "psfem_pcg_pupdate": [_i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "psfem_peer_board_bytes": [_psz],
    "psfem_peer_alloc": [_sz, C.POINTER(C.c_void_p), _ptr],
    "psfem_peer_open": [_ptr, C.POINTER(C.c_void_p)],
    "psfem_peer_close": [_ptr],
    "psfem_peer_free": [_ptr],
    "psfem_dist_pcg_start": [_pctx, _ptr, _u64, _ptr],
    "psfem_dist_pcg_iterations": [_pcsr, _pctx, _i64, _u64, _ptr],
    "psfem_dist_pcg_profile": [_pcsr, _pctx, _i64, _u64, _pdbl, _ptr],
    "psfem_dist_pcg_residual": [_pctx, _u64, _pdbl, _ptr],
    "psfem_dist_cg_halo_bytes": [_i64, _psz],
    "psfem_dist_cg_start": [_pcsr, _pcgctx, _ptr, _ptr, _dbl, _u64, _ptr, _sz, _ptr],
    "psfem_dist_cg_iterations": [_pcsr, _pcgctx, _i64, _i64, _u64, _ptr, _sz, _ptr],
    "psfem_dist_cg_state": [_pcgctx, _i64, _ptr, _pdbl, _ptr],
    "psfem_dist_cg_apply": [_pcsr, _pcgctx, _ptr, _ptr, _u64, _ptr, _sz, _ptr],
    "psfem_axpby": [_i64, _dbl, _ptr, _dbl, _ptr, _ptr, _ptr],
    "psfem_dot_workspace": [_i64, _psz],
    "psfem_dot": [_i64, _ptr, _ptr, _pdbl, _ptr, _sz, _ptr],
    "psfem_multi_dot": [_i64, _i64, _ptr, _ptr, _ptr, _ptr],
    "psfem_multi_axpy": [_i64, _i64, _ptr, _ptr, _ptr, _ptr],
    "psfem_newmark_workspace": [_i64, _i64, _psz],
    "psfem_newmark_run": [_pcsr, _pcsr, _pcsr, _ptr, _pdbl, _i64, _dbl, _dbl, _dbl, _dbl, _dbl, _ptr, _i64,
                          _ptr, _ptr, _ptr, _ptr, _dbl, _i64, _pdbl, _ptr, _sz, _ptr],
    "psfem_explicit_run": [_pcsr, _pcsr, _ptr, _pdbl, _i64, _dbl, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _dbl, _i64, _pdbl,
                           _ptr, _sz, _ptr],
}

for _name, _args in SIGNATURES.items():
    _f = getattr(lib, _name)      # AttributeError here = library / header mismatch: fail loudly
    _f.argtypes = _args
    _f.restype = C.c_int
lib.psfem_last_error.argtypes = []
lib.psfem_last_error.restype = C.c_char_p


def last_error() -> str:
    return (lib.psfem_last_error() or b"").decode()


def check(rc: int, bad_element: int | None = None):
    """Map a C return code to the exception the reference would raise."""
    if rc == OK:
        return
    msg = last_error()
    if rc == EJACOBIAN:
        # Polygones.py:686-687
        raise ValueError(f"Jacobian determinant near zero for element {bad_element}" if bad_element is not None else msg)
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOCONV:
        raise ConvergenceError(msg)
    raise RuntimeError(f"libpsfem CUDA error: {msg}")


def call(name: str, *args):
    check(getattr(lib, name)(*args))


# ---- torch plumbing: device memory + current stream ------------------------------------------
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("psfem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


_bound_device = None


def bind_device():
    """Make the library's CUDA runtime follow torch's current device."""
    global _bound_device
    torch = torch_cuda()
    dev = torch.cuda.current_device()
    if dev != _bound_device:
        call("psfem_set_device", dev)
        _bound_device = dev
    return dev


def stream_ptr():
    torch = torch_cuda()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def workspace(nbytes: int):
    torch = torch_cuda()
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device="cuda")
