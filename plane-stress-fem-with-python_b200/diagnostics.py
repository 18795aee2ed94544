"""Sparse-aware matrix diagnostics.

The reference checks K and M before a dynamic run with dense O(n^3) NumPy calls (main.py:60-94 `check_matrix_properties`:
`np.allclose(K, K.T)`, `np.linalg.cholesky`, `np.linalg.cond`, `np.diag`) and estimates the largest eigenvalue of M^-1 K by
power iteration on a dense inverse (main.py:119-140, :196-198).  These functions return the same quantities for CSR
matrices of any size: the extreme eigenvalues come from the shift-invert Lanczos driver (`drivers.modal`: device SpMV +
Jacobi-PCG solves), the diagonal statistics from the device diagonal, the symmetry test from one sparse transposition.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sparse

from . import _capi
from .Polygones import device_copy
from .drivers import modal


def _identity_like(A):
    return sparse.identity(A.shape[0], dtype=np.float64, format="csr")


def symmetry_defect(A, rtol=1e-5, atol=1e-8):
    """(violations, max |A[r,c] - A[c,r]|) with np.allclose semantics |a - b| <= atol + rtol |b| per stored entry, on the
    device (psfem_csr_symmetry): no transposed copy, no dense matrix."""
    import ctypes as C
    from ._capi import call, ptr, stream_ptr, workspace
    dev = device_copy(A)
    if dev.n_rows != dev.n_cols:
        return dev.nnz, float("inf")
    s = dev.struct()
    bad, worst = C.c_int64(0), C.c_double(0.0)
    ws = workspace(256)
    call("psfem_csr_symmetry", C.byref(s), float(rtol), float(atol), C.byref(bad), C.byref(worst), ptr(ws), 256, stream_ptr())
    return int(bad.value), float(worst.value)


def _is_symmetric(A, rtol=1e-5, atol=1e-8):
    """np.allclose(A, A.T, rtol, atol) (main.py:65-66) for a CSR matrix."""
    return symmetry_defect(A, rtol, atol)[0] == 0


def rigid_body_check(K, nodes, rtol=1e-9):
    """The rigid-body test of an UNCONSTRAINED stiffness matrix (the idea of SparseTest.py:29-36, which applies
    u = ones and expects K u = 0): K must annihilate the two translations and the in-plane rotation
    u = (-(y - yc), x - xc) of the mesh.  Three device SpMVs.  Returns a dict: `residuals` = ||K r|| / (||K||_max ||r||)
    for [translation x, translation y, rotation], `passed` = all below rtol, `patch_test` = np.allclose(K @ ones, 0)
    with the reference's absolute tolerance (1e-8, meaningless for E ~ 1e11 but reported as the script prints it)."""
    torch = _capi.torch_cuda()
    dev = device_copy(K)
    nodes = np.asarray(nodes, dtype=np.float64)
    n2 = 2 * len(nodes)
    if dev.n_rows != n2:
        raise ValueError("rigid_body_check needs the full (unconstrained) matrix of these nodes")
    c = nodes.mean(axis=0)
    modes = np.zeros((3, n2))
    modes[0, 0::2] = 1.0
    modes[1, 1::2] = 1.0
    modes[2, 0::2] = -(nodes[:, 1] - c[1])
    modes[2, 1::2] = nodes[:, 0] - c[0]
    kmax = float(torch.max(torch.abs(dev.vals)).item()) if dev.nnz else 0.0
    res = []
    for r in modes:
        f = dev.matvec(torch.from_numpy(r).cuda())
        res.append(float(torch.linalg.vector_norm(f).item()) / max(kmax * np.linalg.norm(r), 1e-300))
    ones = dev.matvec(torch.ones(n2, dtype=torch.float64, device="cuda"))
    return dict(residuals=res, passed=bool(max(res) <= rtol), patch_test=bool(torch.all(torch.abs(ones) <= 1e-8).item()))


def extreme_eigenvalues(A, tol=1e-10):
    """(lambda_min, lambda_max) of a symmetric positive definite CSR matrix.  lambda_min: smallest eigenvalue of
    (A, I) by shift-invert Lanczos (PCG solves with A); lambda_max: 1 / smallest eigenvalue of (I, A) (Lanczos on A
    itself, the solves with I cost one PCG iteration).  Raises ConvergenceError when a solve with A breaks down (A not
    positive definite)."""
    I = _identity_like(A)
    f_lo, _, _, _ = modal(A, I, [], k=1, tol=tol)
    f_hi, _, _, _ = modal(I, A, [], k=1, tol=tol)
    lam_min = float((2 * np.pi * f_lo[0]) ** 2)
    lam_max = float(1.0 / (2 * np.pi * f_hi[0]) ** 2)
    return lam_min, lam_max


def estimate_max_eigenvalue(K, M, tol=1e-10):
    """Largest eigenvalue of M^-1 K (main.py:119-140 estimates it by power iteration on inv(M) @ K): the reciprocal of
    the smallest eigenvalue of M x = mu K x, one shift-invert Lanczos run whose solves are with the mass matrix."""
    f, _, _, _ = modal(M, K, [], k=1, tol=tol)
    return float(1.0 / (2 * np.pi * f[0]) ** 2)


def critical_time_step(K, M, tol=1e-10):
    """2 / sqrt(lambda_max(M^-1 K)): the stability limit main.py:196-198 derives from its eigenvalue estimate."""
    return 2.0 / np.sqrt(estimate_max_eigenvalue(K, M, tol))


def _definite_and_cond(A, diag):
    if not np.all(diag > 0):
        return False, float("inf")
    try:
        lo, hi = extreme_eigenvalues(A)
    except (_capi.ConvergenceError, RuntimeError, FloatingPointError):
        return False, float("inf")
    if not (lo > 0 and np.isfinite(lo) and np.isfinite(hi)):
        return False, float("inf")
    return True, hi / lo


def check_matrix_properties(K, M):
    """The dictionary of main.py:60-94 for CSR matrices: symmetry (np.allclose semantics, rtol 1e-5, atol 1e-8),
    positive definiteness (a Cholesky factorisation succeeds there; here: positive diagonal and a positive smallest
    eigenvalue), 2-norm condition numbers (lambda_max / lambda_min for these symmetric positive definite matrices),
    smallest diagonal entries and the negative / near-zero diagonal flags."""
    out = {}
    diags = {}
    for name, A in (("K", K), ("M", M)):
        diags[name] = device_copy(A).diagonal().cpu().numpy()
    out["K_symmetric"] = _is_symmetric(K)
    out["M_symmetric"] = _is_symmetric(M)
    pdK, condK = _definite_and_cond(K, diags["K"])
    pdM, condM = _definite_and_cond(M, diags["M"])
    out["K_positive_definite"], out["M_positive_definite"] = pdK, pdM
    out["K_condition_number"], out["M_condition_number"] = condK, condM
    out["K_min_diagonal"] = float(np.min(diags["K"]))
    out["M_min_diagonal"] = float(np.min(diags["M"]))
    out["K_has_negative_diagonal"] = bool(np.any(diags["K"] < 0))
    out["M_has_negative_diagonal"] = bool(np.any(diags["M"] < 0))
    out["K_has_zero_diagonal"] = bool(np.any(np.abs(diags["K"]) < 1e-10))
    out["M_has_zero_diagonal"] = bool(np.any(np.abs(diags["M"]) < 1e-10))
    return out
