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


def _is_symmetric(A, rtol=1e-5, atol=1e-8):
    """np.allclose(A, A.T, rtol, atol) (main.py:65-66) on the union of the two patterns: |a - b| <= atol + rtol |b|."""
    A = sparse.csr_matrix(A)
    At = A.T.tocsr()
    D = (A - At).tocoo()
    if D.nnz == 0:
        return True
    ref = np.abs(np.asarray(At[D.row, D.col]).ravel())
    return bool(np.all(np.abs(D.data) <= atol + rtol * ref))


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
