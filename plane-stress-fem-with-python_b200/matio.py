"""Matrix persistence: the reference's only on-disk format plus a sparse one.

The reference checkpoints K and M as DENSE TEXT -- tab-separated without header
(`pd.DataFrame(K).to_csv('stiffness_matrix.txt', sep='\\t', index=False, header=False)`,
ComputeMatrices.py:34-37, read back by main.py:48-49 and freeModes.py:49-50) or comma-separated
(Beamexemple.py:53-54).  `save_matrix` / `load_matrix` read and write exactly those files, so
`main.py` / `freeModes.py`-style workflows can start from matrices assembled here and vice versa; for
anything beyond toy sizes use the `.npz` form (scipy.sparse CSR, structural zeros kept).

Host-side I/O only: nothing here touches the GPU.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sparse

DENSE_TEXT_LIMIT = 20000      # rows; a dense text dump beyond this is > 3 GB of characters

_SEP = {".txt": "\t", ".tsv": "\t", ".csv": ","}


def _sep_of(path, sep):
    if sep is not None:
        return sep
    ext = os.path.splitext(path)[1].lower()
    if ext not in _SEP:
        raise ValueError(f"cannot infer the format of {path!r}: use .npz, .txt/.tsv (tab) or .csv (comma), or pass sep=")
    return _SEP[ext]


def save_matrix(path, A, sep=None):
    """Write A (scipy sparse or dense) to `path`.

    .npz            -> scipy.sparse.save_npz of the CSR matrix (explicit structural zeros are kept)
    .txt / .tsv     -> dense, tab-separated, no header, no index: ComputeMatrices.py:34-37
    .csv            -> dense, comma-separated, no header, no index: Beamexemple.py:53-54
    Floats are written with repr() (shortest round-trip form, what pandas' to_csv emits), so a file
    written here and read by the reference's `pd.read_csv(..., header=None).to_numpy()` gives back the
    same doubles bit for bit."""
    if os.path.splitext(path)[1].lower() == ".npz":
        sparse.save_npz(path, sparse.csr_matrix(A), compressed=False)
        return
    s = _sep_of(path, sep)
    n = A.shape[0]
    if n > DENSE_TEXT_LIMIT:
        raise ValueError(f"a dense text dump of a {n}x{A.shape[1]} matrix is not sensible; save it as .npz")
    dense_rows = (A[i].toarray().ravel() for i in range(n)) if sparse.issparse(A) else (np.asarray(A[i]).ravel() for i in range(n))
    with open(path, "w") as f:
        for row in dense_rows:
            f.write(s.join(repr(float(v)) for v in row))
            f.write("\n")


def load_matrix(path, sep=None, as_csr=False, pattern=None):
    """Read a matrix written by `save_matrix`, by the reference (ComputeMatrices.py:34-37,
    Beamexemple.py:53-54) or by scipy.sparse.save_npz.

    Dense text comes back as a float64 ndarray like `pd.read_csv(...).to_numpy()` (main.py:48-49), or,
    with as_csr=True, as CSR.  The non-zero pattern of a dense dump is value dependent (exact
    cancellations vanish); pass `pattern=` (a CSR matrix assembled by this
    package for the same mesh) to get the values back on the STRUCTURAL pattern instead."""
    if os.path.splitext(path)[1].lower() == ".npz":
        A = sparse.load_npz(path).tocsr()
        return _on_pattern(A, pattern) if pattern is not None else A
    s = _sep_of(path, sep)
    dense = np.loadtxt(path, delimiter=s, dtype=np.float64, ndmin=2)
    if pattern is not None:
        return _on_pattern(dense, pattern)
    return sparse.csr_matrix(dense) if as_csr else dense


def _on_pattern(A, pattern):
    """Values of A gathered on the sparsity pattern of `pattern` (entries of A outside it must be zero)."""
    P = sparse.csr_matrix(pattern)
    if P.shape != A.shape:
        raise ValueError(f"pattern has shape {P.shape}, matrix has {A.shape}")
    rows = np.repeat(np.arange(P.shape[0]), np.diff(P.indptr))
    if sparse.issparse(A):
        vals = np.asarray(sparse.csr_matrix(A)[rows, P.indices]).ravel()
        outside = sparse.csr_matrix(A).count_nonzero() - np.count_nonzero(vals)
    else:
        vals = np.asarray(A)[rows, P.indices]
        outside = np.count_nonzero(A) - np.count_nonzero(vals)
    if outside:
        raise ValueError(f"{outside} non-zero entries of the matrix lie outside the given pattern")
    return sparse.csr_matrix((vals.astype(np.float64), P.indices.copy(), P.indptr.copy()), shape=P.shape)
