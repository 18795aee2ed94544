"""CPU oracle for the plane-stress FEM hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a vectorised NumPy/SciPy restatement of the reference's algorithm
(`/root/reference/Polygones.py`, `main.py`, `freeModes.py`, `TestBeam/Statictest.py`).
It is the *checker* for the CUDA path and the timed CPU baseline of `bench.py`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product package never does; it fails loudly when the CUDA
library is missing.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here against
(1) the three golden stiffness fixtures shipped by the reference
    (`TestBeam/stiffness_matrix.txt`, `TestBeam/stiffness_matrix.csv`, `comparison/K.csv`,
    committed in CSR form under `tests/golden/`), and
(2) outputs of the unmodified reference imported in the build container
    (`tests/golden/make_golden.py` -> `tests/golden/ref_*.npz`) for all four element
    types, K and M, boundary-condition indexing, static displacements, the Newmark-beta
    series and the modal frequencies.

The solve arithmetic of the reference lives in un-vendored NumPy/SciPy (no lockfile in the
reference => no pinned versions; the container has numpy 2.3.5 / scipy 1.18.1):
`np.linalg.solve` (TestBeam/Statictest.py:69), `scipy.linalg.lu_factor/lu_solve`
(main.py:295,314), `scipy.linalg.eigh` (freeModes.py:74).  The oracle calls the same routines.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.sparse as sparse
import scipy.sparse.linalg as spla

NPE = {("tri", "coarse"): 3, ("tri", "fine"): 6, ("quad", "coarse"): 4, ("quad", "fine"): 9}


# --------------------------------------------------------------------------------------
# mesh  (Polygones.py:62-108)
# --------------------------------------------------------------------------------------
def mesh_counts(N, M=0, mesh_type="coarse"):
    """n, m node counts along x / y.  Polygones.py:64-71."""
    if M == 0:
        M = N
    if mesh_type == "coarse":
        return N, M, N + 1, M + 1
    return N, M, 2 * N + 1, 2 * M + 1


def mesh(L, l, N, M=0, offsetX=0, offsetY=0, element_type="tri", mesh_type="coarse"):
    """Structured rectangular mesh: nodes (n*m,2) f64, conn (nel,npe) int64.

    Polygones.py:62-108.  Node k=i*m+j sits at [i*L/(n-1)-offsetX, j*l/(m-1)-offsetY]
    (:77; the product i*L is formed first, then the division, then the subtraction).
    Element ids run over cells row-major (i outer, j inner); T3 emits two per cell.
    """
    N, M, n, m = mesh_counts(N, M, mesh_type)
    i = np.arange(n)
    j = np.arange(m)
    # Python evaluates (i*L)/(n-1) - offsetX with i an int: for int L the product is an exact
    # integer, for float L it is a rounded double product; numpy does the same on f64 as long
    # as i*L < 2**53.
    xs = (i * L) / (n - 1) - offsetX
    ys = (j * l) / (m - 1) - offsetY
    nodes = np.empty((n * m, 2), dtype=np.float64)
    nodes[:, 0] = np.repeat(xs, m)
    nodes[:, 1] = np.tile(ys, n)
    ii, jj = np.meshgrid(np.arange(N), np.arange(M), indexing="ij")
    ii = ii.ravel()
    jj = jj.ravel()
    if element_type == "tri" and mesh_type == "coarse":      # :83,85
        b = jj + ii * m
        even = np.stack([b, b + m, b + 1], axis=1)
        odd = np.stack([b + m + 1, b + 1, b + m], axis=1)
        conn = np.empty((2 * N * M, 3), dtype=np.int64)
        conn[0::2] = even
        conn[1::2] = odd
    elif element_type == "tri":                               # :91,93
        c = 2 * jj + 2 * ii * m
        even = np.stack([c, c + m, c + 2 * m, c + m + 1, c + 2, c + 1], axis=1)
        odd = np.stack([2 * jj + (2 * ii + 2) * m + 2, 2 * jj + (2 * ii + 1) * m + 2,
                        2 * jj + (2 * ii) * m + 2, 2 * jj + (2 * ii + 1) * m + 1,
                        2 * jj + (2 * ii + 2) * m, 2 * jj + (2 * ii + 2) * m + 1], axis=1)
        conn = np.empty((2 * N * M, 6), dtype=np.int64)
        conn[0::2] = even
        conn[1::2] = odd
    elif mesh_type == "coarse":                               # :100
        b = jj + ii * m
        conn = np.stack([b, b + m, b + m + 1, b + 1], axis=1).astype(np.int64)
    else:                                                     # :106
        c = 2 * jj + 2 * ii * m
        conn = np.stack([c, c + m, c + 2 * m, c + 2 * m + 1, c + 2 * m + 2, c + m + 2,
                         c + 2, c + 1, c + m + 1], axis=1).astype(np.int64)
    return nodes, conn


# --------------------------------------------------------------------------------------
# quadrature and shape functions  (Polygones.py:441-564)
# --------------------------------------------------------------------------------------
def gauss_points(element_type):
    """Polygones.py:534-564.  Triangle weights sum to 1 (no 1/2) -- reproduced as is."""
    if element_type == "tri":
        pts = [(1 / 6, 1 / 6), (2 / 3, 1 / 6), (1 / 6, 2 / 3)]
        w1 = 1 / 3
        return np.array(pts), np.array([w1, w1, w1])
    a = np.sqrt(0.6)
    pts = [(-a, -a), (0, -a), (a, -a), (-a, 0), (0, 0), (a, 0), (-a, a), (0, a), (a, a)]
    w1, w2, w3 = 5 / 9, 8 / 9, 5 / 9
    w = [w1 * w1, w1 * w2, w1 * w3, w2 * w1, w2 * w2, w2 * w3, w3 * w1, w3 * w2, w3 * w3]
    return np.array(pts), np.array(w)


def shape_eval(element_type, mesh_type, xi, eta):
    """N, dN/dxi, dN/deta at (xi,eta): the polynomials of Polygones.py:457-497 evaluated
    directly (the reference evaluates the sympy-simplified forms, :676-677,:1277 -- same
    polynomials, differences are rounding-level)."""
    if element_type == "tri" and mesh_type == "coarse":
        N = [1 - xi - eta, xi, eta]
        dxi = [-1.0, 1.0, 0.0]
        det = [-1.0, 0.0, 1.0]
    elif element_type == "tri":
        s = 1 - xi - eta
        N = [s * (2 * s - 1), xi * (2 * xi - 1), eta * (2 * eta - 1), 4 * xi * s, 4 * xi * eta, 4 * eta * s]
        dxi = [4 * xi + 4 * eta - 3, 4 * xi - 1, 0.0, 4 - 8 * xi - 4 * eta, 4 * eta, -4 * eta]
        det = [4 * xi + 4 * eta - 3, 0.0, 4 * eta - 1, -4 * xi, 4 * xi, 4 - 4 * xi - 8 * eta]
    elif mesh_type == "coarse":
        N = [(1 - xi) * (1 - eta) / 4, (1 + xi) * (1 - eta) / 4, (1 + xi) * (1 + eta) / 4, (1 - xi) * (1 + eta) / 4]
        dxi = [-(1 - eta) / 4, (1 - eta) / 4, (1 + eta) / 4, -(1 + eta) / 4]
        det = [-(1 - xi) / 4, -(1 + xi) / 4, (1 + xi) / 4, (1 - xi) / 4]
    else:
        fx = [xi * (xi - 1) / 2, 1 - xi * xi, xi * (xi + 1) / 2]       # 1-D quadratic Lagrange at -1,0,+1
        fy = [eta * (eta - 1) / 2, 1 - eta * eta, eta * (eta + 1) / 2]
        gx = [xi - 0.5, -2 * xi, xi + 0.5]
        gy = [eta - 0.5, -2 * eta, eta + 0.5]
        # (a,b) = 1-D indices of the 9 functions in the order of Polygones.py:488-496
        ab = [(0, 0), (2, 0), (2, 2), (0, 2), (1, 0), (2, 1), (1, 2), (0, 1), (1, 1)]
        N = [fx[a] * fy[b] for a, b in ab]
        dxi = [gx[a] * fy[b] for a, b in ab]
        det = [fx[a] * gy[b] for a, b in ab]
    return np.array(N, dtype=float), np.array(dxi, dtype=float), np.array(det, dtype=float)


def d_matrix(E, nu):
    """Plane-stress D.  Polygones.py:582-584 (and :881)."""
    return np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]]) * (E / (1 - nu ** 2))


class JacobianError(ValueError):
    pass


# --------------------------------------------------------------------------------------
# element matrices, batched over all elements
# --------------------------------------------------------------------------------------
def _t3_delta(xy):
    """signed 2*area.  Polygones.py:843-845 / :877-879."""
    x, y = xy[:, :, 0], xy[:, :, 1]
    return x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1])


def element_stiffness(nodes, conn, E, nu, t, element_type="tri", mesh_type="coarse"):
    """Ke for every element: (nel, 2npe, 2npe).

    T3: closed form with SIGNED area, Polygones.py:840-883.
    Others: sum over Gauss points of w * B^T D B * |detJ| * t, Polygones.py:593-595, B and J
    from :657-708; raises like :686-687 when |detJ| < 1e-10.
    """
    conn = np.asarray(conn)
    xy = nodes[conn]                                   # (nel, npe, 2)
    nel, npe = conn.shape
    D = d_matrix(E, nu)
    if element_type == "tri" and mesh_type == "coarse":
        delta = _t3_delta(xy)
        x, y = xy[:, :, 0], xy[:, :, 1]
        B = np.zeros((nel, 3, 6))
        for i in range(3):
            b = (y[:, (i + 1) % 3] - y[:, (i + 2) % 3]) / delta
            c = (x[:, (i + 2) % 3] - x[:, (i + 1) % 3]) / delta
            B[:, 0, 2 * i] = b
            B[:, 1, 2 * i + 1] = c
            B[:, 2, 2 * i] = c
            B[:, 2, 2 * i + 1] = b
        vol = 0.5 * delta * t
        return np.matmul(B.transpose(0, 2, 1), np.matmul(D, B)) * vol[:, None, None]
    pts, ws = gauss_points(element_type)
    Ke = np.zeros((nel, 2 * npe, 2 * npe))
    for (xi, eta), w in zip(pts, ws):
        _, dxi, det = shape_eval(element_type, mesh_type, xi, eta)
        J00 = xy[:, :, 0] @ dxi
        J01 = xy[:, :, 1] @ dxi
        J10 = xy[:, :, 0] @ det
        J11 = xy[:, :, 1] @ det
        detJ = J00 * J11 - J01 * J10
        bad = np.nonzero(np.abs(detJ) < 1e-10)[0]
        if bad.size:
            raise JacobianError(f"Jacobian determinant near zero for element {int(bad[0])}")
        i00, i01, i10, i11 = J11 / detJ, -J01 / detJ, -J10 / detJ, J00 / detJ
        dx = i00[:, None] * dxi[None, :] + i01[:, None] * det[None, :]
        dy = i10[:, None] * dxi[None, :] + i11[:, None] * det[None, :]
        B = np.zeros((nel, 3, 2 * npe))
        B[:, 0, 0::2] = dx
        B[:, 1, 1::2] = dy
        B[:, 2, 0::2] = dy
        B[:, 2, 1::2] = dx
        Ke += np.matmul(B.transpose(0, 2, 1), np.matmul(D, B)) * (w * np.abs(detJ) * t)[:, None, None]
    return Ke


def element_mass(nodes, conn, rho, t, element_type="tri", mesh_type="coarse"):
    """Me for every element: (nel, 2npe, 2npe).

    T3: template * rho * signedArea * t / 12, Polygones.py:1218-1235.
    Others: sum_gp w * rho * t * N^T N * |detJ|, Polygones.py:1275-1291.
    """
    conn = np.asarray(conn)
    xy = nodes[conn]
    nel, npe = conn.shape
    if element_type == "tri" and mesh_type == "coarse":
        T = np.zeros((6, 6))
        for i in range(3):
            for j in range(3):
                T[2 * i, 2 * j] = T[2 * i + 1, 2 * j + 1] = 2 if i == j else 1
        vol = 0.5 * _t3_delta(xy) * t
        return T[None] * (rho * vol / 12)[:, None, None]
    pts, ws = gauss_points(element_type)
    Me = np.zeros((nel, 2 * npe, 2 * npe))
    for (xi, eta), w in zip(pts, ws):
        Nv, dxi, det = shape_eval(element_type, mesh_type, xi, eta)
        J00 = xy[:, :, 0] @ dxi
        J01 = xy[:, :, 1] @ dxi
        J10 = xy[:, :, 0] @ det
        J11 = xy[:, :, 1] @ det
        detJ = J00 * J11 - J01 * J10
        bad = np.nonzero(np.abs(detJ) < 1e-10)[0]
        if bad.size:
            raise JacobianError(f"Jacobian determinant near zero for element {int(bad[0])}")
        NN = np.zeros((2 * npe, 2 * npe))
        NN[0::2, 0::2] = np.outer(Nv, Nv)
        NN[1::2, 1::2] = np.outer(Nv, Nv)
        Me += NN[None] * (w * rho * t * np.abs(detJ))[:, None, None]
    return Me


def body_forces(nodes, conn, force_vector, body_force, t, element_type="tri", mesh_type="coarse"):
    """Polygones.py:1151-1200: fe[2i+p] = sum_gp w * N_i * f_p * t * |detJ| with the rule of get_gauss_points
    (triangle weights sum to 1, so T3/T6 carry twice the area), added to force_vector in ascending element order
    (mutated and returned, like the reference).  The T3 branch goes through the same Gauss loop (:1165-1193), not
    through the closed form."""
    conn = np.asarray(conn)
    xy = nodes[conn]
    nel, npe = conn.shape
    pts, ws = gauss_points(element_type)
    fe = np.zeros((nel, npe, 2))
    for (xi, eta), w in zip(pts, ws):
        Nv, dxi, det = shape_eval(element_type, mesh_type, xi, eta)
        J00 = xy[:, :, 0] @ dxi
        J01 = xy[:, :, 1] @ dxi
        J10 = xy[:, :, 0] @ det
        J11 = xy[:, :, 1] @ det
        detJ = J00 * J11 - J01 * J10
        bad = np.nonzero(np.abs(detJ) < 1e-10)[0]
        if bad.size:
            raise JacobianError(f"Jacobian determinant near zero for element {int(bad[0])}")
        for p in (0, 1):
            fe[:, :, p] += (w * Nv[None, :] * body_force[p] * t) * np.abs(detJ)[:, None]
    np.add.at(force_vector, 2 * conn.ravel(), fe[:, :, 0].ravel())        # sequential, ascending element id
    np.add.at(force_vector, 2 * conn.ravel() + 1, fe[:, :, 1].ravel())
    return force_vector


# --------------------------------------------------------------------------------------
# global assembly  (Polygones.py:710-728, :1344-1364) as structural CSR
# --------------------------------------------------------------------------------------
def dof_pairs(conn):
    """rows, cols of every element DOF pair, in the reference's scatter order
    (element ascending, then i, j, then the four (p,q) of Polygones.py:723-726)."""
    conn = np.asarray(conn, dtype=np.int64)
    nel, npe = conn.shape
    dofs = np.empty((nel, 2 * npe), dtype=np.int64)
    dofs[:, 0::2] = 2 * conn
    dofs[:, 1::2] = 2 * conn + 1
    rows = np.repeat(dofs, 2 * npe, axis=1).ravel()
    cols = np.tile(dofs, (1, 2 * npe)).ravel()
    return rows, cols


def structural_pattern(conn, n_nodes):
    """row_ptr, col_idx of the structural CSR pattern (explicit zeros kept).  SURVEY A.6."""
    rows, cols = dof_pairs(conn)
    P = sparse.coo_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)),
                          shape=(2 * n_nodes, 2 * n_nodes)).tocsr()
    P.sum_duplicates()
    P.sort_indices()
    return P.indptr.astype(np.int64), P.indices.astype(np.int32)


def assemble(elem_mats, conn, n_nodes):
    """CSR with exactly the structural pattern (explicit zeros stored)."""
    rows, cols = dof_pairs(conn)
    ndof = 2 * n_nodes
    A = sparse.coo_matrix((elem_mats.ravel(), (rows, cols)), shape=(ndof, ndof)).tocsr()
    A.sum_duplicates()      # coo->csr sums duplicates and never prunes explicit zeros
    A.sort_indices()
    return A


def global_stiffness_matrix(nodes, conn, E, nu, t, element_type="tri", mesh_type="coarse"):
    return assemble(element_stiffness(nodes, conn, E, nu, t, element_type, mesh_type), conn, len(nodes))


def global_mass_matrix(nodes, conn, rho, t, element_type="tri", mesh_type="coarse"):
    return assemble(element_mass(nodes, conn, rho, t, element_type, mesh_type), conn, len(nodes))


# --------------------------------------------------------------------------------------
# boundary conditions  (Polygones.py:898-957, :989-1077)
# --------------------------------------------------------------------------------------
def constrained_dofs_of(bc):
    """Polygones.py:992-997: bc order, duplicates kept, values ignored (non-None = fixed to 0)."""
    out = []
    for node, disp in bc:
        if disp[0] is not None:
            out.append(2 * node)
        if disp[1] is not None:
            out.append(2 * node + 1)
    return out


def free_dofs_of(total_dofs, constrained):
    """Ascending complement, Polygones.py:1000 -- O(n) instead of the list scan."""
    mask = np.ones(total_dofs, dtype=bool)
    if len(constrained):
        mask[np.asarray(constrained, dtype=np.int64)] = False
    return np.nonzero(mask)[0]


def apply_boundary_conditions(K, F, bc):
    """Polygones.py:989-1006 on CSR (or dense) K."""
    constrained = constrained_dofs_of(bc)
    free = free_dofs_of(K.shape[0], constrained)
    if sparse.issparse(K):
        K_red = K.tocsr()[free][:, free]
    else:
        K_red = K[np.ix_(free, free)]
    return K_red, np.asarray(F)[free], constrained


def reduce(M, constrained):
    """Polygones.py:898-911."""
    free = free_dofs_of(M.shape[0], constrained)
    if M.ndim == 1:
        return M[free]
    if M.ndim == 2:
        if sparse.issparse(M):
            return M.tocsr()[free][:, free]
        return M[np.ix_(free, free)]
    raise ValueError("Input must be either a vector (1D) or matrix (2D)")


def reconstruct_full_vector(reduced, constrained, total_dofs):
    """Polygones.py:913-921."""
    full = np.zeros(total_dofs)
    full[free_dofs_of(total_dofs, constrained)] = reduced
    return full


def reconstruct_full_vectors(reduced_vectors, constrained, total_dofs):
    """Polygones.py:923-936 -- returns (modes, total_dofs)."""
    free = free_dofs_of(total_dofs, constrained)
    full = np.zeros((reduced_vectors.shape[1], total_dofs))
    full[:, free] = reduced_vectors.T
    return full


def boundary(direct, disp, N, M=0, element_type="tri", mesh_type="coarse"):
    """Polygones.py:1008-1068."""
    N, M, n, m = mesh_counts(N, M, mesh_type)
    if direct == "top":
        ids = [i * m + m - 1 for i in range(n)]
    elif direct == "bottom":
        ids = [i * m for i in range(n)]
    elif direct == "left":
        ids = list(range(m))
    elif direct == "right":
        ids = [i + (n - 1) * m for i in range(m)]
    else:
        ids = []
    return [(k, disp) for k in ids]


def edge_forces(F, force, direction, l, N, M=0, element_type="tri", mesh_type="coarse"):
    """Polygones.py:1079-1149: f*l/(nodes_on_edge-1) added to EVERY edge node (corners too); mutates F."""
    N, M, n, m = mesh_counts(N, M, mesh_type)
    if direction in ("top", "bottom"):
        fpn = [f * l / (n - 1) for f in force]
    else:
        fpn = [f * l / (m - 1) for f in force]
    if direction == "top":
        ids = np.arange(n) * m + (m - 1)
    elif direction == "bottom":
        ids = np.arange(n) * m
    elif direction == "left":
        ids = np.arange(m)
    elif direction == "right":
        ids = np.arange(m) + (n - 1) * m
    else:
        return F
    F[2 * ids] += fpn[0]
    F[2 * ids + 1] += fpn[1]
    return F


# --------------------------------------------------------------------------------------
# drivers
# --------------------------------------------------------------------------------------
def static_solve_dense(K, F, bc):
    """TestBeam/Statictest.py:66-75: reduce, np.linalg.solve, scatter back."""
    K_red, F_red, constrained = apply_boundary_conditions(K, F, bc)
    Kd = K_red.toarray() if sparse.issparse(K_red) else K_red
    U_red = np.linalg.solve(Kd, F_red)
    return reconstruct_full_vector(U_red, constrained, K.shape[0]), U_red, constrained


def jacobi_pcg(A, b, rtol=1e-8, maxit=100000, x0=None):
    """Jacobi-preconditioned CG, x0=0, stop when ||r||_2 <= rtol*||b||_2.
    The same recurrence the CUDA solver runs; used for iteration-count parity and as the
    timed CPU baseline."""
    A = A.tocsr()
    dinv = 1.0 / A.diagonal()
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    z = dinv * r
    p = z.copy()
    rz = r @ z
    bnorm = np.sqrt(b @ b)
    it = 0
    if bnorm == 0.0:
        return x, 0, 0.0
    while it < maxit:
        if np.sqrt(r @ r) <= rtol * bnorm:
            break
        q = A @ p
        alpha = rz / (p @ q)
        x += alpha * p
        r -= alpha * q
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it, float(np.sqrt(r @ r) / bnorm)


def row_blocks(A, nthreads):
    """The row-block split jacobi_pcg_threads multiplies with (set-up, reusable across calls)."""
    A = A.tocsr()
    bounds = np.linspace(0, A.shape[0], nthreads + 1).astype(np.int64)
    return bounds, [A[bounds[i]:bounds[i + 1]] for i in range(nthreads)], 1.0 / A.diagonal()


def jacobi_pcg_threads(A, b, iters, nthreads, split=None):
    """The recurrence of jacobi_pcg for a FIXED number of iterations with the SpMV and the vector
    updates split over `nthreads` row blocks (SciPy's CSR mat-vec and NumPy's element-wise loops
    release the GIL).  Only used to time the CPU side of bench.py with every host core busy.
    Returns (x, relres, seconds spent in the iteration loop -- the row-block split is set-up)."""
    from concurrent.futures import ThreadPoolExecutor
    bounds, blocks, dinv = split if split is not None else row_blocks(A, nthreads)
    sl = [slice(bounds[i], bounds[i + 1]) for i in range(nthreads)]
    x = np.zeros_like(b)
    r = b.copy()
    z = dinv * r
    p = z.copy()
    q = np.empty_like(b)
    rz = r @ z
    with ThreadPoolExecutor(nthreads) as pool:
        def spmv(i):
            q[sl[i]] = blocks[i] @ p
            return p[sl[i]] @ q[sl[i]]

        def update(i, alpha):
            s = sl[i]
            x[s] += alpha * p[s]
            r[s] -= alpha * q[s]
            z[s] = dinv[s] * r[s]
            return r[s] @ z[s], r[s] @ r[s]

        def pupd(i, beta):
            s = sl[i]
            p[s] = z[s] + beta * p[s]

        import time as _time
        t0 = _time.perf_counter()
        for _ in range(iters):
            pq = sum(pool.map(spmv, range(nthreads)))
            alpha = rz / pq
            parts = list(pool.map(lambda i: update(i, alpha), range(nthreads)))
            rz_new = sum(a for a, _ in parts)
            beta = rz_new / rz
            list(pool.map(lambda i: pupd(i, beta), range(nthreads)))
            rz = rz_new
        seconds = _time.perf_counter() - t0
    return x, float(np.sqrt(sum(c for _, c in parts)) / np.sqrt(b @ b)), seconds


def element_stiffness_threads(nodes, conn, E, nu, t, element_type, mesh_type, nthreads):
    """element_stiffness over `nthreads` element chunks (NumPy's batched matmul releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    conn = np.asarray(conn)
    bounds = np.linspace(0, len(conn), nthreads + 1).astype(np.int64)
    with ThreadPoolExecutor(nthreads) as pool:
        parts = list(pool.map(lambda i: element_stiffness(nodes, conn[bounds[i]:bounds[i + 1]], E, nu, t, element_type, mesh_type),
                              range(nthreads)))
    return np.concatenate(parts, axis=0)


def assemble_chunked(nodes, conn, n_cols, m_cells, E, nu, t, nthreads, col0=0, col1=None, drop_first_node_column=False):
    """global_stiffness_matrix of a Q4 plate mesh (n_cols x m_cells cells, element id = i*m_cells + j, Polygones.py:96-100)
    assembled in slabs of cell columns on `nthreads` threads -- the CPU path with every host core busy, and the only
    way to build the 7.3 GB matrix of the 4096 x 4096 plate without a 17 GB COO intermediate.  A slab assembles every
    element touching its node columns (its own cells + the cell column before them) with `assemble` and keeps the rows
    of its node columns, so every row is complete and identical to the one-shot assembly (tests/test_oracle_golden.py).
    col0/col1: node-column range of the rows to produce (default all).  drop_first_node_column: return
    K[np.ix_(free, free)] for a clamped left edge (Polygones.boundary('left', [0, 0], ...) + apply_boundary_conditions,
    Polygones.py:989-1006) without ever forming K."""
    from concurrent.futures import ThreadPoolExecutor
    m = m_cells + 1
    n_node_cols = n_cols + 1
    col1 = n_node_cols if col1 is None else col1
    first = max(col0, 1 if drop_first_node_column else 0)
    bounds = np.unique(np.linspace(first, col1, max(min(4 * nthreads, col1 - first), 1) + 1).astype(np.int64))
    ndof = 2 * len(nodes)
    shift = 2 * m if drop_first_node_column else 0

    def slab(k):
        c0, c1 = int(bounds[k]), int(bounds[k + 1])                     # node columns whose rows this slab produces
        e0, e1 = max(c0 - 1, 0) * m_cells, min(c1, n_cols) * m_cells     # cells touching them
        Ke = element_stiffness(nodes, conn[e0:e1], E, nu, t, "quad", "coarse")
        A = assemble(Ke, conn[e0:e1], len(nodes))[2 * m * c0:2 * m * c1]
        if shift:
            if c0 <= 1:
                A = A[:, shift:]                                         # rows of node column 1 couple with the clamped column
            else:
                A = sparse.csr_matrix((A.data, A.indices - shift, A.indptr), shape=(A.shape[0], ndof - shift))
        return A

    with ThreadPoolExecutor(nthreads) as pool:
        parts = list(pool.map(slab, range(len(bounds) - 1)))
    K = sparse.vstack(parts, format="csr")
    K.sort_indices()
    return K


def newmark_forces(n_nodes, L, P=1e6, T=10.0, dt=1e-3, accel_ratio=3):
    """Ramp load of main.py:201,208-209,226-236.  Returns times, nbPoints, T0, F_y(t) per node (scalar series)."""
    nb = int(T / dt) + 1
    times = np.linspace(0, T, nb)
    T0 = nb // accel_ratio
    fy = np.empty(nb)
    fy[:T0 + 1] = -(accel_ratio * P * times[:T0 + 1] / T) * L / n_nodes
    fy[T0 + 1:] = -P * L / n_nodes
    return times, nb, T0, fy


def newmark(K_red, M_red, F_red_of_step, n_steps, dt=1e-3, gamma=0.5, beta=0.25, alpha_R=0.1, beta_R=0.01):
    """Newmark-beta exactly as main.py:264-332 implements it (NOT the textbook scheme):
    a0 = M^-1 (F0 - K u0) (:286); K_eff = K + M/(beta dt^2) (:290, no C term); per step
    u_p, v_p (:306-307), F_eff = F_i - K u_p - C v_p with C = alpha_R M + beta_R K (:264-266,:310),
    a = K_eff^-1 F_eff (:314), v = v_p + gamma dt a, u = u_p + beta dt^2 a (:319-321), and the
    data-dependent stabiliser (:324-332).  Dense LU like the reference.
    F_red_of_step(i) -> reduced force vector of step i.  Returns U, V, A (n_free, n_steps)."""
    Kd = K_red.toarray() if sparse.issparse(K_red) else np.asarray(K_red)
    Md = M_red.toarray() if sparse.issparse(M_red) else np.asarray(M_red)
    n = Kd.shape[0]
    U = np.zeros((n, n_steps))
    V = np.zeros((n, n_steps))
    A = np.zeros((n, n_steps))
    A[:, 0] = np.linalg.solve(Md, F_red_of_step(0) - Kd @ U[:, 0])
    C = alpha_R * Md + beta_R * Kd
    K_eff = Kd + 1 / (beta * dt ** 2) * Md
    lu, piv = scipy.linalg.lu_factor(K_eff)
    for i in range(1, n_steps):
        U_pred = U[:, i - 1] + dt * V[:, i - 1] + (0.5 - beta) * dt ** 2 * A[:, i - 1]
        V_pred = V[:, i - 1] + (1 - gamma) * dt * A[:, i - 1]
        F_eff = F_red_of_step(i) - Kd @ U_pred - C @ V_pred
        dA = scipy.linalg.lu_solve((lu, piv), F_eff)
        A[:, i] = dA
        V[:, i] = V_pred + gamma * dt * A[:, i]
        U[:, i] = U_pred + beta * dt ** 2 * A[:, i]
        if i > 10 and np.max(np.abs(U[:, i])) > 100 * np.mean(np.abs(U[:, i - 10:i])):
            s = 10 * np.mean(np.abs(U[:, i - 10:i])) / np.max(np.abs(U[:, i]))
            U[:, i] *= s
            V[:, i] *= s
    return U, V, A


def explicit_euler(K_red, M_red, F_red_of_step, n_steps, dt=1e-3):
    """The modified explicit Euler branch of main.py:334-372 (`use_newmark = False`): a0 = M^-1 (F0 - K u0) (:344);
    per step u_i = u_{i-1} + dt v_{i-1} (:352), a_i = M^-1 (F_i - K u_i) (:355-359), v_i = v_{i-1} + dt/2 (a_{i-1} + a_i)
    (:362) and the stabiliser with threshold 10 (:365-370).  Dense solves like the reference; constant dt (the
    reference takes times[i] - times[i-1] of a linspace)."""
    Kd = K_red.toarray() if sparse.issparse(K_red) else np.asarray(K_red)
    Md = M_red.toarray() if sparse.issparse(M_red) else np.asarray(M_red)
    n = Kd.shape[0]
    U = np.zeros((n, n_steps))
    V = np.zeros((n, n_steps))
    A = np.zeros((n, n_steps))
    lu = scipy.linalg.lu_factor(Md)
    A[:, 0] = scipy.linalg.lu_solve(lu, F_red_of_step(0) - Kd @ U[:, 0])
    for i in range(1, n_steps):
        U[:, i] = U[:, i - 1] + dt * V[:, i - 1]
        A[:, i] = scipy.linalg.lu_solve(lu, F_red_of_step(i) - Kd @ U[:, i])
        V[:, i] = V[:, i - 1] + 0.5 * dt * (A[:, i - 1] + A[:, i])
        if i > 10 and np.max(np.abs(U[:, i])) > 10 * np.mean(np.abs(U[:, i - 10:i])):
            s = 10 * np.mean(np.abs(U[:, i - 10:i])) / np.max(np.abs(U[:, i]))
            U[:, i] *= s
            V[:, i] *= s
    return U, V, A


def energies(K_red, M_red, U, V):
    """main.py:377-385."""
    kin = 0.5 * np.einsum("it,it->t", V, M_red @ V)
    pot = 0.5 * np.einsum("it,it->t", U, K_red @ U)
    return kin, pot, kin + pot


def modal_dense(K, M, bc, k=None):
    """freeModes.py:53-77: reduce, scale by max |diag|, dense generalized eigh, clip, Hz, rebuild modes."""
    K_red, _, constrained = apply_boundary_conditions(K, np.zeros(K.shape[0]), bc)
    M_red = reduce(M, constrained)
    Kd = K_red.toarray() if sparse.issparse(K_red) else K_red
    Md = M_red.toarray() if sparse.issparse(M_red) else M_red
    k_max = np.max(np.abs(np.diag(Kd)))
    m_max = np.max(np.abs(np.diag(Md)))
    scaling = k_max / m_max
    w, v = scipy.linalg.eigh(Kd / k_max, Md / m_max)
    w = np.maximum(w, 0)
    freqs = np.sqrt(w * scaling) / (2 * np.pi)
    modes = reconstruct_full_vectors(v, constrained, K.shape[0])
    if k is not None:
        return freqs[:k], modes[:k], constrained
    return freqs, modes, constrained
