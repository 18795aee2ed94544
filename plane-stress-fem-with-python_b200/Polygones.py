"""Drop-in for the hot-path functions of the reference's `Polygones.py` (same names, argument
order, defaults and error behaviour), running on hand-written sm_100a kernels behind the C ABI of
libpsfem.so.  `import Polygones` semantics are kept: `mesh`, `global_stiffness_matrix`,
`global_mass_matrix`, `element_*_matrix`, `apply_boundary_conditions`, `reduce`,
`reconstruct_full_vector(s)`, `reconstruct_full_matrix`, `boundary`, `merge_boudary_conditions`,
`edgeForces`, `displacement`, `get_deformed_nodes_list`.

Differences a user sees, all forced by scale:
  * K and M come back as `scipy.sparse.csr_matrix` on the STRUCTURAL pattern (explicit zeros kept,
    SURVEY A.6), not as dense (2n,2n) arrays (Polygones.py:715, :1349);
  * `elements` is an array-backed Mapping (`ElementTable`) that behaves like the reference's
    dict[int, list[int]]; plain dicts are accepted everywhere;
  * `integration='symbolic'` (Polygones.py:597-650; never used by any reference script) is not
    implemented and raises NotImplementedError.
The returned matrices carry a handle to their device-resident copy (`_psfem_dev`), which the
drivers in `drivers.py` use instead of re-uploading.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sparse

from . import _capi
from ._capi import ETYPE, NPE
from .device import (Assembler, DeviceCSR, ElementTable, FreeMap, as_element_table, constrained_dofs_of,
                     device_mesh)

__all__ = ["shape_functions", "get_shape_derivatives", "create_shape_functions", "compute_B_matrix_at_point",
           "mesh", "global_stiffness_matrix", "global_stiffness_matrix_linear", "get_gauss_points", "global_mass_matrix", "element_stiffness_matrix",
           "element_mass_matrix", "apply_boundary_conditions", "reduce", "reconstruct_full_vector",
           "reconstruct_full_vectors", "reconstruct_full_matrix", "boundary", "merge_boudary_conditions",
           "edgeForces", "body_forces", "convert_to_banded", "convert_to_banded_optimized", "calculate_bandwidth",
           "calculate_bandwidth_optimized", "normalize_matrix", "shapeFunctionCoeff", "shapeFunction", "strainDisplacement",
           "local_stiffness_matrix", "localMassMtrix", "displacement", "get_deformed_nodes_list", "globalEnergy", "ElementTable"]


def _etype(element_type, mesh_type):
    try:
        return ETYPE[(element_type, mesh_type)]
    except KeyError:
        raise ValueError(f"unknown element_type/mesh_type {element_type!r}/{mesh_type!r}") from None


def _check_integration(integration):
    # Polygones.py:590-653
    if integration == "gauss":
        return
    if integration == "symbolic":
        raise NotImplementedError("integration='symbolic' is outside the GPU hot path (no reference script uses it)")
    raise ValueError("Integration method must be either 'gauss' or 'symbolic'")


# ---------------------------------------------------------------------------------------------
# mesh  (Polygones.py:62-108)
# ---------------------------------------------------------------------------------------------
def mesh(L, l, N, M=0, offsetX=0, offsetY=0, element_type='tri', mesh_type='coarse'):
    """Structured rectangular mesh; returns (nodes (n*m,2) float64 ndarray, elements Mapping).
    Coordinates are bit-identical to Polygones.py:77, connectivity to :83-106."""
    torch = _capi.torch_cuda()
    et = _etype(element_type, mesh_type)
    nodes_dev, conn_dev = device_mesh(L, l, N, M, offsetX, offsetY, element_type, mesh_type)
    torch.cuda.current_stream().synchronize()
    nodes = nodes_dev.cpu().numpy()
    elements = ElementTable(conn_dev.cpu().numpy())
    elements._dev = conn_dev
    return nodes, elements


# ---------------------------------------------------------------------------------------------
# assembly  (Polygones.py:710-728, :1344-1364)
# ---------------------------------------------------------------------------------------------
def _assembler(elements, n_nodes, et) -> tuple[Assembler, ElementTable]:
    tab = as_element_table(elements, NPE[et])
    plan = tab._plan
    if plan is None or plan.n_nodes != n_nodes or plan.etype != et:
        plan = Assembler(tab.device(), n_nodes, et)
        tab._plan = plan
    return plan, tab


class DeviceBackedCSR(sparse.csr_matrix):
    """The `scipy.sparse.csr_matrix` the assembly functions return: an ordinary CSR matrix for the caller that also
    remembers its device-resident copy, so that the drivers do not upload 7 GB again.

    The cached copy must never go stale (a penalty boundary condition `K[i, i] = 1e20` after assembly has to reach the
    solver), so the value array is handed out READ-ONLY and every mutating method of this class (`K[i, j] = v`,
    `setdiag`, `K *= a`, `K /= a`, `eliminate_zeros`, `sum_duplicates`, `sort_indices`, `prune`, `resize`) first drops
    the device copy and makes the values writeable again.  A direct write `K.data[k] = v` raises numpy's
    "assignment destination is read-only" instead of being lost; `K.data.flags.writeable = True` (or assigning a new
    `K.data`) is seen by `device_copy`, which then uploads the host values."""

    _psfem_dev = None
    _psfem_key = None

    def _psfem_attach(self, dev):
        self.data.flags.writeable = False
        self._psfem_dev = dev
        self._psfem_key = (self.data.ctypes.data, self.data.size, self.indices.ctypes.data, self.indptr.ctypes.data)
        return self

    def _psfem_valid(self):
        return (self._psfem_dev is not None and not self.data.flags.writeable and self._psfem_key ==
                (self.data.ctypes.data, self.data.size, self.indices.ctypes.data, self.indptr.ctypes.data))

    def _psfem_touch(self):
        self._psfem_dev = None
        self._psfem_key = None
        data = self.__dict__.get("data")        # (scipy's constructor calls prune() before / while the arrays are set)
        if data is not None and not data.flags.writeable:
            try:
                data.flags.writeable = True
            except ValueError:                  # a view of a buffer that is itself read-only: take a private copy
                self.data = data.copy()


def _mutator(name):
    base = getattr(sparse.csr_matrix, name)

    def method(self, *args, **kwargs):
        self._psfem_touch()
        return base(self, *args, **kwargs)
    method.__name__ = name
    method.__doc__ = base.__doc__
    return method


for _name in ("__setitem__", "setdiag", "__imul__", "__itruediv__", "__iadd__", "__isub__", "eliminate_zeros", "sum_duplicates",
              "sort_indices", "prune", "resize"):
    if hasattr(sparse.csr_matrix, _name):
        setattr(DeviceBackedCSR, _name, _mutator(_name))


def _to_host_csr(dev: DeviceCSR):
    A = dev.to_scipy(cls=DeviceBackedCSR)
    return A._psfem_attach(dev)


def device_copy(A) -> DeviceCSR:
    """Device CSR of a matrix: the copy cached at assembly while the host matrix is provably untouched (see
    DeviceBackedCSR), else an upload of the host arrays."""
    if isinstance(A, DeviceCSR):
        return A
    if isinstance(A, DeviceBackedCSR) and A._psfem_valid():
        return A._psfem_dev
    return DeviceCSR.from_scipy(sparse.csr_matrix(A))


def global_stiffness_matrix(nodes, elements, E, nu, t, element_type='tri', mesh_type='coarse', integration='gauss'):
    """Global stiffness matrix as CSR on the structural pattern (Polygones.py:710-728)."""
    _check_integration(integration)
    et = _etype(element_type, mesh_type)
    torch = _capi.torch_cuda()
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    plan, _ = _assembler(elements, len(nodes), et)
    K_vals, _ = plan.assemble(torch.from_numpy(nodes).cuda(), E=E, nu=nu, t=t, what=1)
    return _to_host_csr(plan.pattern_csr(K_vals))


def global_stiffness_matrix_linear(nodes, element, E, nu, t):
    """The reference's T3-only assembly variant (Polygones.py:885-896, no caller): same matrix as
    global_stiffness_matrix(..., 'tri', 'coarse')."""
    return global_stiffness_matrix(nodes, element, E, nu, t, 'tri', 'coarse')


def get_gauss_points(element_type='tri'):
    """Quadrature rule of the reference (Polygones.py:534-564): (points, weights); the triangle weights sum to 1,
    not 1/2, exactly as there; 3x3 Gauss-Legendre for quadrilaterals (eta outer, xi inner)."""
    if element_type == 'tri':
        return [(1 / 6, 1 / 6), (2 / 3, 1 / 6), (1 / 6, 2 / 3)], [1 / 3, 1 / 3, 1 / 3]
    a = np.sqrt(0.6)                       # every other element_type is a quadrilateral there (`else:`, :547)
    w = [5 / 9, 8 / 9, 5 / 9]
    g = [-a, 0, a]
    return [(g[c], g[r]) for r in range(3) for c in range(3)], [w[r] * w[c] for r in range(3) for c in range(3)]


# ---------------------------------------------------------------------------------------------
# shape functions (Polygones.py:441-532) and the single-point B matrix (:657-708)
# ---------------------------------------------------------------------------------------------
def _shape_tables(element_type, mesh_type):
    """Numeric (N, dN/dxi, dN/deta)(xi, eta) of the four element kinds: the closed forms the device kernels
    integrate (csrc/psfem_assemble.cu shape_eval_host), node order as Polygones.py:457-497."""
    et = _etype(element_type, mesh_type)
    if et == _capi.T3:
        def f(xi, eta):
            z = 0 * (xi + eta)
            return ([1 - xi - eta, xi + z, eta + z], [z - 1, z + 1, z], [z - 1, z, z + 1])
    elif et == _capi.Q4:
        sx, sy = (-1, 1, 1, -1), (-1, -1, 1, 1)

        def f(xi, eta):
            return ([(1 + a * xi) * (1 + b * eta) / 4 for a, b in zip(sx, sy)],
                    [a * (1 + b * eta) / 4 + 0 * xi for a, b in zip(sx, sy)],
                    [b * (1 + a * xi) / 4 + 0 * eta for a, b in zip(sx, sy)])
    elif et == _capi.T6:
        def f(xi, eta):
            s = 1 - xi - eta
            z = 0 * (xi + eta)
            return ([s * (2 * s - 1), xi * (2 * xi - 1) + z, eta * (2 * eta - 1) + z, 4 * xi * s, 4 * xi * eta, 4 * eta * s],
                    [4 * xi + 4 * eta - 3, 4 * xi - 1 + z, z, 4 - 8 * xi - 4 * eta, 4 * eta + z, -4 * eta + z],
                    [4 * xi + 4 * eta - 3, z, 4 * eta - 1 + z, -4 * xi + z, 4 * xi + z, 4 - 4 * xi - 8 * eta])
    else:
        ab = ((0, 0), (2, 0), (2, 2), (0, 2), (1, 0), (2, 1), (1, 2), (0, 1), (1, 1))   # corners, mid-sides, centre (:487-497)

        def f(xi, eta):
            fx = (xi * (xi - 1) / 2, 1 - xi * xi, xi * (xi + 1) / 2)
            fy = (eta * (eta - 1) / 2, 1 - eta * eta, eta * (eta + 1) / 2)
            gx = (xi - 0.5, -2 * xi, xi + 0.5)
            gy = (eta - 0.5, -2 * eta, eta + 0.5)
            return ([fx[a] * fy[b] for a, b in ab], [gx[a] * fy[b] for a, b in ab], [fx[a] * gy[b] for a, b in ab])
    return f


def shape_functions(element_type='tri', mesh_type='coarse'):
    """(N, xi, eta): the shape functions as sympy expressions in the symbols xi, eta, like Polygones.py:441-502 (so
    `expr.subs({'xi': .., 'eta': ..})` keeps working).  Built by evaluating the closed forms on the symbols; sympy is
    imported here only -- the assembly never touches it (the reference spends 97 % of its Q4 time in sympy.simplify)."""
    import sympy as sp
    xi, eta = sp.symbols('xi eta')
    N, _, _ = _shape_tables(element_type, mesh_type)(xi, eta)
    return [sp.expand(sp.nsimplify(n)) for n in N], xi, eta


def get_shape_derivatives(element_type='tri', mesh_type='coarse'):
    """(dN_dxi, dN_deta) as sympy expressions (Polygones.py:504-518)."""
    import sympy as sp
    N, xi, eta = shape_functions(element_type, mesh_type)
    return [sp.diff(n, xi) for n in N], [sp.diff(n, eta) for n in N]


def create_shape_functions(element_type='tri', mesh_type='coarse'):
    """(N_funcs, dN_dxi_funcs, dN_deta_funcs): lists of NumPy callables f(xi, eta) (Polygones.py:520-532), here plain
    closed forms instead of lambdified sympy expressions."""
    f = _shape_tables(element_type, mesh_type)
    n = NPE[_etype(element_type, mesh_type)]

    def pick(part, i):
        return lambda xi, eta: f(np.asarray(xi, dtype=np.float64), np.asarray(eta, dtype=np.float64))[part][i] + 0.0
    return [pick(0, i) for i in range(n)], [pick(1, i) for i in range(n)], [pick(2, i) for i in range(n)]


def compute_B_matrix_at_point(nodes, elements, element_number, xi, eta, element_type='tri', mesh_type='coarse'):
    """(B, J, detJ) of one element at the reference point (xi, eta), Polygones.py:657-708: B is (3, 2*npe) with rows
    eps_x, eps_y, gamma_xy, J the 2x2 Jacobian [[dx/dxi, dy/dxi], [dx/deta, dy/deta]].  Evaluated on the device
    (psfem_b_matrix_at_point); raises the reference's ValueError when |detJ| < 1e-10 (:686-687)."""
    import ctypes as C
    torch = _capi.torch_cuda()
    _capi.bind_device()
    et = _etype(element_type, mesh_type)
    npe = NPE[et]
    conn = np.asarray(elements[element_number], dtype=np.int32).reshape(-1)
    if conn.size != npe:
        raise ValueError(f"element {element_number} has {conn.size} nodes, expected {npe}")
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    if conn.min() < 0 or conn.max() >= len(nodes):
        raise IndexError(f"element {element_number} refers to a node outside the node array")
    # only the element's own nodes travel to the device
    pts = torch.from_numpy(np.ascontiguousarray(nodes[conn])).cuda()
    local = torch.arange(npe, dtype=torch.int32, device="cuda")
    out = torch.empty(6 * npe + 6, dtype=torch.float64, device="cuda")
    h = (C.c_double * (6 * npe + 6))()
    rc = _capi.lib.psfem_b_matrix_at_point(_capi.ptr(pts), _capi.ptr(local), et, float(xi), float(eta), _capi.ptr(out), h,
                                           _capi.stream_ptr())
    _capi.check(rc, element_number)
    v = np.frombuffer(h, dtype=np.float64).copy()
    return v[:6 * npe].reshape(3, 2 * npe), v[6 * npe:6 * npe + 4].reshape(2, 2), float(v[6 * npe + 4])


def global_mass_matrix(nodes, elements, rho, t, element_type='tri', mesh_type='coarse', integration='gauss'):
    """Global mass matrix on the same pattern as K (Polygones.py:1344-1364)."""
    _check_integration(integration)
    et = _etype(element_type, mesh_type)
    torch = _capi.torch_cuda()
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    plan, _ = _assembler(elements, len(nodes), et)
    _, M_vals = plan.assemble(torch.from_numpy(nodes).cuda(), rho=rho, t=t, what=2)
    return _to_host_csr(plan.pattern_csr(M_vals))


def _single_element(nodes, elements, element_number, et, **mat):
    torch = _capi.torch_cuda()
    conn = np.asarray(elements[element_number], dtype=np.int32).reshape(1, -1)
    if conn.shape[1] != NPE[et]:
        raise ValueError(f"element {element_number} has {conn.shape[1]} nodes, expected {NPE[et]}")
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    asm = Assembler.__new__(Assembler)          # no pattern needed for one element matrix
    asm.etype, asm.npe, asm.nel, asm.n_nodes = et, NPE[et], 1, len(nodes)
    asm.conn = torch.from_numpy(conn).cuda()
    asm._asm_ws = {}
    try:
        return asm.element_matrices(torch.from_numpy(nodes).cuda(), **mat)
    except ValueError as e:
        if "Jacobian determinant near zero" in str(e):      # report the caller's element number
            raise ValueError(f"Jacobian determinant near zero for element {element_number}") from None
        raise


def element_stiffness_matrix(nodes, elements, element_number, E, nu, t, element_type='tri', mesh_type='coarse',
                             integration='gauss'):
    """(2npe,2npe) element stiffness (Polygones.py:566-655; T3 closed form :875-883)."""
    _check_integration(integration)
    Ke, _ = _single_element(nodes, elements, element_number, _etype(element_type, mesh_type), E=E, nu=nu, t=t, what=1)
    return Ke[0].cpu().numpy()


def element_mass_matrix(nodes, elements, element_number, rho, t, element_type='tri', mesh_type='coarse',
                        integration='gauss'):
    """(2npe,2npe) element mass (Polygones.py:1250-1342; T3 :1228-1235)."""
    _check_integration(integration)
    _, Me = _single_element(nodes, elements, element_number, _etype(element_type, mesh_type), rho=rho, t=t, what=2)
    return Me[0].cpu().numpy()


# ---------------------------------------------------------------------------------------------
# boundary conditions  (Polygones.py:898-957, :989-1006)
# ---------------------------------------------------------------------------------------------
def _free_dofs_host(total_dofs, constrained_dofs):
    """Ascending complement of the constrained list (Polygones.py:1000/:901/:916) in O(n)."""
    mask = np.ones(total_dofs, dtype=bool)
    cd = np.asarray(list(constrained_dofs), dtype=np.int64)
    cd = cd[(cd >= 0) & (cd < total_dofs)]
    mask[cd] = False
    return np.nonzero(mask)[0]


def _reduce_matrix(A, constrained_dofs):
    if sparse.issparse(A):
        fm = FreeMap(A.shape[0], constrained_dofs)
        red = fm.reduce_csr(device_copy(A))
        return _to_host_csr(red)
    free = _free_dofs_host(A.shape[0], constrained_dofs)
    return A[np.ix_(free, free)]


def apply_boundary_conditions(K, F, bc):
    """(K_reduced, F_reduced, constrained_dofs) exactly as Polygones.py:989-1006: constrained_dofs in
    bc order with duplicates kept and values ignored; free DOFs ascending.  CSR in -> CSR out."""
    constrained_dofs = constrained_dofs_of(bc)
    K_reduced = _reduce_matrix(K, constrained_dofs)
    F_reduced = np.asarray(F)[_free_dofs_host(K.shape[0], constrained_dofs)]
    return K_reduced, F_reduced, constrained_dofs


def reduce(M, constrained_dofs):
    """Polygones.py:898-911."""
    if M.ndim == 1:
        return M[_free_dofs_host(M.shape[0], constrained_dofs)]
    elif M.ndim == 2:
        return _reduce_matrix(M, constrained_dofs)
    else:
        raise ValueError("Input must be either a vector (1D) or matrix (2D)")


def reconstruct_full_vector(reduced_vector, constrained_dofs, total_dofs):
    """Polygones.py:913-921."""
    full_vector = np.zeros(total_dofs)
    full_vector[_free_dofs_host(total_dofs, constrained_dofs)] = reduced_vector
    return full_vector


def reconstruct_full_vectors(reduced_vectors, constrained_dofs, total_dofs):
    """Polygones.py:923-936 -- note the (modes, total_dofs) result shape."""
    num_modes = reduced_vectors.shape[1]
    full_vectors = np.zeros((num_modes, total_dofs))
    full_vectors[:, _free_dofs_host(total_dofs, constrained_dofs)] = np.asarray(reduced_vectors).T
    return full_vectors


def reconstruct_full_matrix(reduced_matrix, constrained_dofs, total_dofs):
    """Polygones.py:938-957; CSR in -> CSR out, dense in -> dense out."""
    free = _free_dofs_host(total_dofs, constrained_dofs)
    if sparse.issparse(reduced_matrix):
        R = sparse.coo_matrix(reduced_matrix)
        return sparse.csr_matrix((R.data, (free[R.row], free[R.col])), shape=(total_dofs, total_dofs))
    full_matrix = np.zeros((total_dofs, total_dofs))
    full_matrix[np.ix_(free, free)] = reduced_matrix
    return full_matrix


def _counts(N, M, mesh_type):
    if M == 0:
        M = N
    if mesh_type == 'coarse':
        return N, M, N + 1, M + 1
    return N, M, 2 * N + 1, 2 * M + 1


def _edge_nodes(direct, N, M, mesh_type):
    N, M, n, m = _counts(N, M, mesh_type)
    if direct == 'top':
        return np.arange(n) * m + (m - 1)
    if direct == 'bottom':
        return np.arange(n) * m
    if direct == 'left':
        return np.arange(m)
    if direct == 'right':
        return np.arange(m) + (n - 1) * m
    return np.arange(0)


def boundary(direct, disp, N, M=0, element_type='tri', mesh_type='coarse'):
    """List of (node, disp) along one edge (Polygones.py:1008-1068)."""
    return [(int(k), disp) for k in _edge_nodes(direct, N, M, mesh_type)]


def merge_boudary_conditions(bc1, bc2):
    """Polygones.py:1070-1077 (order-preserving union; name kept with the reference's spelling)."""
    bc = list(bc1)
    seen = set()
    hashable = True
    try:
        seen = {(n, tuple(d)) for n, d in bc}
    except TypeError:
        hashable = False
    for i in bc2:
        if hashable:
            key = (i[0], tuple(i[1]))
            if key not in seen:
                seen.add(key)
                bc.append(i)
        elif i not in bc:
            bc.append(i)
    return bc


def edgeForces(F, force, dir, l, N, M=0, element_type='tri', mesh_type='coarse'):
    """Adds f*l/(nodes_on_edge-1) to EVERY node of the edge, corners included; mutates and returns F
    (Polygones.py:1079-1149)."""
    N_, M_, n, m = _counts(N, M, mesh_type)
    if dir in ['top', 'bottom']:
        force_per_node = [f * l / (n - 1) for f in force]
    else:
        force_per_node = [f * l / (m - 1) for f in force]
    ids = _edge_nodes(dir, N, M, mesh_type)
    F[2 * ids] += force_per_node[0]
    F[2 * ids + 1] += force_per_node[1]
    return F


def body_forces(nodes, elements, force_vector, body_force, t, element_type='tri', mesh_type='coarse'):
    """Adds the consistent nodal forces of a uniform body force (fx, fy) to force_vector and returns it, like
    Polygones.py:1151-1200 (fe[2i+p] = sum_gp w N_i f_p t |detJ|, elements in ascending order; no density).

    On the device this is one mass assembly and one SpMV: sum_gp w N_i |detJ| t is the i-th row sum of the element
    mass matrix with rho = 1 (the shape functions sum to one), so the added vector is M(rho=1) @ [fx, fy, fx, fy, ...].
    The reference's 'tri'/'coarse' branch integrates with its triangle rule whose weights sum to 1 instead of 1/2
    (Polygones.py:543-546) while its T3 mass is the closed form: its T3 body force is exactly twice M @ g, reproduced
    here (tests: golden vectors of the reference for all four element types)."""
    et = _etype(element_type, mesh_type)
    torch = _capi.torch_cuda()
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    plan, _ = _assembler(elements, len(nodes), et)
    _, M_vals = plan.assemble(torch.from_numpy(nodes).cuda(), rho=1.0, t=t, what=2)
    g = torch.tensor([float(body_force[0]), float(body_force[1])], dtype=torch.float64, device="cuda").repeat(len(nodes))
    c = 2.0 if et == _capi.T3 else 1.0
    add = plan.pattern_csr(M_vals).matvec(g, alpha=c)
    torch.cuda.current_stream().synchronize()
    force_vector += add.cpu().numpy()
    return force_vector


# ---------------------------------------------------------------------------------------------
# banded helpers of the reference's scripts (Polygones.py:959-987, :1205-1216), CSR-aware, O(nnz) on the host
# ---------------------------------------------------------------------------------------------
def _coo_parts(K):
    if sparse.issparse(K):
        C = sparse.coo_matrix(K)
        return C.row.astype(np.int64), C.col.astype(np.int64), C.data
    K = np.asarray(K)
    r, c = np.nonzero(K)
    return r.astype(np.int64), c.astype(np.int64), K[r, c]


def convert_to_banded(K, lower_bandwidth, upper_bandwidth):
    """LAPACK band storage ab[u + i - j, j] = K[i, j] of the entries with -lower <= j - i <= upper, what
    scipy.linalg.solve_banded takes (Polygones.py:959-965; 2Dexemple.py:33-36).  Accepts the CSR matrices this
    package returns as well as dense arrays."""
    n = K.shape[0]
    ab = np.zeros((lower_bandwidth + upper_bandwidth + 1, n))
    r, c, v = _coo_parts(K)
    keep = (c - r <= upper_bandwidth) & (r - c <= lower_bandwidth)
    np.add.at(ab, (upper_bandwidth + r[keep] - c[keep], c[keep]), v[keep])     # duplicates of an unsummed sparse matrix add up
    return ab


convert_to_banded_optimized = convert_to_banded      # same result (Polygones.py:967-974)


def calculate_bandwidth(K):
    """(lower, upper) bandwidth over the NON-ZERO VALUES of K (Polygones.py:976-987): explicit zeros of the
    structural CSR pattern do not count, like the reference's `K[i, j] != 0` test on its dense matrix."""
    r, c, v = _coo_parts(K)
    nz = v != 0
    d = c[nz] - r[nz]
    lower = int(max(0, -d.min())) if d.size else 0
    upper = int(max(0, d.max())) if d.size else 0
    return lower, upper


def calculate_bandwidth_optimized(K):
    """Polygones.py:1205-1216: the largest super-diagonal offset holding a non-zero value, returned twice.
    (An all-zero upper triangle makes the reference loop forever; here it gives (0, 0).)"""
    r, c, v = _coo_parts(K)
    up = (v != 0) & (c >= r)
    b = int((c[up] - r[up]).max()) if up.any() else 0
    return b, b


def normalize_matrix(matrix):
    """Polygones.py:1368-1371."""
    max_value = abs(matrix).max() if sparse.issparse(matrix) else np.max(np.abs(matrix))
    if max_value == 0:
        return matrix, max_value
    return matrix / max_value, max_value


# ---------------------------------------------------------------------------------------------
# T3 element helpers of the reference (Polygones.py:840-883, :1228-1235): small host-side closed forms
# ---------------------------------------------------------------------------------------------
def shapeFunctionCoeff(nodes, elements, nbElement):
    """[[a_i, b_i, c_i]] / (signed 2*area): N_i(x, y) = a_i + b_i x + c_i y  (Polygones.py:840-851)."""
    p = np.asarray(nodes, dtype=np.float64)[list(elements[nbElement])]
    x, y = p[:, 0], p[:, 1]
    delta = x[0] * (y[1] - y[2]) + x[1] * (y[2] - y[0]) + x[2] * (y[0] - y[1])
    L = []
    for i in range(3):
        j, k = (i + 1) % 3, (i + 2) % 3
        L.append([(x[j] * y[k] - x[k] * y[j]) / delta, (y[j] - y[k]) / delta, (x[k] - x[j]) / delta])
    return L


def shapeFunction(nodes, elements, nbElement):
    """Callable N(x, y) -> [N_0, N_1, N_2] of a T3 element (Polygones.py:853-860; 2Dexemple.py:40)."""
    L = shapeFunctionCoeff(nodes, elements, nbElement)

    def N(x, y):
        return [a + b * x + c * y for a, b, c in L]
    return N


def strainDisplacement(nodes, elements, nbElement):
    """Constant (3, 6) B matrix of a T3 element (Polygones.py:862-873)."""
    L = shapeFunctionCoeff(nodes, elements, nbElement)
    B = np.zeros((3, 6))
    for i, (_, b, c) in enumerate(L):
        B[0, 2 * i] = b
        B[1, 2 * i + 1] = c
        B[2, 2 * i], B[2, 2 * i + 1] = c, b
    return B


def local_stiffness_matrix(nodes, element, nbElement, E, nu, t):
    """T3 element stiffness (Polygones.py:875-883) -- the device closed form of element_stiffness_matrix."""
    return element_stiffness_matrix(nodes, element, nbElement, E, nu, t, 'tri', 'coarse')


def localMassMtrix(nodes, elements, nbElement, rho, t):
    """T3 element mass (Polygones.py:1228-1235; the reference's spelling)."""
    return element_mass_matrix(nodes, elements, nbElement, rho, t, 'tri', 'coarse')


# ---------------------------------------------------------------------------------------------
# post-processing  (Polygones.py:399-439, :1202-1203)
# ---------------------------------------------------------------------------------------------
def displacement(nodes, U, scale=1):
    """Polygones.py:399-403."""
    U = np.asarray(U)
    deformed_nodes = np.copy(nodes)
    deformed_nodes[:, 0] = nodes[:, 0] + U[0::2] * scale
    deformed_nodes[:, 1] = nodes[:, 1] + U[1::2] * scale
    return deformed_nodes


def get_deformed_nodes_list(nodes, U_reduced, constrained_dofs, scale=1):
    """Polygones.py:405-439."""
    n_steps = U_reduced.shape[1]
    U = np.zeros((2 * len(nodes), n_steps))
    U[_free_dofs_host(U.shape[0], constrained_dofs), :] = U_reduced
    return [displacement(nodes, U[:, i], scale) for i in range(n_steps)]


def globalEnergy(U, K):
    """Polygones.py:1202-1203."""
    return 0.5 * np.dot(U.T, K @ U)
