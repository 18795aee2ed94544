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

__all__ = ["mesh", "global_stiffness_matrix", "global_mass_matrix", "element_stiffness_matrix",
           "element_mass_matrix", "apply_boundary_conditions", "reduce", "reconstruct_full_vector",
           "reconstruct_full_vectors", "reconstruct_full_matrix", "boundary", "merge_boudary_conditions",
           "edgeForces", "body_forces", "displacement", "get_deformed_nodes_list", "globalEnergy", "ElementTable"]


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


def _to_host_csr(dev: DeviceCSR):
    A = dev.to_scipy()
    A._psfem_dev = dev
    A._psfem_fingerprint = _fingerprint(A)
    return A


def _fingerprint(A):
    """Cheap staleness check of the cached device copy: buffer address + 64 sampled values."""
    d = A.data
    idx = np.linspace(0, max(d.size - 1, 0), num=min(64, d.size), dtype=np.int64)
    return (d.ctypes.data, d.size, float(d[idx].sum()) if d.size else 0.0)


def device_copy(A) -> DeviceCSR:
    """Device CSR of a matrix: the cached copy made at assembly if still valid, else an upload."""
    if isinstance(A, DeviceCSR):
        return A
    dev = getattr(A, "_psfem_dev", None)
    if dev is not None and getattr(A, "_psfem_fingerprint", None) == _fingerprint(A):
        return dev
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
