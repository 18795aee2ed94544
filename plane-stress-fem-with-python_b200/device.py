"""Device-resident objects of the hot path: element tables, the symbolic pattern / assembly plan,
CSR matrices with their SpMV plan, boundary-condition reduction, PCG.

Everything here calls the C ABI of libpsfem.so through `_capi`; torch only owns the buffers.
"""
from __future__ import annotations

import ctypes as C
import os
from collections.abc import Mapping

import numpy as np
import scipy.sparse as sparse

from . import _capi
from ._capi import ETYPE, NPE, call, check, lib, ptr, stream_ptr, workspace


# --------------------------------------------------------------------------------------------
# elements: array-backed stand-in for the reference's dict[int, list[int]] (Polygones.py:74-107)
# --------------------------------------------------------------------------------------------
class ElementTable(Mapping):
    """Behaves like the reference's `elements` dict (`for e in elements`, `elements[e][k]`,
    `len(elements[e])`, Polygones.py:717-721) but is backed by one (nel, npe) int32 array, so a
    16.7 M-element mesh does not become 16.7 M Python lists."""

    def __init__(self, conn: np.ndarray):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        assert conn.ndim == 2
        self.conn = conn
        self._dev = None          # cached device copy
        self._plan = None         # cached Assembler (symbolic phase), built on first assembly

    def __len__(self):
        return self.conn.shape[0]

    def __iter__(self):
        return iter(range(self.conn.shape[0]))

    def __getitem__(self, e):
        if isinstance(e, (int, np.integer)):
            if e < 0 or e >= self.conn.shape[0]:
                raise KeyError(e)
            return self.conn[e].tolist()
        raise KeyError(e)

    def __repr__(self):
        return f"ElementTable(nel={self.conn.shape[0]}, npe={self.conn.shape[1]})"

    def device(self):
        torch = _capi.torch_cuda()
        if self._dev is None:
            self._dev = torch.from_numpy(self.conn).cuda()
        return self._dev


def as_element_table(elements, npe_expected=None) -> ElementTable:
    """Accept an ElementTable, a plain dict (2Dexemple.py:12-17 passes a literal) or an array.
    A dict is read in ITERATION order, which is the order of the reference's scatter loop
    (Polygones.py:717)."""
    if isinstance(elements, ElementTable):
        tab = elements
    elif isinstance(elements, np.ndarray):
        tab = ElementTable(elements)
    else:
        rows = [elements[k] for k in elements]
        tab = ElementTable(np.array(rows, dtype=np.int32).reshape(len(rows), -1))
    if npe_expected is not None and len(tab) and tab.conn.shape[1] != npe_expected:
        raise ValueError(f"elements have {tab.conn.shape[1]} nodes, element type expects {npe_expected}")
    return tab


PIN_THRESHOLD = 1 << 20    # elements; smaller arrays are not worth a page-locked allocation


def to_host(t):
    """Device tensor -> host tensor (asynchronous into pinned memory when large; caller synchronises)."""
    torch = _capi.torch_cuda()
    if t.numel() < PIN_THRESHOLD:
        return t.cpu()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    return h


def host_array(h):
    """NumPy view of a host tensor whose WRITEABLE flag can be cleared AND set again.  `tensor.numpy()` arrays cannot be
    made writeable once locked (their base is not a buffer exporter), so DeviceBackedCSR -- which hands its values out
    read-only and unlocks them on the first in-place edit -- would have to copy 4.8 GB; a ctypes view of the same
    (page-locked) memory can be toggled.  The tensor stays alive as long as the array does."""
    n = h.numel() * h.element_size()
    if n == 0:
        return h.numpy()
    buf = (C.c_ubyte * n).from_address(h.data_ptr())
    buf._psfem_keep = h
    return np.frombuffer(buf, dtype=h.numpy().dtype).reshape(tuple(h.shape))


def to_device(a, dtype=None):
    """Host ndarray -> device tensor; page-locked sources (e.g. arrays this package returned) go by DMA."""
    import warnings
    torch = _capi.torch_cuda()
    a = np.ascontiguousarray(a, dtype=dtype)
    with warnings.catch_warnings():          # the read-only value arrays of DeviceBackedCSR are only read here
        warnings.filterwarnings("ignore", message="The given NumPy array is not writable")
        t = torch.from_numpy(a)
    return t.to("cuda", non_blocking=False)


# --------------------------------------------------------------------------------------------
# CSR on the device
# --------------------------------------------------------------------------------------------
class DeviceCSR:
    """CSR matrix in HBM: int32 row_ptr/col_idx, fp64 vals, plus the SpMV row-block plan."""

    def __init__(self, n_rows, n_cols, row_ptr, col_idx, vals, nnz=None):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.row_ptr, self.col_idx, self.vals = row_ptr, col_idx, vals
        self.nnz = int(col_idx.numel()) if nnz is None else int(nnz)
        self._plan = None
        self._sl = None           # symmetric-lattice storage of self.vals (sym_lattice)

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    def with_vals(self, vals):
        """Same pattern (and SpMV plan), other values: M, K_eff share K's pattern."""
        o = DeviceCSR(self.n_rows, self.n_cols, self.row_ptr, self.col_idx, vals, self.nnz)
        o._plan = self._plan
        return o

    def plan(self):
        if self._plan is None:
            torch = _capi.torch_cuda()
            _capi.bind_device()
            nb = C.c_int64(0)
            call("psfem_spmv_plan_size", self.n_rows, self.nnz, C.byref(nb))
            rowblk = torch.empty(nb.value + 2, dtype=torch.int32, device="cuda")
            mx = C.c_int64(0)
            call("psfem_spmv_plan", self.n_rows, ptr(self.row_ptr), self.nnz, ptr(rowblk), nb.value, C.byref(mx), stream_ptr())
            self._plan = (rowblk, nb.value, mx.value, self._node_block_plan())
        return self._plan

    def _node_block_plan(self):
        """Node-block view for the 9-byte-per-non-zero SpMV (psfem_spmv_nb_plan); None if the matrix does not
        have the 2x2 node-block structure (e.g. a reduction that removed single DOFs of a node)."""
        torch = _capi.torch_cuda()
        if self.n_rows == 0 or self.n_rows % 2 or self.nnz % 4 or self.nnz == 0:
            return None
        nb = C.c_int64(0)
        call("psfem_spmv_nb_plan_size", self.n_rows, self.nnz, C.byref(nb))
        i32 = dict(dtype=torch.int32, device="cuda")
        node_ptr = torch.empty(self.n_rows // 2 + 9, **i32)
        pcol = torch.empty(self.nnz // 4 + 8, **i32)
        nodeblk = torch.empty(nb.value + 4, **i32)
        info = (C.c_int64 * 3)()
        call("psfem_spmv_nb_plan", self.n_rows, ptr(self.row_ptr), ptr(self.col_idx), self.nnz, ptr(node_ptr), ptr(pcol),
             ptr(nodeblk), nb.value, info, stream_ptr())
        if info[0] != 1:
            return None
        pcol16 = None
        if info[2] < 32768 and os.environ.get("PSFEM_NB_IDX16"):      # 16-bit offsets, 8.5 B per non-zero (opt-in)
            pcol16 = torch.empty(self.nnz // 4 + 16, dtype=torch.int16, device="cuda")
            call("psfem_spmv_nb_compress", self.n_rows, ptr(node_ptr), ptr(pcol), self.nnz, ptr(pcol16), stream_ptr())
            pcol = None
        return (node_ptr, pcol, nodeblk, nb.value, int(info[1]), pcol16)

    SL_ROUNDING_TOL = 1e-14   # |K(r,c) - K(c,r)| <= tol * |diagonal|: asymmetry from rounding only

    def sym_lattice(self, row_node_offset=0, tol=None):
        """Symmetric-lattice storage for the SpMV (psfem_spmv_sl_build): for a symmetric matrix whose nodes couple
        with their 3x3 lattice neighbours only -- every K / M / K + a M of a Polygones.mesh mesh, its reductions by whole
        node columns and the strips of the multi-GPU partition -- the SpMV then streams half the bytes and returns
        the same bits.  Built once per value array; returns False (and changes nothing) for any other matrix.
        row_node_offset: x node of the first row node (owned rows of a strip start after the left halo column).
        tol: 0.0 accepts bitwise symmetric matrices only (the Q4 matrices assembled here; SpMV bit-identical to the CSR
        kernels); a positive value accepts |K(r,c) - K(c,r)| <= tol * |diagonal| and applies the upper triangle for both
        (T3 matrices: the reference's own B^T D B is symmetric up to rounding only, Polygones.py:877-882); None tries
        0.0, then SL_ROUNDING_TOL."""
        if os.environ.get("PSFEM_NO_SYMLATTICE") or self.vals is None:
            return False
        if self._sl is not None and self._sl["src"] == self.vals.data_ptr() and self._sl["roff"] == row_node_offset:
            return True
        self._sl = None
        nbp = self.plan()[3]
        if nbp is None or nbp[1] is None or self.n_cols % 2 or self.n_rows < 64:
            return False
        torch = _capi.torch_cuda()
        s = self.struct()
        x_nodes = self.n_cols // 2
        cand, cnt = (C.c_int64 * 8)(), C.c_int64(0)
        call("psfem_spmv_sl_candidates", C.byref(s), int(row_node_offset), x_nodes, cand, C.byref(cnt), stream_ptr())
        tols = (0.0, self.SL_ROUNDING_TOL) if tol is None else (float(tol),)
        for m, tol in [(m_, t_) for t_ in tols for m_ in list(cand[: cnt.value])]:
            sz = C.c_size_t(0)
            call("psfem_spmv_sl_size", self.n_rows // 2, int(row_node_offset), m, C.byref(sz))
            # the fused PCG kernel reads the storage through 116-row window copies that start two rows before a strip group
            # and end after the last one: 256 bytes of slack in front, 2 KB behind (psfem_cg_sl.cuh)
            nbytes = sz.value + 256
            whole = torch.empty((nbytes + 7) // 8 + 32 + 256, dtype=torch.float64, device="cuda")
            buf = whole[32:]
            ok = C.c_int64(0)
            call("psfem_spmv_sl_build", C.byref(s), int(row_node_offset), x_nodes, m, float(tol), ptr(buf), nbytes, C.byref(ok),
                 stream_ptr())
            if ok.value == 1:
                self._sl = dict(vals=buf, src=self.vals.data_ptr(), m=int(m), roff=int(row_node_offset), x_nodes=x_nodes, tol=tol)
                return True
            del buf
        return False

    def struct(self, vals=None):
        rowblk, nb, mx, nbp = self.plan()
        s = _capi.psfem_csr()
        s.n_rows, s.nnz = self.n_rows, self.nnz
        s.row_ptr, s.col_idx = self.row_ptr.data_ptr(), self.col_idx.data_ptr()
        s.vals = (self.vals if vals is None else vals).data_ptr()
        s.rowblk, s.n_blocks, s.max_row_nnz = rowblk.data_ptr(), nb, mx
        if nbp is not None:
            s.nb_node_ptr, s.nb_nodeblk = nbp[0].data_ptr(), nbp[2].data_ptr()
            s.nb_pcol = nbp[1].data_ptr() if nbp[1] is not None else None
            s.nb_pcol16 = nbp[5].data_ptr() if nbp[5] is not None else None
            s.nb_n_blocks, s.nb_max_pairs = nbp[3], nbp[4]
        sl = self._sl
        if sl is not None and sl["src"] == s.vals:
            s.sl_vals, s.sl_src_vals = sl["vals"].data_ptr(), sl["src"]
            s.sl_m, s.sl_roff, s.sl_x_nodes = sl["m"], sl["roff"], sl["x_nodes"]
        return s

    def to_scipy(self, cls=None):
        """Host scipy CSR.  Large arrays land in page-locked host memory (torch's caching host allocator), so the
        7 GB of a 33.6 M-DOF matrix cross PCIe at DMA speed instead of through the pageable staging path; the
        arrays are ordinary NumPy arrays for the caller (and upload at DMA speed when handed back to a driver)."""
        torch = _capi.torch_cuda()
        parts = [to_host(t) for t in (self.vals, self.col_idx, self.row_ptr)]
        torch.cuda.current_stream().synchronize()
        A = (cls or sparse.csr_matrix)(tuple(host_array(h) for h in parts), shape=self.shape)
        # device-built patterns are sorted and duplicate-free: tell scipy, so that it never scans (or rewrites) the arrays
        A.has_canonical_format = True
        return A

    @classmethod
    def from_scipy(cls, A):
        torch = _capi.torch_cuda()
        _capi.bind_device()
        A = sparse.csr_matrix(A)
        if A.nnz >= 2 ** 31 - 8192:
            raise ValueError("matrix has too many non-zeros for one GPU (int32 indices): partition it")

        def upload(M):
            return cls(M.shape[0], M.shape[1], to_device(M.indptr, np.int32), to_device(M.indices, np.int32),
                       to_device(M.data, np.float64), nnz=M.nnz)
        dev = upload(A)
        # sorted / in-range columns are checked on the device (one pass at HBM speed) instead of scipy's
        # has_sorted_indices host scan -- 0.3 s for the 604 M indices of the 33.6 M-DOF plate
        flags = C.c_int64(0)
        ws = workspace(256)
        call("psfem_csr_check", dev.n_rows, dev.n_cols, ptr(dev.row_ptr), ptr(dev.col_idx), dev.nnz, C.byref(flags), ptr(ws), 256,
             stream_ptr())
        if flags.value & 6:
            raise ValueError("malformed CSR matrix (row pointers not monotone or column index out of range)")
        if flags.value & 1:                       # rare: unsorted or duplicated columns -> canonical form on the host
            A = A.copy()
            A.sum_duplicates()
            A.sort_indices()
            dev = upload(A)
        return dev

    # y = alpha*A@x + gamma*z
    def matvec(self, x, alpha=1.0, gamma=0.0, z=None, out=None):
        torch = _capi.torch_cuda()
        _capi.bind_device()
        if out is None:
            out = torch.empty(self.n_rows, dtype=torch.float64, device="cuda")
        s = self.struct()
        call("psfem_spmv", C.byref(s), ptr(x), float(alpha), float(gamma), ptr(z), ptr(out), stream_ptr())
        return out

    def diagonal(self, invert=False):
        torch = _capi.torch_cuda()
        _capi.bind_device()
        d = torch.empty(self.n_rows, dtype=torch.float64, device="cuda")
        s = self.struct()
        call("psfem_extract_diag", C.byref(s), 0, 1 if invert else 0, ptr(d), stream_ptr())
        return d


# --------------------------------------------------------------------------------------------
# symbolic phase + numeric assembly
# --------------------------------------------------------------------------------------------
class Assembler:
    """Symbolic phase (structural CSR pattern + deterministic assembly plan) of one mesh; reused by
    K, M and every re-assembly.  Replaces the dense target of Polygones.py:715 / :1349."""

    def __init__(self, conn_dev, n_nodes: int, etype: int):
        torch = _capi.torch_cuda()
        _capi.bind_device()
        self.etype, self.npe = etype, NPE[etype]
        self.conn = conn_dev
        self.nel = int(conn_dev.shape[0])
        self.n_nodes = int(n_nodes)
        self._asm_ws = {}
        self._csr = None
        self._tiles = None            # tile plan of the fused T3/Q4 path, built at the first assembly
        self.use_tiles = etype in (_capi.T3, _capi.Q4)
        self.tile_nodes_R = 64        # 8x8-node Morton tiles: best measured on B200 (profiles/r1_tiles.md)
        self.perm = self.pair_row = self.pair_ptr = None
        # lattice topology of Polygones.mesh (any coordinates): closed-form pattern, plan-free marching-warp assembly
        nm = (C.c_int64 * 2)()
        self._small_ws = workspace(256)
        call("psfem_lattice_detect", ptr(conn_dev), self.nel, etype, self.n_nodes, nm, ptr(self._small_ws), 256, stream_ptr())
        self.lattice = (int(nm[0]), int(nm[1])) if nm[0] > 0 else None
        self.use_lattice = self.lattice is not None
        i32 = dict(dtype=torch.int32, device="cuda")
        if self.lattice is not None:
            n, m = self.lattice
            npairs = C.c_int64(0)
            call("psfem_lattice_pattern", n, m, etype, None, None, None, C.byref(npairs), stream_ptr())
            self.n_pairs = npairs.value
            self.nnz = 4 * self.n_pairs
            self.row_ptr = torch.empty(2 * self.n_nodes + 1, **i32)
            self.col_idx = torch.empty(self.nnz, **i32)
            self.node_ptr = torch.empty(self.n_nodes + 1, **i32)
            call("psfem_lattice_pattern", n, m, etype, ptr(self.row_ptr), ptr(self.col_idx), ptr(self.node_ptr),
                 C.byref(npairs), stream_ptr())
        else:
            self._sort_pattern(keep_pattern=True)

    def _sort_pattern(self, keep_pattern):
        """General symbolic phase (radix sort of the node-pair keys): pattern + segmented-reduction plan.
        A lattice already has its pattern in closed form and only needs the plan when the two-kernel path
        is asked for (keep_pattern=False: the sorted pattern is identical and is dropped)."""
        torch = _capi.torch_cuda()
        nk = self.nel * self.npe * self.npe
        sz = C.c_size_t(0)
        call("psfem_pattern_workspace", self.nel, self.npe, self.n_nodes, C.byref(sz))
        ws = workspace(sz.value)
        self.perm = torch.empty(max(nk, 1), dtype=torch.int32, device="cuda")
        npairs = C.c_int64(0)
        call("psfem_pattern_count", ptr(self.conn), self.nel, self.npe, self.n_nodes, ptr(self.perm), ptr(ws), sz.value,
             C.byref(npairs), stream_ptr())
        i32 = dict(dtype=torch.int32, device="cuda")
        n_pairs = npairs.value
        row_ptr = torch.empty(2 * self.n_nodes + 1, **i32)
        col_idx = torch.empty(max(4 * n_pairs, 1), **i32)[: 4 * n_pairs]
        node_ptr = torch.empty(self.n_nodes + 1, **i32)
        self.pair_row = torch.empty(max(n_pairs, 1), **i32)
        self.pair_ptr = torch.empty(n_pairs + 1, **i32)
        call("psfem_pattern_fill", self.nel, self.npe, self.n_nodes, n_pairs, ptr(ws), sz.value, ptr(row_ptr),
             ptr(col_idx), ptr(node_ptr), ptr(self.pair_row), ptr(self.pair_ptr), stream_ptr())
        torch.cuda.current_stream().synchronize()
        del ws
        if keep_pattern:
            self.n_pairs, self.nnz = n_pairs, 4 * n_pairs
            self.row_ptr, self.col_idx, self.node_ptr = row_ptr, col_idx, node_ptr
        else:
            assert n_pairs == self.n_pairs
        return row_ptr, col_idx, node_ptr

    def tile_plan(self, nodes_dev):
        """Symbolic side of the tile-fused assembly (psfem_tileplan_build); needs coordinates once."""
        if self._tiles is None:
            torch = _capi.torch_cuda()
            i32 = dict(dtype=torch.int32, device="cuda")
            R = self.tile_nodes_R
            n_tiles = (self.n_nodes + R - 1) // R
            sz = C.c_size_t(0)
            call("psfem_tileplan_workspace", self.nel, self.npe, self.n_nodes, R, C.byref(sz))
            ws = workspace(sz.value)
            ninc = self.nel * self.npe
            tile_nodes = torch.empty(self.n_nodes, **i32)
            tile_elem_ptr = torch.empty(n_tiles + 1, **i32)
            tile_elems = torch.empty(max(ninc, 1), **i32)
            tile_hdr = torch.empty(8 * n_tiles, **i32)
            tile_conn = torch.empty(4 * max(ninc, 1), **i32)
            tile_rec = torch.empty(32 * max(ninc, 1), dtype=torch.uint8, device="cuda")
            nmeta = torch.empty(2 * n_tiles * R, **i32)
            info = (C.c_int64 * 8)()
            call("psfem_tileplan_build", ptr(nodes_dev), ptr(self.conn), self.nel, self.npe, self.n_nodes, R,
                 ptr(self.node_ptr), ptr(self.col_idx), ptr(tile_nodes), ptr(tile_elem_ptr), ptr(tile_elems), ptr(tile_hdr),
                 ptr(tile_conn), ptr(tile_rec), ptr(nmeta), info, ptr(ws), sz.value, stream_ptr())
            n_te = int(info[0])
            tile_elems = tile_elems[:n_te].clone()                  # drop the upper-bound slack
            tile_conn = tile_conn[: 4 * n_te].clone()
            tile_rec = tile_rec[: 32 * n_te].clone()
            del ws
            self._tiles = dict(R=R, tile_hdr=tile_hdr, tile_conn=tile_conn, tile_rec=tile_rec, tile_elems=tile_elems, nmeta=nmeta,
                               info=info, max_elems=int(info[1]), max_rounds=int(info[3]), mean_elems=int(info[5]),
                               ws=workspace(256))
            if info[4] != 1:
                self.use_tiles = False
        return self._tiles

    def pattern_csr(self, vals):
        """DeviceCSR over this pattern (K and M share row_ptr/col_idx and the SpMV plan)."""
        if self._csr is None:
            self._csr = DeviceCSR(2 * self.n_nodes, 2 * self.n_nodes, self.row_ptr, self.col_idx, vals, self.nnz)
            return self._csr
        return self._csr.with_vals(vals)

    def _ws(self, what):
        if what not in self._asm_ws:
            sz = C.c_size_t(0)
            call("psfem_assemble_workspace", self.nel, self.etype, what, C.byref(sz))
            self._asm_ws[what] = (workspace(sz.value), sz.value)
        return self._asm_ws[what]

    def release_workspace(self):
        self._asm_ws = {}

    def assemble(self, nodes_dev, E=0.0, nu=0.0, rho=0.0, t=1.0, what=1, K_out=None, M_out=None):
        """Numeric phase.  what: 1 = K, 2 = M, 3 = both.  Returns (K_vals, M_vals) device tensors."""
        torch = _capi.torch_cuda()
        _capi.bind_device()
        f64 = dict(dtype=torch.float64, device="cuda")
        K_vals = (K_out if K_out is not None else torch.empty(max(self.nnz, 1), **f64)[: self.nnz]) if what & 1 else None
        M_vals = (M_out if M_out is not None else torch.empty(max(self.nnz, 1), **f64)[: self.nnz]) if what & 2 else None
        bad = C.c_int64(-1)
        if self.use_lattice and self.nel > 0:
            n, m = self.lattice
            rc = lib.psfem_assemble_lattice(ptr(nodes_dev), n, m, self.etype, float(E), float(nu), float(rho), float(t), what,
                                            ptr(K_vals), ptr(M_vals), ptr(self._small_ws), 256, C.byref(bad), stream_ptr())
            check(rc, bad.value)
            return K_vals, M_vals
        if self.use_tiles and self.nel > 0:
            tp = self.tile_plan(nodes_dev)
        if self.use_tiles and self.nel > 0:
            rc = lib.psfem_assemble_tiled(ptr(nodes_dev), self.nel, self.etype, self.n_nodes, float(E), float(nu), float(rho),
                                          float(t), what, tp["R"], tp["info"], ptr(tp["tile_hdr"]), ptr(tp["tile_conn"]),
                                          ptr(tp["tile_rec"]), ptr(tp["tile_elems"]), ptr(tp["nmeta"]),
                                          ptr(K_vals), ptr(M_vals), ptr(tp["ws"]), 256, C.byref(bad), stream_ptr())
            if rc == _capi.EINVAL and "shared memory" in _capi.last_error():
                self.use_tiles = False          # tile too large for one SM: use the two-kernel path below
            else:
                check(rc, bad.value)
                return K_vals, M_vals
        if self.perm is None:
            self._sort_pattern(keep_pattern=False)
        ws, nbytes = self._ws(what)
        rc = lib.psfem_assemble(ptr(nodes_dev), ptr(self.conn), self.nel, self.etype, self.n_nodes, float(E), float(nu),
                                float(rho), float(t), what, self.n_pairs, ptr(self.node_ptr), ptr(self.pair_row),
                                ptr(self.pair_ptr), ptr(self.perm), ptr(K_vals), ptr(M_vals), ptr(ws), nbytes,
                                C.byref(bad), stream_ptr())
        check(rc, bad.value)
        return K_vals, M_vals

    def element_matrices(self, nodes_dev, E=0.0, nu=0.0, rho=0.0, t=1.0, what=1):
        torch = _capi.torch_cuda()
        _capi.bind_device()
        nd = 2 * self.npe
        f64 = dict(dtype=torch.float64, device="cuda")
        Ke = torch.empty((self.nel, nd, nd), **f64) if what & 1 else None
        Me = torch.empty((self.nel, nd, nd), **f64) if what & 2 else None
        ws, nbytes = self._ws(what)
        bad = C.c_int64(-1)
        rc = lib.psfem_element_matrices(ptr(nodes_dev), ptr(self.conn), self.nel, self.etype, float(E), float(nu), float(rho),
                                        float(t), what, ptr(Ke), ptr(Me), ptr(ws), nbytes, C.byref(bad), stream_ptr())
        check(rc, bad.value)
        return Ke, Me


def device_mesh(L, l, N, M=0, offsetX=0, offsetY=0, element_type="tri", mesh_type="coarse",
                col0=None, ncols=None, cell0=None, ncells=None):
    """Polygones.mesh on the device: (nodes (ncols*m,2) f64, conn (nel,npe) int32) torch tensors.
    col/cell ranges select a strip (ids local to the strip) for the multi-GPU partition."""
    torch = _capi.torch_cuda()
    _capi.bind_device()
    et = ETYPE[(element_type, mesh_type)]
    N, M = int(N), int(M)
    if M == 0:
        M = N
    fine = mesh_type == "fine"
    n, m = (2 * N + 1, 2 * M + 1) if fine else (N + 1, M + 1)
    col0 = 0 if col0 is None else col0
    ncols = n - col0 if ncols is None else ncols
    cell0 = 0 if cell0 is None else cell0
    ncells = N - cell0 if ncells is None else ncells
    per_cell = 2 if element_type == "tri" else 1
    nodes = torch.empty((ncols * m, 2), dtype=torch.float64, device="cuda")
    conn = torch.empty((ncells * M * per_cell, NPE[et]), dtype=torch.int32, device="cuda")
    call("psfem_mesh_rect", float(L), float(l), N, M, float(offsetX), float(offsetY), et, col0, ncols, cell0, ncells,
         ptr(nodes), ptr(conn), stream_ptr())
    return nodes, conn


# --------------------------------------------------------------------------------------------
# boundary conditions
# --------------------------------------------------------------------------------------------
def constrained_dofs_of(bc):
    """Polygones.py:992-997: bc order, duplicates kept, values ignored (any non-None = fixed to 0)."""
    out = []
    for node, disp in bc:
        if disp[0] is not None:
            out.append(2 * node)
        if disp[1] is not None:
            out.append(2 * node + 1)
    return out


class FreeMap:
    """free/constrained DOF bookkeeping of Polygones.py:1000 / :901 / :916 in O(n) on the device."""

    def __init__(self, n_total: int, constrained):
        torch = _capi.torch_cuda()
        _capi.bind_device()
        self.n_total = int(n_total)
        cd = np.asarray(list(constrained), dtype=np.int64)
        # DOFs beyond int32 can never match a row index; drop them before the int32 cast
        cd = cd[(cd >= 0) & (cd < self.n_total)].astype(np.int32)
        self.constrained_dev = torch.from_numpy(cd).cuda() if cd.size else None
        self.new_id = torch.empty(self.n_total, dtype=torch.int32, device="cuda")
        self.free_dofs = torch.empty(self.n_total, dtype=torch.int32, device="cuda")
        sz = C.c_size_t(0)
        call("psfem_scan_workspace", self.n_total, C.byref(sz))
        self._ws = workspace(sz.value)
        self._ws_bytes = sz.value
        nf = C.c_int64(0)
        call("psfem_free_map", self.n_total, ptr(self.constrained_dev), int(cd.size), ptr(self.new_id), ptr(self.free_dofs),
             C.byref(nf), ptr(self._ws), sz.value, stream_ptr())
        self.n_free = nf.value
        self.free_dofs = self.free_dofs[: self.n_free]
        self._keep = {}

    def reduce_csr(self, A: DeviceCSR) -> DeviceCSR:
        """K[np.ix_(free, free)] (Polygones.py:1003 / :907) on CSR."""
        torch = _capi.torch_cuda()
        assert A.n_rows == self.n_total and A.n_cols == self.n_total
        key = (A.row_ptr.data_ptr(), A.col_idx.data_ptr())
        if key not in self._keep:
            row_ptr_red = torch.empty(self.n_free + 1, dtype=torch.int32, device="cuda")
            nnz = C.c_int64(0)
            call("psfem_reduce_count", self.n_total, ptr(A.row_ptr), ptr(A.col_idx), ptr(self.new_id), self.n_free,
                 ptr(row_ptr_red), C.byref(nnz), ptr(self._ws), self._ws_bytes, stream_ptr())
            col_red = torch.empty(max(nnz.value, 1), dtype=torch.int32, device="cuda")[: nnz.value]
            keep = torch.empty(max(nnz.value, 1), dtype=torch.int32, device="cuda")[: nnz.value]
            call("psfem_reduce_fill", self.n_total, ptr(A.row_ptr), ptr(A.col_idx), ptr(self.new_id), ptr(row_ptr_red),
                 ptr(col_red), ptr(keep), stream_ptr())
            proto = DeviceCSR(self.n_free, self.n_free, row_ptr_red, col_red, None, nnz.value)
            self._keep[key] = (proto, keep, A.row_ptr, A.col_idx)    # keep refs alive: key is a pointer pair
        proto, keep, _, _ = self._keep[key]
        vals = torch.empty(max(proto.nnz, 1), dtype=torch.float64, device="cuda")[: proto.nnz]
        call("psfem_gather", ptr(A.vals), ptr(keep), proto.nnz, ptr(vals), stream_ptr())
        out = DeviceCSR(self.n_free, self.n_free, proto.row_ptr, proto.col_idx, vals, proto.nnz)
        if proto._plan is None:
            proto.vals = vals
            proto.plan()
        out._plan = proto._plan
        return out

    def reduce_vec(self, v_dev):
        """F[free_dofs] (Polygones.py:1004 / :905)."""
        torch = _capi.torch_cuda()
        out = torch.empty(self.n_free, dtype=torch.float64, device="cuda")
        call("psfem_gather", ptr(v_dev), ptr(self.free_dofs), self.n_free, ptr(out), stream_ptr())
        return out

    def expand(self, red_dev, nvec=1):
        """reconstruct_full_vector(s), Polygones.py:913-936: zeros at constrained DOFs."""
        torch = _capi.torch_cuda()
        full = torch.empty((nvec, self.n_total) if nvec > 1 else self.n_total, dtype=torch.float64, device="cuda")
        call("psfem_scatter_free", ptr(red_dev), ptr(self.free_dofs), self.n_free, self.n_total, nvec, ptr(full), stream_ptr())
        return full


# --------------------------------------------------------------------------------------------
# solvers
# --------------------------------------------------------------------------------------------
SL_MIN_ROWS = 16384       # systems up to this size run in the single-CTA solver


class LatticeOperand:
    """A K (or M) of a Q4 lattice that exists ONLY as the symmetric-lattice operand of the solvers (psfem_assemble_lattice_sl):
    19 doubles per node, no CSR values, no column indices.  Quacks like a DeviceCSR for the fused PCG and the SpMV."""

    def __init__(self, n_rows, m, roff_nodes, x_nodes, sl_vals, keep=None):
        self.n_rows, self.n_cols, self.nnz = int(n_rows), 2 * int(x_nodes), 0
        self.vals = None
        self._sl = dict(vals=sl_vals, src=None, m=int(m), roff=int(roff_nodes), x_nodes=int(x_nodes), tol=0.0)
        self._keep = keep

    def plan(self):
        return (None, 0, 0, None)

    def sym_lattice(self, row_node_offset=0, tol=None):
        return True

    def struct(self):
        s = _capi.psfem_csr()
        s.n_rows, s.nnz = self.n_rows, 0
        sl = self._sl
        s.sl_vals, s.sl_src_vals = sl["vals"].data_ptr(), None          # vals is NULL as well: "storage of exactly these values"
        s.sl_m, s.sl_roff, s.sl_x_nodes = sl["m"], sl["roff"], sl["x_nodes"]
        return s

    def matvec(self, x, alpha=1.0, gamma=0.0, z=None, out=None):
        torch = _capi.torch_cuda()
        _capi.bind_device()
        if out is None:
            out = torch.empty(self.n_rows, dtype=torch.float64, device="cuda")
        s = self.struct()
        call("psfem_spmv", C.byref(s), ptr(x), float(alpha), float(gamma), ptr(z), ptr(out), stream_ptr())
        return out


def assemble_lattice_sl(nodes_dev, n, m, c_lo, S, E=0.0, nu=0.0, rho=0.0, t=1.0, what=1):
    """Q4 lattice (n node columns x m rows, coordinates nodes_dev) assembled straight into symmetric-lattice storage for the
    node columns [c_lo, c_lo + S) (psfem_assemble_lattice_sl).  Returns the storage tensor (padded for the window copies of
    the fused PCG kernel)."""
    torch = _capi.torch_cuda()
    _capi.bind_device()
    mpad = (int(m) + 3) & ~3
    nd = int(S) * 19 * mpad
    whole = torch.empty(nd + 32 + 256, dtype=torch.float64, device="cuda")
    buf = whole[32:]
    ws = workspace(256)
    bad = C.c_int64(-1)
    rc = lib.psfem_assemble_lattice_sl(ptr(nodes_dev), int(n), int(m), _capi.Q4, float(E), float(nu), float(rho), float(t), int(what),
                                       int(c_lo), int(S), ptr(buf), nd * 8, ptr(ws), 256, C.byref(bad), stream_ptr())
    check(rc, bad.value)
    return buf


def fused_cg_enabled():
    """PSFEM_PCG=classic selects the three-kernel recurrence everywhere (A/B measurements, bit-comparison with the oracle's PCG)."""
    return os.environ.get("PSFEM_PCG", "fused") != "classic"


class FusedCG:
    """Jacobi-PCG with one kernel per iteration on an operand that carries its symmetric-lattice storage
    (psfem_cg_sl_*: Chronopoulos-Gear recurrence fused into the marching-warp SpMV).  `new_id` (FreeMap.new_id) keeps
    constrained DOFs in the matrix as masked rows."""

    def __init__(self, A: "DeviceCSR", new_id=None):
        _capi.bind_device()
        self.A, self.new_id = A, new_id
        self.n = A.n_rows
        sz = C.c_size_t(0)
        call("psfem_cg_sl_workspace", self.n, C.byref(sz))
        self.ws, self.ws_bytes = workspace(sz.value), sz.value
        self.launched = 0
        self.x = None

    @staticmethod
    def usable(A: "DeviceCSR") -> bool:
        if not fused_cg_enabled() or A._sl is None:
            return False
        ok = C.c_int(0)
        s = A.struct()
        call("psfem_cg_sl_usable", C.byref(s), C.byref(ok))
        return bool(ok.value)

    def start(self, b, x, rtol=1e-8):
        """x: start vector on entry (updated in place by the iterations)."""
        self.x, self.launched = x, 0
        s = self.A.struct()
        call("psfem_cg_sl_start", C.byref(s), ptr(b), ptr(x), ptr(self.new_id), float(rtol), ptr(self.ws), self.ws_bytes, stream_ptr())

    def iterations(self, k):
        s = self.A.struct()
        call("psfem_cg_sl_iterations", C.byref(s), ptr(self.x), 1 if self.new_id is not None else 0, int(k), self.launched,
             ptr(self.ws), self.ws_bytes, stream_ptr())
        self.launched += int(k)

    def finish(self):
        """x is written every other iteration only; adds the pending half of the update (call before reading x)."""
        if self.x is not None:
            call("psfem_cg_sl_finish", ptr(self.x), self.n, ptr(self.ws), stream_ptr())

    def state(self):
        h = (C.c_double * 5)()
        call("psfem_cg_sl_state", ptr(self.ws), self.n, h, stream_ptr())
        return dict(iterations=int(h[0]), relres=float(h[1]), bnorm=float(h[2]), done=bool(h[3]), status=int(h[4]))

    def solve(self, b, x, rtol, maxit, check_every, polish=None):
        """Solve, then evaluate the TRUE residual b - A x and continue from x while it is above the tolerance and still
        improving (psfem_cg_sl_solve_checked; at most `polish` cycles, default PSFEM_PCG_POLISH or 3; negative: no check).
        info["relres"] is the recurrence residual at convergence (what scipy's cg tests as well), info["relres_true"] the
        last evaluated ||b - A x|| / ||b||."""
        if polish is None:
            polish = int(os.environ.get("PSFEM_PCG_POLISH", "3"))
        s = self.A.struct()
        info = (C.c_double * 8)()
        rc = lib.psfem_cg_sl_solve_checked(C.byref(s), ptr(b), ptr(x), ptr(self.new_id), float(rtol), int(maxit), int(check_every),
                                           int(polish), info, ptr(self.ws), self.ws_bytes, stream_ptr())
        out = dict(iterations=int(info[0]), relres=float(info[1]), bnorm=float(info[2]), converged=(rc == _capi.OK), pcg="fused")
        if info[3] >= 0.0:
            out.update(relres_true=float(info[3]), polish_cycles=int(info[4]), polish_iterations=int(info[5]))
        return rc, out


def continue_from_true_residual(run, residual, rtol, out, polish=None):
    """The classic (three-kernel) PCG paths after convergence, like psfem_cg_sl_solve_checked for the fused one: evaluate
    ||b - A x|| (residual() -> relative norm), and while it is above rtol and improves by a factor 2 per cycle, run the
    solver again FROM x (run() -> (rc, iterations); the C entry points take x0 = d_x).  At most `polish` cycles
    (PSFEM_PCG_POLISH, default 3; negative: no evaluation).  Adds relres_true / polish_cycles / polish_iterations to `out`."""
    if polish is None:
        polish = int(os.environ.get("PSFEM_PCG_POLISH", "3"))
    if polish < 0 or not out.get("converged", False):
        return out
    cycles, extra, prev = 0, 0, None
    while True:
        true = residual()
        if true <= rtol or cycles >= polish or (cycles > 0 and not true < 0.5 * prev):
            break
        prev = true
        cycles += 1
        rc, its = run()
        extra += its
        if rc != _capi.OK:
            break
    out["iterations"] += extra
    out.update(relres_true=true, polish_cycles=cycles, polish_iterations=extra)
    return out


def pcg_jacobi(A: DeviceCSR, b, x0=None, rtol=1e-8, maxit=None, check_every=50, raise_on_fail=True):
    """Jacobi-PCG on the device.  Returns (x, info) with info = dict(iterations, relres, bnorm)."""
    torch = _capi.torch_cuda()
    _capi.bind_device()
    n = A.n_rows
    x = torch.zeros(n, dtype=torch.float64, device="cuda") if x0 is None else x0.clone()
    if maxit is None:
        maxit = 10 * n + 100
    if n > SL_MIN_ROWS:
        A.sym_lattice()           # half the SpMV traffic when A is a symmetric lattice matrix (same SpMV results bit for bit)
        if FusedCG.usable(A):     # ... and then the whole iteration is ONE kernel
            rc, out = FusedCG(A).solve(b, x, rtol, maxit, check_every)
            if rc == _capi.ENOCONV and not raise_on_fail:
                return x, out
            check(rc)
            return x, out
    s = A.struct()
    sz = C.c_size_t(0)
    call("psfem_pcg_workspace", n, s.n_blocks, C.byref(sz))
    ws = workspace(sz.value)
    info = (C.c_double * 4)()
    rc = lib.psfem_pcg_jacobi(C.byref(s), ptr(b), ptr(x), float(rtol), int(maxit), int(check_every), info, ptr(ws), sz.value,
                              stream_ptr())
    out = dict(iterations=int(info[0]), relres=float(info[1]), bnorm=float(info[2]), converged=(rc == _capi.OK), pcg="classic")
    if rc == _capi.ENOCONV and not raise_on_fail:
        return x, out
    check(rc)
    if n > SL_MIN_ROWS and out["bnorm"] > 0:      # the multi-kernel path (the single-CTA kernel of small systems runs too few iterations to drift)

        def again():
            h = (C.c_double * 4)()
            rc2 = lib.psfem_pcg_jacobi(C.byref(s), ptr(b), ptr(x), float(rtol), int(maxit), int(check_every), h, ptr(ws), sz.value,
                                       stream_ptr())
            return rc2, int(h[0])
        continue_from_true_residual(again, lambda: float(torch.linalg.vector_norm(A.matvec(x, alpha=-1.0, gamma=1.0, z=b)).item())
                                    / out["bnorm"], rtol, out)
    return x, out


def direct_limits():
    """(max rows, max half bandwidth) of the banded direct solver (psfem_direct_limits)."""
    a, b = C.c_int64(0), C.c_int64(0)
    call("psfem_direct_limits", C.byref(a), C.byref(b))
    return a.value, b.value


def half_bandwidth(A: DeviceCSR) -> int:
    """max |row - col| over the pattern (psfem_band_width)."""
    _capi.bind_device()
    hb = C.c_int64(0)
    ws = workspace(256)
    s = A.struct()
    call("psfem_band_width", C.byref(s), C.byref(hb), ptr(ws), 256, stream_ptr())
    return int(hb.value)


def direct_solve(A: DeviceCSR, b, refine=2, hb=None):
    """Banded Cholesky + `refine` rounds of iterative refinement on a double-double residual (psfem_direct_solve): the
    device counterpart of the dense LU the reference's static scripts call (Statictest.py:69, Beamexemple.py:63).
    Returns (x, info)."""
    torch = _capi.torch_cuda()
    _capi.bind_device()
    n = A.n_rows
    if hb is None:
        hb = half_bandwidth(A)
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    sz = C.c_size_t(0)
    call("psfem_direct_workspace", n, hb, C.byref(sz))
    ws = workspace(sz.value)
    info = (C.c_double * 5)()
    s = A.struct()
    rc = lib.psfem_direct_solve(C.byref(s), ptr(b), ptr(x), hb, int(refine), info, ptr(ws), sz.value, stream_ptr())
    if rc == _capi.ENOCONV:
        raise np.linalg.LinAlgError(_capi.last_error())        # what np.linalg.solve / cholesky raise on such a matrix
    check(rc)
    return x, dict(method="direct", iterations=0, refinement_rounds=int(info[0]), relres=float(info[1]), bnorm=float(info[2]),
                   last_correction=float(info[4]), half_bandwidth=int(hb), converged=True)


def residual_dd(A: DeviceCSR, x, b, out=None):
    """r = b - A x accumulated in double-double (psfem_residual_dd)."""
    torch = _capi.torch_cuda()
    _capi.bind_device()
    if out is None:
        out = torch.empty(A.n_rows, dtype=torch.float64, device="cuda")
    s = A.struct()
    call("psfem_residual_dd", C.byref(s), ptr(x), ptr(b), ptr(out), stream_ptr())
    return out


class Blas:
    """Small device BLAS-1 set for the Lanczos driver."""

    def __init__(self):
        sz = C.c_size_t(0)
        call("psfem_dot_workspace", 0, C.byref(sz))
        self._ws = workspace(sz.value)
        self._bytes = sz.value

    def dot(self, x, y):
        out = C.c_double(0)
        call("psfem_dot", x.numel(), ptr(x), ptr(y), C.byref(out), ptr(self._ws), self._bytes, stream_ptr())
        return out.value

    @staticmethod
    def axpby(a, x, b, y, out):
        call("psfem_axpby", x.numel(), float(a), ptr(x), float(b), ptr(y), ptr(out), stream_ptr())
        return out

    @staticmethod
    def multi_dot(V, m, w, out):
        call("psfem_multi_dot", m, w.numel(), ptr(V), ptr(w), ptr(out), stream_ptr())
        return out

    @staticmethod
    def multi_axpy(V, m, c, w):
        call("psfem_multi_axpy", m, w.numel(), ptr(V), ptr(c), ptr(w), stream_ptr())
        return w
