"""Strip partition of a structured plate across GPUs (one process per GPU, torch.distributed).

Node id = i*m + j with i the x index (Polygones.py:77), so contiguous node / DOF ranges are
x-strips.  Rank r owns the node columns [c0, c1); it meshes every cell column touching an owned node
(its own cells plus its left neighbour's last cell column) and all their node columns (a halo of one
cell column on the left, one node column on the right), assembles that extended strip with
no communication (the interface cell columns are computed on both sides, 1/4096 of the work) and
keeps the rows of its owned DOFs, which are complete and bit-identical to the same rows of the
single-GPU matrix.  Per PCG iteration the only exchanges are
  * the halo: the free DOFs of my first / last owned column go to my left / right neighbour
    (2*(M+1) doubles each, grouped NCCL send/recv straight out of / into the extended p vector),
  * two all-reduces of scalars (p.q, then r.z and r.r together).
All scalars stay on the device; the host never waits inside the iteration.

`StripPartition` is pure index arithmetic (tested on CPU); `pcg_iterations` is the rank-local
iteration loop written against a small `ops` interface so that the CPU/gloo tests drive exactly the
same control flow with NumPy ops; `StripProblem` binds it to the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _capi
from ._capi import call, ptr, stream_ptr


class StripPartition:
    """Which node columns / cell columns / DOFs a rank owns and meshes (coarse or fine meshes).

    Cell columns are split evenly; rank r owns the node columns from the left edge of its first cell
    column up to (not including) the left edge of the next rank's first cell column, the last rank
    also owns the right edge.  It meshes every cell column touching an owned node column, i.e. its
    own cells plus the last cell column of its left neighbour, and all node columns of those cells:
    a left halo of `step` node columns (1 coarse, 2 fine) and a right halo of one node column."""

    def __init__(self, N: int, M: int, world: int, rank: int, fine: bool = False):
        assert 0 <= rank < world
        self.N, self.M, self.world, self.rank, self.fine = N, M, world, rank, fine
        self.n = (2 * N + 1) if fine else (N + 1)           # node columns (Polygones.py:66-71)
        self.m = (2 * M + 1) if fine else (M + 1)           # nodes per column
        self.step = step = 2 if fine else 1                 # node columns per cell column
        base, rem = divmod(N, world)
        assert base >= 1, "more ranks than cell columns"
        cs = [r * base + min(r, rem) for r in range(world + 1)]
        self.cell_start = cs
        self.c0 = step * cs[rank]                           # owned node columns [c0, c1)
        self.c1 = step * cs[rank + 1] if rank < world - 1 else self.n
        self.cell0, self.cell1 = max(cs[rank] - 1, 0), cs[rank + 1]      # meshed cell columns
        self.e0, self.e1 = step * self.cell0, step * self.cell1 + 1       # meshed node columns
        self.left_halo_cols = self.c0 - self.e0             # 0 or step
        self.right_halo_cols = self.e1 - self.c1            # 0 or 1
        self.n_ext_nodes = (self.e1 - self.e0) * self.m
        self.n_owned_nodes = (self.c1 - self.c0) * self.m
        self.dof_shift = 2 * self.m * self.e0               # global dof = extended-local dof + shift
        self.own_lo = 2 * self.m * (self.c0 - self.e0)      # owned DOF range in extended-local numbering
        self.own_hi = 2 * self.m * (self.c1 - self.e0)
        self.n_ext_dof = 2 * self.n_ext_nodes

    @property
    def left(self):
        return self.rank - 1 if self.rank > 0 else None

    @property
    def right(self):
        return self.rank + 1 if self.rank < self.world - 1 else None

    def local_constrained(self, constrained_global):
        """global constrained DOF list -> extended-local ids (those outside the strip dropped)."""
        cd = np.asarray(list(constrained_global), dtype=np.int64) - self.dof_shift
        return cd[(cd >= 0) & (cd < self.n_ext_dof)]

    def halo_plan(self, free_mask_ext=None, count_free=None):
        """Index ranges, in the REDUCED (free-DOF) extended vector, of the halo exchange.
        `count_free(a, b)` = number of free extended-local DOFs in [a, b) (or pass the boolean mask).
        own = (lo, hi): owned entries; recv_left = [0, lo), recv_right = [hi, n_ext_free);
        send_left = my first node column (my left neighbour's right halo), send_right = my last `step`
        node columns (my right neighbour's left halo).  Both sides count the same global DOFs, so the
        message sizes agree without negotiation."""
        if count_free is None:
            csum = np.concatenate([[0], np.cumsum(np.asarray(free_mask_ext).astype(np.int64))])
            count_free = lambda a, b: int(csum[b] - csum[a])
        col = 2 * self.m
        lo = count_free(0, self.own_lo)
        hi = lo + count_free(self.own_lo, self.own_hi)
        n_ext_free = hi + count_free(self.own_hi, self.n_ext_dof)
        plan = dict(own=(lo, hi), n_ext_free=n_ext_free, send_left=None, send_right=None, recv_left=None, recv_right=None)
        if self.left is not None:
            plan["recv_left"] = (0, lo)
            plan["send_left"] = (lo, lo + count_free(self.own_lo, self.own_lo + col))
        if self.right is not None:
            plan["recv_right"] = (hi, n_ext_free)
            plan["send_right"] = (hi - count_free(self.own_hi - self.step * col, self.own_hi), hi)
        return plan


def constrains_whole_nodes(constrained_dofs) -> bool:
    """True when every constrained node has BOTH of its DOFs fixed (clamped edges): the reduced matrix of
    Polygones.py:1003 then keeps the 2x2 node-block structure.  Rollers and other single-DOF conditions return False:
    the strip solver keeps those DOFs in the matrix as masked rows instead (StripProblem.setup)."""
    cd = np.unique(np.asarray(list(constrained_dofs), dtype=np.int64))
    nodes = np.unique(cd >> 1)
    return cd.size == 2 * nodes.size


# ------------------------------------------------------------------------------------------------
# rank-local PCG iteration loop (backend-agnostic)
# ------------------------------------------------------------------------------------------------
class Comm:
    """torch.distributed plumbing of the strip solver: halo send/recv + scalar all-reduce."""

    def __init__(self, part: StripPartition, plan: dict, dist_module=None):
        self.part, self.plan, self.dist = part, plan, dist_module
        self.active = part.world > 1

    def halo(self, p_ext):
        if not self.active:
            return
        d = self.dist
        ops = []
        pl = self.plan
        if self.part.left is not None:
            a, b = pl["send_left"]
            ops.append(d.P2POp(d.isend, p_ext[a:b], self.part.left))
            a, b = pl["recv_left"]
            ops.append(d.P2POp(d.irecv, p_ext[a:b], self.part.left))
        if self.part.right is not None:
            a, b = pl["send_right"]
            ops.append(d.P2POp(d.isend, p_ext[a:b], self.part.right))
            a, b = pl["recv_right"]
            ops.append(d.P2POp(d.irecv, p_ext[a:b], self.part.right))
        for w in d.batch_isend_irecv(ops):
            w.wait()

    def allreduce(self, t):
        if self.active:
            self.dist.all_reduce(t)


def pcg_iterations(ops, comm: Comm, st: dict, k: int):
    """k Jacobi-PCG iterations on the strip-partitioned system.  `st` holds x, r, q, dinv (owned rows),
    p_ext (extended, halo up to date on entry), scalar tensors rz (2 slots), pq, out2, and `parity`.
    Identical recurrence to the single-GPU solver; only the two reductions and the halo are global."""
    lo, hi = comm.plan["own"]
    for _ in range(k):
        cur, nxt = st["parity"], st["parity"] ^ 1
        ops.spmv_dot(st["p_ext"], lo, st["q"], st["pq"])                 # q = A p ; pq = p_own . q (local)
        comm.allreduce(st["pq"])
        ops.update(st["r"], st["q"], st["dinv"], st["rz"][cur:cur + 1], st["pq"], st["out2"])   # r -= alpha q
        comm.allreduce(st["out2"])                                        # {r.z, r.r}
        ops.copy_scalar(st["out2"][0:1], st["rz"][nxt:nxt + 1])
        # x += alpha p (old p), p = dinv r + beta p
        ops.pupdate(st["p_ext"][lo:hi], st["x"], st["r"], st["dinv"], st["rz"][cur:cur + 1], st["rz"][nxt:nxt + 1], st["pq"])
        comm.halo(st["p_ext"])
        st["parity"] = nxt
        st["iters"] += 1


def pcg_start(ops, comm: Comm, st: dict, b_own):
    """x = 0, r = b, p = dinv r (+ halo), rz = r.z, rr, bb."""
    lo, hi = comm.plan["own"]
    ops.init(st["r"], st["dinv"], b_own, st["p_ext"][lo:hi], st["out3"])   # r already holds b
    comm.allreduce(st["out3"])
    ops.copy_scalar(st["out3"][0:1], st["rz"][0:1])
    comm.halo(st["p_ext"])
    st["parity"] = 0
    st["iters"] = 0


# ------------------------------------------------------------------------------------------------
# peer memory (CUDA IPC over NVLink / NVSwitch) for the fused strip solver
# ------------------------------------------------------------------------------------------------
class _RawDeviceArray:
    """__cuda_array_interface__ view of library-owned device memory (torch.as_tensor aliases it)."""

    def __init__(self, ptr_, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr_), False), "version": 2}


class PeerFabric:
    """This rank's board and extended p vector live in memory that the other ranks map (psfem_peer_alloc /
    psfem_peer_open); the PCG kernels then exchange their scalars and halo columns with plain stores over
    NVLink instead of NCCL calls (include/psfem.h, psfem_dist_*).  The host only swaps the 64-byte IPC
    handles and the halo sizes once, through torch.distributed."""

    def __init__(self, part: StripPartition, plan: dict, dist_module):
        """Collective: every rank must call it.  Whether the peer transport is usable is decided TOGETHER -- a failure of
        cudaIpcGetMemHandle / cudaIpcOpenMemHandle on any rank (no P2P between one pair of GPUs) is all-reduced, every
        rank then releases what it had allocated or mapped and raises, so that all ranks fall back to NCCL in step
        (a rank that skipped the collectives below on its own would leave the others hanging in them)."""
        torch = _capi.torch_cuda()
        self.part, self.plan, self.dist = part, plan, dist_module
        self.opened = []
        self.own = []
        self.p_ext = None
        world, rank = part.world, part.rank
        if world > _capi.MAX_RANKS:
            raise RuntimeError(f"fused strip solver supports up to {_capi.MAX_RANKS} ranks")
        self.n_ext_free = plan["n_ext_free"]
        lo, hi = plan["own"]

        def agree(err):
            """all-reduce(MIN) of 'my step worked'; on any failure release everything on every rank and raise"""
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device="cuda")
            dist_module.all_reduce(ok, op=dist_module.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self._release_local()
                raise RuntimeError(f"peer-memory transport unavailable on at least one rank ({err or 'a peer rank failed'})")

        # stage 1 (local): my board and my extended p vector
        err, hb, hp = None, None, None
        try:
            sz = C.c_size_t(0)
            call("psfem_peer_board_bytes", C.byref(sz))
            self.board_ptr, hb = self._alloc(sz.value)
            self.p_ptr, hp = self._alloc(max(self.n_ext_free, 1) * 8)
            call("psfem_dist_cg_halo_bytes", part.m, C.byref(sz))      # LL halo buffer of the fused one-kernel iteration
            self.ll_ptr, hl = self._alloc(sz.value)
        except Exception as e:                                # noqa: BLE001 -- any failure means "no peer transport"
            err = str(e)
            hl = None
        agree(err)
        # stage 2 (collective): swap the IPC handles and the halo sizes
        mine = dict(board=hb, p=hp, ll=hl, lo=lo, hi=hi, n_ext_free=self.n_ext_free,
                    send_left=plan["send_left"], send_right=plan["send_right"])
        everyone = [None] * world
        dist_module.all_gather_object(everyone, mine)
        # stage 3 (local): map the other ranks' boards and my neighbours' vectors
        err = None
        try:
            self.boards = [None] * world
            for w in range(world):
                self.boards[w] = self.board_ptr if w == rank else self._open(everyone[w]["board"])
            self.left_p = self.right_p = None
            self.left_ll = self.right_ll = None
            self.left_dst_off = self.right_dst_off = 0
            self.send_left_n = self.send_right_n = 0
            if part.left is not None:
                nb = everyone[part.left]
                self.left_p = self._open(nb["p"])
                self.left_ll = self._open(nb["ll"])
                self.left_dst_off = nb["hi"]                                   # its recv_right region [hi, n_ext_free)
                self.send_left_n = plan["send_left"][1] - plan["send_left"][0]
                if self.send_left_n != nb["n_ext_free"] - nb["hi"] or plan["send_left"][0] != lo:
                    raise RuntimeError("halo plan mismatch with the left neighbour")
            if part.right is not None:
                nb = everyone[part.right]
                self.right_p = self._open(nb["p"])
                self.right_ll = self._open(nb["ll"])
                self.right_dst_off = 0                                           # its recv_left region [0, lo)
                self.send_right_n = plan["send_right"][1] - plan["send_right"][0]
                if self.send_right_n != nb["lo"] or plan["send_right"][1] != hi:
                    raise RuntimeError("halo plan mismatch with the right neighbour")
        except Exception as e:                                # noqa: BLE001
            err = str(e)
        agree(err)
        self.p_ext = torch.as_tensor(_RawDeviceArray(self.p_ptr, self.n_ext_free), device="cuda")
        dist_module.barrier()

    def _release_local(self):
        """Unmap and free without any collective (the failure path of __init__: every rank runs it after the all-reduce)."""
        self.p_ext = None
        for q in self.opened:
            try:
                call("psfem_peer_close", C.c_void_p(q))
            except Exception:                                  # noqa: BLE001
                pass
        self.opened = []
        for q in self.own:
            try:
                call("psfem_peer_free", C.c_void_p(q))
            except Exception:                                  # noqa: BLE001
                pass
        self.own = []

    def _alloc(self, nbytes):
        p = C.c_void_p(0)
        h = (C.c_ubyte * 64)()
        call("psfem_peer_alloc", int(nbytes), C.byref(p), h)
        self.own.append(p.value)
        return p.value, bytes(h)

    def _open(self, handle):
        p = C.c_void_p(0)
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        call("psfem_peer_open", h, C.byref(p))
        self.opened.append(p.value)
        return p.value

    def close(self):
        torch = _capi.torch_cuda()
        torch.cuda.synchronize()
        self.dist.barrier()                       # nobody still stores into / reads from a mapping
        self.p_ext = None
        for q in self.opened:
            call("psfem_peer_close", C.c_void_p(q))
        self.opened = []
        self.dist.barrier()
        for q in self.own:
            call("psfem_peer_free", C.c_void_p(q))
        self.own = []


# ------------------------------------------------------------------------------------------------
# CUDA backend
# ------------------------------------------------------------------------------------------------
class CudaOps:
    def __init__(self, A_struct, n_own, n_blocks, masked=False):
        torch = _capi.torch_cuda()
        self.A = A_struct
        self.n = n_own
        self._update = "psfem_pcg_update_masked" if masked else "psfem_pcg_update"
        npart = max(n_blocks, 148 * 8) * 3
        self.partials = torch.empty(npart, dtype=torch.float64, device="cuda")
        self.counter = torch.zeros(4, dtype=torch.int32, device="cuda")

    def spmv_dot(self, p_ext, x_offset, q, pq):
        call("psfem_pcg_spmv_dot", C.byref(self.A), ptr(p_ext), x_offset, ptr(q), ptr(pq), ptr(self.partials),
             ptr(self.counter), stream_ptr())

    def update(self, r, q, dinv, rz, pq, out2):
        call(self._update, self.n, ptr(r), ptr(q), ptr(dinv), ptr(rz), ptr(pq), ptr(out2),
             ptr(self.partials), C.c_void_p(self.counter.data_ptr() + 4), stream_ptr())

    def pupdate(self, p_own, x, r, dinv, rz_old, rz_new, pq):
        call("psfem_pcg_pupdate", self.n, ptr(p_own), ptr(x), ptr(r), ptr(dinv), ptr(rz_old), ptr(rz_new), ptr(pq), stream_ptr())

    def init(self, r, dinv, b, p_own, out3):
        call("psfem_pcg_init", self.n, ptr(r), ptr(dinv), ptr(b), ptr(p_own), ptr(out3), ptr(self.partials),
             C.c_void_p(self.counter.data_ptr() + 8), stream_ptr())

    @staticmethod
    def copy_scalar(src, dst):
        call("psfem_axpby", 1, 1.0, ptr(src), 0.0, ptr(None), ptr(dst), stream_ptr())


class StripProblem:
    """Clamped-left / loaded-right Q4 plate of BASELINE configs[3]/[4] on `world` GPUs: strip mesh,
    symbolic pattern, numeric assembly, BC reduction, distributed Jacobi-PCG.  Also the single-GPU
    benchmark object (world = 1)."""

    def __init__(self, L, l, N, M, rank=0, world=1, E=210e9, nu=0.3, t=0.1, element_type="quad", mesh_type="coarse",
                 load=(0.0, -1e6), transport=None, bc=None, F=None):
        """bc: boundary conditions like the reference's lists [(global node, [ux, uy]), ...] (None entries = free
        component); default Polygones.boundary('left', [0, 0], N, M).  F: global load, either a callable
        (dof_lo, dof_hi) -> ndarray of that DOF range (so no rank builds the whole vector) or a full-length array;
        default Polygones.edgeForces(zeros, load, 'right', l, N, M)."""
        import os
        self.bc, self.F = bc, F
        # "peer": collectives fused into the PCG kernels over CUDA-IPC peer memory (default for world > 1);
        # "nccl": torch.distributed all-reduce + send/recv between the kernels
        self.transport = transport or os.environ.get("PSFEM_DIST_TRANSPORT", "peer")
        self.fabric = None
        self.seq = 0
        self.L, self.l, self.N, self.M = L, l, N, M
        self.rank, self.world = rank, world
        self.E, self.nu, self.t = E, nu, t
        self.element_type, self.mesh_type, self.load = element_type, mesh_type, load
        self.part = StripPartition(N, M, world, rank, fine=(mesh_type == "fine"))

    # -- set-up ---------------------------------------------------------------------------------
    def setup(self, nodes_conn=None, lattice_only=False):
        """lattice_only (Q4 meshes): K is assembled straight into the symmetric-lattice operand of the fused PCG
        (psfem_assemble_lattice_sl) -- no CSR values, column indices, reduction copy or SpMV plans exist; solve-only flows."""
        from .device import Assembler, FreeMap, device_mesh
        from ._capi import ETYPE
        torch = _capi.torch_cuda()
        _capi.bind_device()
        p = self.part
        self.lattice_only = bool(lattice_only)
        if lattice_only:
            return self._setup_lattice_only(nodes_conn)
        if nodes_conn is None:
            self.nodes, self.conn = device_mesh(self.L, self.l, self.N, self.M, 0, 0, self.element_type, self.mesh_type,
                                                col0=p.e0, ncols=p.e1 - p.e0, cell0=p.cell0, ncells=p.cell1 - p.cell0)
        else:
            self.nodes, self.conn = nodes_conn
        et = ETYPE[(self.element_type, self.mesh_type)]
        self.asm = Assembler(self.conn, p.n_ext_nodes, et)
        f64 = dict(dtype=torch.float64, device="cuda")
        self.K_vals = torch.empty(self.asm.nnz, **f64)
        self.asm.assemble(self.nodes, E=self.E, nu=self.nu, t=self.t, what=1, K_out=self.K_vals)
        K_ext = self.asm.pattern_csr(self.K_vals)
        # boundary conditions: left edge clamped (global node column 0), Polygones.boundary('left',[0,0],...)
        m = p.m
        if self.bc is None:
            cd_global = np.arange(2 * m, dtype=np.int64)          # DOFs 2k, 2k+1 of nodes k = 0..m-1
        else:
            from .device import constrained_dofs_of
            cd_global = np.asarray(constrained_dofs_of(self.bc), dtype=np.int64)
        self.constrained_global = cd_global
        cd_local = p.local_constrained(cd_global)
        # Boundary conditions that fix single DOFs of a node (rollers) would take the 2x2 node-block structure away
        # from the reduced matrix (Polygones.py:1003) and with it the fast SpMV kernels: such problems are solved on the
        # UNREDUCED strip matrix with the constrained DOFs masked (D^-1 = 0, zero load and residual there) -- the
        # iterates on the free DOFs are those of the reduced system, like psfem_pcg_jacobi_masked on one GPU.
        import os
        self.masked = (not constrains_whole_nodes(cd_local)) and K_ext.plan()[3] is not None and not os.environ.get("PSFEM_DIST_NO_MASK")
        from .device import DeviceCSR
        if self.masked:
            self.fm = None
            self.K_ext_red = K_ext
            self.plan = p.halo_plan(count_free=lambda a, b: b - a)
            self._cd_owned = torch.from_numpy(np.unique(cd_local[(cd_local >= p.own_lo) & (cd_local < p.own_hi)]) - p.own_lo).cuda()
        else:
            self.fm = FreeMap(p.n_ext_dof, cd_local)
            self.K_ext_red = self.fm.reduce_csr(K_ext)
            free = self.fm.free_dofs                                   # sorted extended-local ids of free DOFs

            def count_free(a, b):
                lo_ = int(torch.searchsorted(free, torch.tensor([a], dtype=free.dtype, device="cuda")).item())
                hi_ = int(torch.searchsorted(free, torch.tensor([b], dtype=free.dtype, device="cuda")).item())
                return hi_ - lo_
            self.plan = p.halo_plan(count_free=count_free)
        lo, hi = self.plan["own"]
        self.n_owned_free = hi - lo                                  # rows of this rank's operand (masked: constrained rows included)
        self.n_ext_free = self.plan["n_ext_free"]
        # owned rows of the (reduced) extended matrix: a row slice (pointer offset), columns stay extended
        rp = self.K_ext_red.row_ptr[lo:hi + 1]
        self.K_red = DeviceCSR(hi - lo, self.n_ext_free, rp, self.K_ext_red.col_idx, self.K_ext_red.vals, nnz=None)
        self.K_red.nnz = int((rp[-1] - rp[0]).item())
        # symmetric-lattice storage of the owned rows (+ the column before them): half the SpMV traffic, same bits
        self.sym_lattice = (lo % 2 == 0) and self.K_red.sym_lattice(row_node_offset=lo // 2)
        self.A = self.K_red.struct()
        nb = self.K_red.plan()[1]
        # load: edgeForces(F,[0,-1e6],'right',l,N,M): f*l/(m-1) on every node of the right edge (global column n-1)
        if self.F is None:
            F_ext = torch.zeros(p.n_ext_dof, **f64)
            if p.e1 == p.n:
                k0 = (p.n - 1 - p.e0) * m
                F_ext[2 * k0:2 * (k0 + m):2] = self.load[0] * self.l / (m - 1)
                F_ext[2 * k0 + 1:2 * (k0 + m):2] = self.load[1] * self.l / (m - 1)
        else:
            a, b = p.dof_shift, p.dof_shift + p.n_ext_dof
            Fh = self.F(a, b) if callable(self.F) else np.asarray(self.F)[a:b]
            F_ext = torch.from_numpy(np.ascontiguousarray(Fh, dtype=np.float64)).cuda()
        if self.masked:
            self.b_own = F_ext[lo:hi].clone()
            self.b_own[self._cd_owned] = 0.0                        # a load on a constrained DOF is dropped (Polygones.py:1004)
        else:
            self.b_own = self.fm.reduce_vec(F_ext)[lo:hi].contiguous()
        # global sizes
        self.n_dof_global = 2 * p.n * p.m
        cdg = cd_global[(cd_global >= 0) & (cd_global < self.n_dof_global)]
        self.n_free_global = self.n_dof_global - len(np.unique(cdg))
        per_cell = 2 if self.element_type == "tri" else 1
        self.nel_global = self.N * self.M * per_cell
        self.nel_owned = self.asm.nel
        return self._setup_solver(lo, nb)

    def _setup_lattice_only(self, nodes_conn):
        """Solve-only strip: mesh -> K in symmetric-lattice storage -> fused PCG state.  The default boundary condition (left
        edge clamped) is the reduction of Polygones.py:1003 by dropping node column 0 from the stored columns; any other
        `bc` keeps its DOFs in the system as masked rows."""
        from .device import LatticeOperand, assemble_lattice_sl, device_mesh, fused_cg_enabled
        torch = _capi.torch_cuda()
        p = self.part
        if (self.element_type, self.mesh_type) != ("quad", "coarse"):
            raise ValueError("lattice_only: Q4 meshes only")
        if not fused_cg_enabled():
            raise RuntimeError("lattice_only needs the fused PCG (PSFEM_PCG=classic is set)")
        if nodes_conn is None:
            self.nodes, _ = device_mesh(self.L, self.l, self.N, self.M, 0, 0, self.element_type, self.mesh_type,
                                        col0=p.e0, ncols=p.e1 - p.e0, cell0=p.cell0, ncells=0)
        else:
            self.nodes = nodes_conn[0]
        self.conn = None
        f64 = dict(dtype=torch.float64, device="cuda")
        m = p.m
        if self.bc is None:
            cd_global = np.arange(2 * m, dtype=np.int64)
        else:
            from .device import constrained_dofs_of
            cd_global = np.asarray(constrained_dofs_of(self.bc), dtype=np.int64)
        self.constrained_global = cd_global
        cd_local = p.local_constrained(cd_global)
        pre = p.left_halo_cols                               # 0 or 1 node columns before the owned ones
        drop0 = 1 if (self.bc is None and p.c0 == 0) else 0  # the clamped column 0 leaves the system
        own_first = (p.c0 - p.e0) + drop0                    # strip-local index of the first owned node column
        c_lo = own_first - pre
        n_loc = p.e1 - p.e0
        own_end = p.c1 - p.e0
        S = own_end - c_lo
        n_cols_own = S - pre
        if n_cols_own < 1:
            raise ValueError("lattice_only: a rank without owned node columns")
        n = 2 * m * n_cols_own
        self.masked = self.bc is not None
        sl = assemble_lattice_sl(self.nodes, n_loc, m, c_lo, S, E=self.E, nu=self.nu, t=self.t, what=1)
        self.K_red = LatticeOperand(n, m, pre * m, (S + p.right_halo_cols) * m, sl)
        self.K_vals = self.K_ext_red = self.fm = self.asm = None
        self.sym_lattice = True
        self.A = self.K_red.struct()
        dof0, dof1 = 2 * m * own_first, 2 * m * own_end      # owned DOFs in strip-local numbering
        if self.masked:
            own_cd = np.unique(cd_local[(cd_local >= dof0) & (cd_local < dof1)]) - dof0
            self._cd_owned = torch.from_numpy(own_cd).cuda()
            count_free = lambda a, b: b - a
        else:
            c0a, c0b = (0, 2 * m) if (p.e0 == 0) else (0, 0)  # strip-local DOFs of the clamped global column 0

            def count_free(a, b):
                return (b - a) - max(0, min(b, c0b) - max(a, c0a))
        self.plan = p.halo_plan(count_free=count_free)
        lo, hi = self.plan["own"]
        assert hi - lo == n, (lo, hi, n)
        self.n_owned_free = n
        self.n_ext_free = self.plan["n_ext_free"]
        if self.F is None:
            F_ext = torch.zeros(p.n_ext_dof, **f64)
            if p.e1 == p.n:
                k0 = (p.n - 1 - p.e0) * m
                F_ext[2 * k0:2 * (k0 + m):2] = self.load[0] * self.l / (m - 1)
                F_ext[2 * k0 + 1:2 * (k0 + m):2] = self.load[1] * self.l / (m - 1)
        else:
            a, b = p.dof_shift, p.dof_shift + p.n_ext_dof
            Fh = self.F(a, b) if callable(self.F) else np.asarray(self.F)[a:b]
            F_ext = torch.from_numpy(np.ascontiguousarray(Fh, dtype=np.float64)).cuda()
        self.b_own = F_ext[dof0:dof1].clone()
        if self.masked:
            self.b_own[self._cd_owned] = 0.0
        self.n_dof_global = 2 * p.n * p.m
        cdg = cd_global[(cd_global >= 0) & (cd_global < self.n_dof_global)]
        self.n_free_global = self.n_dof_global - len(np.unique(cdg))
        self.nel_global = self.N * self.M
        self.nel_owned = (n_loc - 1) * self.M
        return self._setup_solver(lo, 0)

    def _setup_solver(self, lo, nb):
        """PCG state, peer fabric and the fused / classic iteration contexts (shared by the CSR and the lattice-only set-up)."""
        torch = _capi.torch_cuda()
        p = self.part
        f64 = dict(dtype=torch.float64, device="cuda")
        # PCG state
        self.ops = CudaOps(self.A, self.n_owned_free, nb, masked=self.masked)
        import torch.distributed as dist
        self.comm = Comm(p, self.plan, dist if self.world > 1 else None)
        n = self.n_owned_free
        if self.world > 1 and self.transport == "peer":
            try:
                self.fabric = PeerFabric(p, self.plan, dist)
            except RuntimeError as e:                   # no P2P / IPC on this box: PeerFabric decided that COLLECTIVELY and
                import warnings                         # released its memory on every rank, so all ranks switch together
                warnings.warn(f"{e}; using NCCL")
                self.fabric, self.transport = None, "nccl"
        p_ext = self.fabric.p_ext if self.fabric else torch.zeros(self.n_ext_free, **f64)
        self.st = dict(x=torch.zeros(n, **f64), r=self.b_own.clone(), q=torch.empty(n, **f64),
                       dinv=torch.empty(n, **f64), p_ext=p_ext,
                       rz=torch.zeros(2, **f64), pq=torch.zeros(1, **f64), out2=torch.zeros(2, **f64),
                       out3=torch.zeros(8, **f64), parity=0, iters=0)
        if not self.lattice_only:
            call("psfem_extract_diag", C.byref(self.A), lo, 1, ptr(self.st["dinv"]), stream_ptr())
            if self.masked:
                self.st["dinv"][self._cd_owned] = 0.0
        self._new_id = None
        if self.masked:                                   # free map of the owned rows for the fused kernels (negative = constrained)
            self._new_id = torch.zeros(n, dtype=torch.int32, device="cuda")
            self._new_id[self._cd_owned] = -1
        if self.fabric and not self.lattice_only:
            fb, st = self.fabric, self.st
            c = self.ctx = _capi.psfem_dist_ctx()
            c.rank, c.world = self.rank, self.world
            c.has_left, c.has_right = int(p.left is not None), int(p.right is not None)
            for w in range(self.world):
                c.board[w] = fb.boards[w]
            c.p_ext, c.left_p_ext, c.right_p_ext = fb.p_ptr, fb.left_p, fb.right_p
            c.own_lo, c.n_own = lo, n
            c.send_left_n, c.left_dst_off = fb.send_left_n, fb.left_dst_off
            c.send_right_n, c.right_dst_off = fb.send_right_n, fb.right_dst_off
            c.x, c.r, c.q, c.dinv = st["x"].data_ptr(), st["r"].data_ptr(), st["q"].data_ptr(), st["dinv"].data_ptr()
            c.out3 = st["out3"].data_ptr()
            c.partials, c.counter = self.ops.partials.data_ptr(), self.ops.counter.data_ptr()
            c.masked = int(self.masked)
        # strips with the peer transport and symmetric-lattice storage: the fused one-kernel iteration with the halo of u and
        # ONE scalar exchange inside it (psfem_dist_cg_*); PSFEM_PCG=classic keeps the three-kernel strip solver
        self.fused_dist = False
        if self.fabric and self.sym_lattice and not (self.masked and lo % 2):
            from .device import FusedCG, fused_cg_enabled, workspace
            ok = C.c_int(0)
            call("psfem_cg_sl_usable", C.byref(self.A), C.byref(ok))
            if fused_cg_enabled() and ok.value in (1, 2):
                sz = C.c_size_t(0)
                call("psfem_cg_sl_workspace", n, C.byref(sz))
                self.cg_ws, self.cg_ws_bytes = workspace(sz.value), sz.value
                g = self.cg_ctx = _capi.psfem_dist_cg_ctx()
                fb = self.fabric
                g.rank, g.world = self.rank, self.world
                g.has_left, g.has_right = int(p.left is not None), int(p.right is not None)
                for w_ in range(self.world):
                    g.board[w_] = fb.boards[w_]
                g.ll_mine, g.ll_left, g.ll_right = fb.ll_ptr, fb.left_ll, fb.right_ll
                g.m = p.m
                g.x = self.st["x"].data_ptr()
                g.counter = self.ops.counter.data_ptr()
                g.masked = int(self.masked)
                self.fused_dist = True
                self.launches = 0
        # one GPU, symmetric-lattice operand: the whole iteration is one kernel (psfem_cg_sl.cuh)
        self.fused = None
        if self.world == 1 and self.sym_lattice:
            from .device import FusedCG
            if FusedCG.usable(self.K_red):
                self.fused = FusedCG(self.K_red, new_id=self._new_id)
        if self.lattice_only and self.fused is None and not self.fused_dist:
            raise RuntimeError("lattice_only: the fused PCG is not available for this strip (no peer transport / operand rejected)")
        self.start()
        torch.cuda.synchronize()
        return self


    def start(self, b=None):
        """x = 0, r = b, p = D^-1 r with its halo, r.z / r.r / b.b.  (b: another right-hand side on the owned rows, fused strip
        solver only -- the correction equation of _restart_from_x.)"""
        if b is None:
            self._x_acc = None
        if self.fused is not None:
            self.st["x"].zero_()
            self.fused.start(self.b_own, self.st["x"], rtol=0.0)
            self.st["iters"] = 0
            return
        if self.fused_dist:
            torch = _capi.torch_cuda()
            if self.seq > 0:                      # restart on live boards / halo buffers: everybody has consumed the previous sequence
                import torch.distributed as dist
                torch.cuda.synchronize()
                dist.barrier()
            self.seq += 1
            call("psfem_dist_cg_start", C.byref(self.A), C.byref(self.cg_ctx), ptr(self.b_own if b is None else b), ptr(self._new_id),
                 0.0, self.seq, ptr(self.cg_ws), self.cg_ws_bytes, stream_ptr())
            self.launches = 1
            self.st["iters"] = 0
            self._bb = None
            return
        if self.fabric:
            if self.seq > 0:
                # restart on live boards: every rank must have consumed the previous sequence numbers of stage 1 before
                # a fast rank publishes the start step into the same slots
                torch = _capi.torch_cuda()
                import torch.distributed as dist
                torch.cuda.synchronize()
                dist.barrier()
            self.seq += 1
            call("psfem_dist_pcg_start", C.byref(self.ctx), ptr(self.b_own), self.seq, stream_ptr())
            self.st["iters"] = 0
            h = (C.c_double * 3)()
            call("psfem_dist_pcg_residual", C.byref(self.ctx), self.seq, h, stream_ptr())
            self._bb = float(h[2])
        else:
            self.st["x"].zero_()
            self.st["r"].copy_(self.b_own)
            pcg_start(self.ops, self.comm, self.st, self.b_own)

    # -- what the benchmark reports about the iteration ------------------------------------------------
    @property
    def pcg_variant(self):
        if getattr(self, "fused", None) is not None or getattr(self, "fused_dist", False):
            return ("fused Chronopoulos-Gear Jacobi-PCG: pointwise updates inside the marching-warp symmetric-lattice SpMV "
                    "(1 kernel, 1 reduction per iteration)")
        return "classic Jacobi-PCG: SpMV+dot | r-update+dots | x,p-update (3 kernels, 2 reductions per iteration)"

    @property
    def kernels_per_iteration(self):
        return 1 if (getattr(self, "fused", None) is not None or getattr(self, "fused_dist", False)) else 3

    def iteration_bytes(self, spmv_bytes):
        """algorithmic HBM bytes of one PCG iteration on this rank.  Classic: the SpMV + update (q, r, dinv read, r
        written) + p-update (p, x, r, dinv read, p, x written) = spmv + 80 B per DOF.  Fused: 152 B of matrix per node +
        r, w, s, p read and r, s, w, p written (128) + x read and written every other iteration (16) = 296 B per node
        (D^-1 comes from the streamed diagonal)."""
        if self.kernels_per_iteration == 1:
            return (152 + 128 + 16) * (self.n_owned_free // 2)
        return spmv_bytes + (4 + 6) * 8 * self.n_owned_free

    @staticmethod
    def dominant_kernel(sym_lat, node_block, idx_bytes):
        if sym_lat:
            return ("spmv_sl_tma_kernel<true> (SpMV + dot on symmetric-lattice storage: marching warps, matrix staged by TMA bulk copies, "
                    "152 B of matrix per node, bit-identical to the CSR kernels)")
        if node_block:
            return f"spmv_nb_kernel<true> (PCG SpMV + dot on the 2x2 node-block structure, TMA-staged, {8 + idx_bytes / 4:g} B/non-zero)"
        return "spmv2_kernel<true> (PCG SpMV + dot, CSR, TMA-staged, 12 B/non-zero)"

    # -- timed pieces -----------------------------------------------------------------------------
    def assemble(self):
        self.asm.assemble(self.nodes, E=self.E, nu=self.nu, t=self.t, what=1, K_out=self.K_vals)
        # masked strips solve on K_vals itself (no reduced copy): a re-assembly changes the operand under its
        # symmetric-lattice storage, which is rebuilt before the next iteration
        if getattr(self, "masked", False) and getattr(self, "sym_lattice", False):
            self._sl_stale = True

    def _refresh_operand(self):
        if getattr(self, "_sl_stale", False):
            lo, _ = self.plan["own"]
            self.K_red._sl = None
            self.sym_lattice = self.K_red.sym_lattice(row_node_offset=lo // 2)
            self.A = self.K_red.struct()
            self.ops.A = self.A
            self._sl_stale = False

    def cg_iterations(self, k):
        self._refresh_operand()
        if self.fused is not None:
            self.fused.iterations(k)
            self.st["iters"] += int(k)
            return
        if self.fused_dist:   # one kernel per iteration: halo of u and the scalar exchange inside it
            call("psfem_dist_cg_iterations", C.byref(self.A), C.byref(self.cg_ctx), int(k), self.launches, self.seq + 1,
                 ptr(self.cg_ws), self.cg_ws_bytes, stream_ptr())
            self.seq += int(k)
            self.launches += int(k)
            self.st["iters"] += int(k)
            return
        if self.fabric:   # three kernels per iteration, no collective launches: scalars and halo go through peer memory
            call("psfem_dist_pcg_iterations", C.byref(self.A), C.byref(self.ctx), int(k), self.seq + 1, stream_ptr())
            self.seq += int(k)
            self.st["iters"] += int(k)
        else:
            pcg_iterations(self.ops, self.comm, self.st, k)

    def spmv_only(self, k):
        self._refresh_operand()
        lo, _ = self.plan["own"]
        for _ in range(k):
            self.ops.spmv_dot(self.st["p_ext"], lo, self.st["q"], self.st["pq"])

    def residual(self):
        """(||r||, ||b||) from the recurrence (global)."""
        if self.fused is not None:
            st = self.fused.state()
            return st["relres"] * st["bnorm"], st["bnorm"]
        if self.fused_dist:
            h = (C.c_double * 5)()
            call("psfem_dist_cg_state", C.byref(self.cg_ctx), self.n_owned_free, ptr(self.cg_ws), h, stream_ptr())
            return float(h[1] * h[2]), float(h[2])
        if self.fabric:
            h = (C.c_double * 3)()
            call("psfem_dist_pcg_residual", C.byref(self.ctx), self.seq, h, stream_ptr())
            return float(np.sqrt(h[1])), float(np.sqrt(self._bb))
        rr = self.st["out2"][1].item() if self.st["iters"] else self.st["out3"][1].item()
        bb = self.st["out3"][2].item()
        return float(np.sqrt(rr)), float(np.sqrt(bb))

    def apply(self, x_own):
        """y = K x over all strips (owned rows in and out, the halo columns travel inside the kernel); fused strip solver only,
        collective, overwrites the recurrence vectors."""
        torch = _capi.torch_cuda()
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        y = torch.empty_like(x_own)
        self.seq += 1
        call("psfem_dist_cg_apply", C.byref(self.A), C.byref(self.cg_ctx), ptr(x_own), ptr(y), self.seq, ptr(self.cg_ws),
             self.cg_ws_bytes, stream_ptr())
        return y

    def _restart_from_x(self):
        """||b - K x|| of the current solution, evaluated (not recurred), with the solve left started FROM x: the residual the
        recurrence carries drifts away from the true one over 10^4..10^5 iterations.  One GPU: the fused solver starts again
        at x.  Strips: x so far is set aside, y = K x by an apply pass, and the solve starts on the correction equation
        K d = b - K x (the solution is x + d)."""
        torch = _capi.torch_cuda()
        self.finish_x()
        if self.fused is not None:
            self.fused.start(self.b_own, self.st["x"], rtol=0.0)
            self.st["iters"] = 0
            st = self.fused.state()
            return st["relres"] * st["bnorm"]
        import torch.distributed as dist
        x = self.st["x"]
        if getattr(self, "_x_acc", None) is None:
            self._x_acc = x.clone()
        else:
            self._x_acc += x
        rt = self.b_own - self.apply(self._x_acc)
        if self.masked:
            rt *= (self._new_id >= 0).to(rt.dtype)
        rr = (rt * rt).sum().reshape(1)
        dist.all_reduce(rr)
        self._b_cycle = rt
        self.start(b=rt)
        return float(rr.sqrt().item())

    def _fold_x(self):
        """strips after _restart_from_x: the solution is the part set aside + the correction."""
        if getattr(self, "_x_acc", None) is not None:
            self.finish_x()
            self.st["x"] += self._x_acc
            self._x_acc = None

    def solve(self, rtol=1e-8, maxit=200000, check_every=100, polish=None):
        """Continue the PCG until ||r|| <= rtol ||b|| (host checks every check_every iterations).  With the fused solvers the
        TRUE residual b - K x is evaluated after convergence and the solve continued from x while it is above the tolerance
        and still improving by a factor 2 per cycle (at most `polish` cycles, default PSFEM_PCG_POLISH or 3; negative: no
        evaluation).  relres: the recurrence residual at the first convergence; relres_true: the last evaluated one."""
        torch = _capi.torch_cuda()
        t0 = time.perf_counter()
        if polish is None:
            polish = int(os.environ.get("PSFEM_PCG_POLISH", "3"))
        can_check = self.fused is not None or self.fused_dist
        total, cycles, polish_iters = 0, 0, 0
        bnorm = first = true = prev = None
        while True:
            while total + self.st["iters"] < maxit:
                r, b = self.residual()
                if r <= rtol * (b if bnorm is None else bnorm):
                    break
                self.cg_iterations(check_every)
            torch.cuda.synchronize()
            r, b = self.residual()
            if bnorm is None:
                bnorm, first = b, (r / b if b > 0 else 0.0)
            total += self.st["iters"]
            if cycles:
                polish_iters += self.st["iters"]
            if not can_check or polish < 0 or not r <= rtol * bnorm:
                self.st["iters"] = total
                break
            true = self._restart_from_x()          # st["iters"] = 0 again
            if true <= rtol * bnorm or cycles >= polish or (cycles > 0 and not true < 0.5 * prev) or total >= maxit:
                self.st["iters"] = total
                break
            prev = true
            cycles += 1
        self._fold_x()
        out = dict(iterations=total, relres=first, seconds=time.perf_counter() - t0)
        if true is not None:
            out.update(relres_true=true / bnorm if bnorm > 0 else 0.0, polish_cycles=cycles, polish_iterations=polish_iters)
        return out

    def finish_x(self):
        """The fused PCG writes x every other iteration; add the pending half before x is read."""
        if getattr(self, "fused", None) is not None:
            self.fused.finish()
        elif getattr(self, "fused_dist", False):
            call("psfem_cg_sl_finish", ptr(self.st["x"]), self.n_owned_free, ptr(self.cg_ws), stream_ptr())

    def gather_solution(self, dst=0):
        """Global displacement vector (zeros at constrained DOFs, like reconstruct_full_vector Polygones.py:913-921)
        on rank `dst` (None elsewhere).  Ranks own contiguous, ascending DOF ranges, so the owned solutions concatenated in
        rank order ARE the reduced global vector: NCCL transfers of the device buffers into one buffer on `dst`, one scatter
        over the free DOFs on the device (psfem_scatter_free), one D2H into page-locked memory."""
        torch = _capi.torch_cuda()
        from .device import FreeMap, to_host
        self.finish_x()
        x = self.st["x"]
        if self.world == 1:
            red = x
        else:
            import torch.distributed as dist
            sizes = [None] * self.world
            dist.all_gather_object(sizes, int(x.numel()))
            # the strips own different numbers of columns (the last one also owns the right edge): point-to-point
            # transfers into slices of one device buffer (a NCCL gather wants equal sizes)
            if self.rank == dst:
                red = torch.empty(sum(sizes), dtype=torch.float64, device="cuda")
                parts = list(red.split(sizes))
                parts[dst].copy_(x)
                ops = [dist.P2POp(dist.irecv, parts[r], r) for r in range(self.world) if r != dst]
                for w_ in dist.batch_isend_irecv(ops):
                    w_.wait()
            else:
                for w_ in dist.batch_isend_irecv([dist.P2POp(dist.isend, x, dst)]):
                    w_.wait()
                return None
        if self.masked:                      # masked strips keep every DOF of their columns (zeros at the constrained ones)
            full = red
        else:
            fm = FreeMap(self.n_dof_global, self.constrained_global)
            full = fm.expand(red)
        h = to_host(full)
        torch.cuda.current_stream().synchronize()
        return h.numpy()

    def full_solve(self, rtol):
        torch = _capi.torch_cuda()
        self.start()
        torch.cuda.synchronize()
        return self.solve(rtol)

    def e2e_distributed(self, iters, steps):
        """End-to-end at N>1: every step uploads the rank's strip mesh from pinned host memory,
        assembles, reduces, runs `iters` PCG iterations and reads the residual norm back."""
        torch = _capi.torch_cuda()
        import torch.distributed as dist
        nodes_h = self.nodes.cpu().pin_memory()
        conn_h = self.conn.cpu().pin_memory()
        self.release()
        torch.cuda.empty_cache()
        times = []
        for s in range(steps + 1):
            if self.world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nodes = nodes_h.cuda(non_blocking=True)
            conn = conn_h.cuda(non_blocking=True)
            self.setup((nodes, conn))
            self.cg_iterations(iters)
            r, b = self.residual()
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()
            times.append(time.perf_counter() - t0)
            self.release()
        t = torch.tensor([sum(times[1:])], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"value": self.n_free_global * iters * steps / float(t.item()), "unit": "DOF*iter/s",
                "h2d_bytes_per_step": int(nodes_h.numel() * 8 + conn_h.numel() * 4), "d2h_bytes_per_step": 16,
                "steps": steps, "pcg_iters_per_call": iters,
                "calls": "per rank: pinned host strip mesh -> H2D -> pattern + assembly + BC reduction -> PCG iterations -> residual D2H"}

    def release(self):
        if getattr(self, "fabric", None) is not None:
            self.st["p_ext"] = None
            self.fabric.close()
            self.fabric = None
        for k in ("asm", "K_vals", "K_ext_red", "K_red", "fm", "st", "ops", "A", "b_own", "nodes", "conn", "ctx", "_cd_owned", "fused", "cg_ws", "cg_ctx", "_x_acc", "_b_cycle"):
            if hasattr(self, k):
                delattr(self, k)


# ------------------------------------------------------------------------------------------------
# public multi-GPU entry point
# ------------------------------------------------------------------------------------------------
def static_solve(L, l, N, M=0, bc=None, F=None, E=210e9, nu=0.3, t=0.1, element_type="quad", mesh_type="coarse",
                 nodes=None, rtol=1e-8, maxit=200000, check_every=100, dst=0, raise_on_fail=True, transport=None,
                 lattice_only=None):
    """Static solve of a `Polygones.mesh(L, l, N, M, element_type=..., mesh_type=...)` plate on ALL ranks of the default
    torch.distributed process group (one process per GPU): the multi-GPU counterpart of `drivers.static_solve`, i.e. of
    mesh -> global_stiffness_matrix -> apply_boundary_conditions -> solve -> reconstruct (Statictest.py:49-75).

    Every rank calls it with the same arguments, all HOST data:
      bc     boundary conditions like the reference's lists [(node, [ux, uy]), ...] (default: left edge clamped),
      F      global load vector (2 * nodes), or a callable (dof_lo, dof_hi) -> slice (default: edgeForces on the right edge),
      nodes  optional (n*m, 2) coordinates replacing the regular ones (a deformed / jittered structured mesh); each rank
             uploads only the rows of its strip.
    The strip of each rank is meshed, assembled and reduced on its GPU without communication; the PCG runs with the
    collectives fused into its kernels over peer memory (NCCL when peer access is unavailable).
    lattice_only (default: on for Q4 meshes when the fused PCG can run): K is assembled straight into the symmetric-lattice
    operand of the solver and never exists as CSR (StripProblem.setup(lattice_only=True)); same iterates, bit for bit.
    Returns (U, info): U = the global displacement vector with zeros at constrained DOFs (reconstruct_full_vector,
    Polygones.py:913-921) on rank `dst`, None elsewhere; info = dict(iterations, relres, converged, seconds, transport,
    h2d_bytes, d2h_bytes of this rank)."""
    import torch.distributed as dist
    torch = _capi.torch_cuda()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    M = M or N
    prob = StripProblem(L=L, l=l, N=N, M=M, rank=rank, world=world, E=E, nu=nu, t=t, element_type=element_type,
                        mesh_type=mesh_type, bc=bc, F=F, transport=transport)
    h2d = 0
    nodes_conn = None
    if nodes is not None:
        from .device import device_mesh
        p = prob.part
        if tuple(nodes.shape) != (p.n * p.m, 2):
            raise ValueError(f"nodes must have shape ({p.n * p.m}, 2) for this mesh")
        _, conn = device_mesh(L, l, N, M, 0, 0, element_type, mesh_type, col0=p.e0, ncols=p.e1 - p.e0, cell0=p.cell0,
                              ncells=p.cell1 - p.cell0)
        part = nodes[p.e0 * p.m:p.e1 * p.m]
        part = part if isinstance(part, np.ndarray) and part.dtype == np.float64 else np.ascontiguousarray(part, dtype=np.float64)
        nd = torch.from_numpy(part).to("cuda", non_blocking=True)
        h2d += part.nbytes
        nodes_conn = (nd, conn)
    from .device import fused_cg_enabled
    if lattice_only is None:
        lattice_only = ((element_type, mesh_type) == ("quad", "coarse") and fused_cg_enabled()
                        and (world == 1 or prob.transport == "peer"))
    if lattice_only:
        try:
            prob.setup(nodes_conn, lattice_only=True)
        except RuntimeError as e:                           # no peer transport on this box (decided by ALL ranks together)
            if "lattice_only" not in str(e):
                raise
            prob.setup(nodes_conn)
    else:
        prob.setup(nodes_conn)
    if F is not None:
        h2d += 8 * prob.part.n_ext_dof
    info = prob.solve(rtol=rtol, maxit=maxit, check_every=check_every)
    info["converged"] = bool(info["relres"] <= rtol)
    info["transport"] = prob.transport if world > 1 else "single GPU"
    info["lattice_only"] = bool(getattr(prob, "lattice_only", False))
    U = prob.gather_solution(dst=dst)
    info["h2d_bytes"] = int(h2d)
    info["d2h_bytes"] = int(8 * prob.n_dof_global) if rank == dst else 0       # the gathered vector leaves the device on rank dst only
    prob.release()
    if raise_on_fail and not info["converged"]:
        raise _capi.ConvergenceError(f"strip PCG: relres {info['relres']:.3e} after {info['iterations']} iterations")
    return U, info
