"""Static, Newmark-beta and modal drivers.

The reference has no driver functions: these are the module-level arithmetic of its scripts turned
into functions, with the dense LAPACK calls replaced by device kernels:
  static_solve : TestBeam/Statictest.py:66-78, StaticTest2.py:66-81, Beamexemple.py:59-65
  newmark      : main.py:226-332 + 377-385 (reproduced literally, quirks included -- SURVEY A.9)
  modal        : freeModes.py:53-77 (Beamexemple.py:86-160)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import call, check, lib, ptr, stream_ptr, workspace
from .device import (SL_MIN_ROWS, Blas, DeviceCSR, FreeMap, FusedCG, constrained_dofs_of, continue_from_true_residual, direct_limits, direct_solve, half_bandwidth,
                     pcg_jacobi, residual_dd, to_host)
from .Polygones import device_copy


def _scaled(x, a, out=None):
    """out = a*x on the device (psfem_axpby)."""
    torch = _capi.torch_cuda()
    if out is None:
        out = torch.empty_like(x)
    call("psfem_axpby", x.numel(), float(a), ptr(x), 0.0, ptr(None), ptr(out), stream_ptr())
    return out


def _to_dev(v):
    torch = _capi.torch_cuda()
    if isinstance(v, np.ndarray) or not hasattr(v, "data_ptr"):
        return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).cuda()
    return v


# ---------------------------------------------------------------------------------------------
# static
# ---------------------------------------------------------------------------------------------
MASKED_MIN_ROWS = 16384      # below: the single-CTA solver on the reduced system


def _pcg_masked(Kd: DeviceCSR, fm: FreeMap, F_dev, rtol, maxit, check_every, raise_on_fail=True):
    torch = _capi.torch_cuda()
    _capi.bind_device()
    n = Kd.n_rows
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    if maxit is None:
        maxit = 10 * n + 100
    Kd.sym_lattice()              # the unreduced matrix of a Polygones.mesh mesh is a symmetric lattice matrix
    if FusedCG.usable(Kd):        # one kernel per iteration (constrained DOFs as masked rows)
        rc, out = FusedCG(Kd, new_id=fm.new_id).solve(F_dev, x, rtol, maxit, check_every)
        out.update(method="pcg", masked=True)
        if rc == _capi.ENOCONV and not raise_on_fail:
            return x, out
        check(rc)
        return x, out
    s = Kd.struct()
    sz = C.c_size_t(0)
    call("psfem_pcg_workspace", n, s.n_blocks, C.byref(sz))
    ws = workspace(sz.value)
    info = (C.c_double * 4)()
    rc = lib.psfem_pcg_jacobi_masked(C.byref(s), ptr(F_dev), ptr(x), ptr(fm.new_id), float(rtol), int(maxit), int(check_every),
                                     info, ptr(ws), sz.value, stream_ptr())
    out = dict(method="pcg", iterations=int(info[0]), relres=float(info[1]), bnorm=float(info[2]), converged=(rc == _capi.OK), masked=True,
               pcg="classic")
    if rc == _capi.ENOCONV and not raise_on_fail:
        return x, out
    check(rc)
    if out["bnorm"] > 0:
        cd_idx = fm.constrained_dev.long() if fm.constrained_dev is not None else None

        def again():
            h = (C.c_double * 4)()
            rc2 = lib.psfem_pcg_jacobi_masked(C.byref(s), ptr(F_dev), ptr(x), ptr(fm.new_id), float(rtol), int(maxit), int(check_every),
                                              h, ptr(ws), sz.value, stream_ptr())
            return rc2, int(h[0])

        def true_residual():
            r = Kd.matvec(x, alpha=-1.0, gamma=1.0, z=F_dev)
            if cd_idx is not None:
                r.index_fill_(0, cd_idx, 0.0)          # rows of constrained DOFs are not part of the system
            return float(torch.linalg.vector_norm(r).item()) / out["bnorm"]
        continue_from_true_residual(again, true_residual, rtol, out)
    return x, out


def _norm(blas: Blas, v):
    return float(np.sqrt(max(blas.dot(v, v), 0.0)))


def _refine(solve, residual, x, bnorm, rounds, info):
    """Iterative refinement with a compensated residual: r = b - K x in double-double (psfem_residual_dd), K d = r by
    the inner solver, x += d.  Stops early once the true relative residual is at the rounding level."""
    blas = Blas()
    done = 0
    for _ in range(int(rounds)):
        r = residual(x)
        rel = _norm(blas, r) / bnorm if bnorm > 0 else 0.0
        info["relres_true"] = rel
        if rel <= 4 * np.finfo(float).eps:
            break
        d, inner = solve(r)
        info["iterations"] += inner["iterations"]
        Blas.axpby(1.0, x, 1.0, d, x)
        done += 1
    info["refinement_rounds"] = done
    r = residual(x)
    info["relres_true"] = _norm(blas, r) / bnorm if bnorm > 0 else 0.0
    return x


def static_solve(K, F, bc, rtol=1e-8, maxit=None, check_every=50, return_device=False, method="auto", refine=None,
                 raise_on_fail=True):
    """Solve K U = F with the DOFs of `bc` fixed to zero.

    Same steps as Statictest.py:66-75: apply_boundary_conditions -> solve -> scatter back with zeros at constrained DOFs.
    The dense LU of the reference (`np.linalg.solve`, :69) is replaced on the device by
      method="direct": banded Cholesky + iterative refinement on a double-double residual (psfem_direct_solve) -- small
                       systems (<= 16 384 free DOFs, half bandwidth <= 320), accurate to rounding like the LU it replaces;
      method="pcg":    Jacobi-PCG (x0 = 0, stop at ||r|| <= rtol ||b||), the path of the large plates; `refine` > 0 adds
                       that many rounds of refinement with the compensated residual around it (each an inner PCG solve);
      method="auto":   direct when the reduced system fits the band solver, else pcg.
    Returns (U, U_reduced, constrained_dofs, info)."""
    torch = _capi.torch_cuda()
    if method not in ("auto", "direct", "pcg"):
        raise ValueError("method must be 'auto', 'direct' or 'pcg'")
    constrained = constrained_dofs_of(bc)
    Kd = device_copy(K)
    fm = FreeMap(Kd.n_rows, constrained)
    max_rows, max_hb = direct_limits()
    K_red = None
    hb = None
    if method != "pcg" and 0 < fm.n_free <= max_rows:
        K_red = fm.reduce_csr(Kd)
        hb = half_bandwidth(K_red)
        if hb > max_hb:
            if method == "direct":
                raise ValueError(f"method='direct': half bandwidth {hb} exceeds the band solver's limit {max_hb}")
            method = "pcg"
        else:
            method = "direct"
    elif method == "direct":
        raise ValueError(f"method='direct': {fm.n_free} free DOFs exceed the band solver's limit {max_rows}")
    else:
        method = "pcg"
    if method == "direct":
        F_red = fm.reduce_vec(_to_dev(F))
        U_red, info = direct_solve(K_red, F_red, refine=2 if refine is None else refine, hb=hb)
        U = fm.expand(U_red)
    elif Kd.n_rows > MASKED_MIN_ROWS and Kd.plan()[3] is not None:
        # large system whose matrix has the 2x2 node-block structure: solve on the UNREDUCED matrix with the
        # constrained DOFs masked out (psfem_pcg_jacobi_masked) -- same iterates on the free DOFs as the reduced
        # system of Polygones.py:1003, the 9-byte SpMV for any boundary conditions, no reduced copy of K
        F_dev = _to_dev(F)
        U, info = _pcg_masked(Kd, fm, F_dev, rtol, maxit, check_every, raise_on_fail)
        if refine:
            cd_idx = fm.constrained_dev.long() if fm.constrained_dev is not None else None

            def masked_residual(x):
                r = residual_dd(Kd, x, F_dev)
                if cd_idx is not None:
                    r.index_fill_(0, cd_idx, 0.0)          # rows of constrained DOFs are not part of the system (reactions)
                return r
            _refine(lambda r: _pcg_masked(Kd, fm, r, rtol, maxit, check_every, raise_on_fail), masked_residual, U,
                    info["bnorm"], refine, info)
        U_red = fm.reduce_vec(U)
    else:
        if K_red is None:
            K_red = fm.reduce_csr(Kd)
        F_red = fm.reduce_vec(_to_dev(F))
        U_red, info = pcg_jacobi(K_red, F_red, rtol=rtol, maxit=maxit, check_every=check_every, raise_on_fail=raise_on_fail)
        info["method"] = "pcg"
        if refine:
            _refine(lambda r: pcg_jacobi(K_red, r, rtol=rtol, maxit=maxit, check_every=check_every, raise_on_fail=raise_on_fail),
                    lambda x: residual_dd(K_red, x, F_red), U_red, info["bnorm"], refine, info)
        U = fm.expand(U_red)
    if return_device:
        return U, U_red, constrained, info
    Uh, Urh = to_host(U), to_host(U_red)          # page-locked landing buffers: DMA instead of the pageable path
    torch.cuda.current_stream().synchronize()
    return Uh.numpy(), Urh.numpy(), constrained, info


# ---------------------------------------------------------------------------------------------
# Newmark-beta (main.py)
# ---------------------------------------------------------------------------------------------
def ramp_load(n_nodes, L, P=1e6, T=10.0, dt=1e-3, accel_ratio=3):
    """Load history of main.py:201,208-209,226-236: every node gets F_y(t) = -(accel_ratio*P*t/T)*L/n
    up to step T0 = nbPoints//accel_ratio, then -P*L/n; F_x = 0.
    Returns (shape (2n,) with ones at y DOFs, scale (nbPoints,), times)."""
    nb = int(T / dt) + 1
    times = np.linspace(0, T, nb)
    T0 = nb // accel_ratio
    scale = np.empty(nb)
    scale[:T0 + 1] = -(accel_ratio * P * times[:T0 + 1] / T) * L / n_nodes
    scale[T0 + 1:] = -P * L / n_nodes
    shape = np.zeros(2 * n_nodes)
    shape[1::2] = 1.0
    return shape, scale, times


def newmark(K, M, bc, f_shape, f_scale, dt=1e-3, gamma=0.5, beta=0.25, alpha_R=0.1, beta_R=0.01,
            store_every=1, energies=True, rtol=1e-13, maxit=None):
    """Newmark-beta exactly as main.py:264-332 implements it, on the device.

    K, M: full matrices (CSR); bc: boundary conditions (reduced with apply_boundary_conditions /
    reduce, main.py:56-57); load of step i = f_shape * f_scale[i] (full-size shape vector).
    Returns a dict: U, V, A (n_free, n_stored) like main.py's U_reduced / V_reduced / A_reduced,
    kinetic/potential/total energy series (main.py:377-385), constrained_dofs, info."""
    torch = _capi.torch_cuda()
    _capi.bind_device()
    constrained = constrained_dofs_of(bc)
    Kd, Md = device_copy(K), device_copy(M)
    fm = FreeMap(Kd.n_rows, constrained)
    K_red = fm.reduce_csr(Kd)
    M_red = fm.reduce_csr(Md)
    if M_red.nnz != K_red.nnz:
        raise ValueError("K and M must share one sparsity pattern (assemble both with this package)")
    n = K_red.n_rows
    f_scale = np.ascontiguousarray(f_scale, dtype=np.float64)
    n_steps = int(f_scale.size)
    fshape_red = fm.reduce_vec(_to_dev(f_shape))
    f64 = dict(dtype=torch.float64, device="cuda")
    # K_eff = K + M/(beta dt^2)   (main.py:290; no damping term)
    Keff_vals = torch.empty(K_red.nnz, **f64)
    call("psfem_axpby", K_red.nnz, 1.0, ptr(K_red.vals), 1.0 / (beta * dt ** 2), ptr(M_red.vals), ptr(Keff_vals), stream_ptr())
    # a0 = M^-1 (F0 - K u0), u0 = 0   (main.py:286)
    f0 = torch.empty(n, **f64)
    call("psfem_axpby", n, float(f_scale[0]), ptr(fshape_red), 0.0, ptr(None), ptr(f0), stream_ptr())
    a0, _ = pcg_jacobi(M_red, f0, rtol=rtol, maxit=maxit)
    n_store = (n_steps - 1) // store_every + 1
    U = torch.empty((n_store, n), **f64)
    V = torch.empty((n_store, n), **f64)
    A = torch.empty((n_store, n), **f64)
    E = torch.empty((n_steps, 3), **f64) if energies else None
    Keff = K_red.with_vals(Keff_vals)
    if n > SL_MIN_ROWS:           # large lattice systems: symmetric-lattice storage for the PCG operand and the two of the right-hand side
        for A_ in (K_red, M_red, Keff):
            A_.sym_lattice()
    s, sM, sKeff = K_red.struct(), M_red.struct(), Keff.struct()
    sz = C.c_size_t(0)
    call("psfem_newmark_workspace", n, s.n_blocks, C.byref(sz))
    ws = workspace(sz.value)
    info = (C.c_double * 4)()
    scale_c = f_scale.ctypes.data_as(C.POINTER(C.c_double))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    rc = lib.psfem_newmark_run(C.byref(s), C.byref(sM), C.byref(sKeff), ptr(fshape_red), scale_c, n_steps,
                               float(dt), float(gamma), float(beta), float(alpha_R), float(beta_R), ptr(a0),
                               int(store_every), ptr(U), ptr(V), ptr(A), ptr(E), float(rtol),
                               int(maxit if maxit is not None else 10 * n + 100), info, ptr(ws), sz.value, stream_ptr())
    ev1.record()
    check(rc)
    torch.cuda.current_stream().synchronize()
    out = dict(U=U.cpu().numpy().T, V=V.cpu().numpy().T, A=A.cpu().numpy().T, constrained_dofs=constrained,
               info=dict(pcg_iterations=int(info[0]), max_relres=float(info[1]), stabiliser_triggers=int(info[2]),
                         run_ms=float(ev0.elapsed_time(ev1))))
    if energies:
        e = E.cpu().numpy()
        out.update(kinetic_energy=e[:, 0].copy(), potential_energy=e[:, 1].copy(), total_energy=e[:, 2].copy())
    return out


def explicit_euler(K, M, bc, f_shape, f_scale, dt=1e-3, store_every=1, energies=True, rtol=1e-13, maxit=None):
    """The modified explicit Euler scheme of main.py:334-372 (its `use_newmark = False` branch), on the device:
    u_i = u_{i-1} + dt v_{i-1}; a_i = M^-1 (F_i - K u_i); v_i = v_{i-1} + dt/2 (a_{i-1} + a_i); the stabiliser of
    :365-370 (threshold 10).  Arguments and the returned dict are those of `newmark`; the dense `np.linalg.solve` with M
    per step (:359) is a Jacobi-PCG solve warm-started from the previous acceleration (small banded systems: the
    banded Cholesky factor of M inside one persistent kernel).  Constant dt (the reference takes times[i] - times[i-1]
    of a linspace, equal to dt up to rounding)."""
    torch = _capi.torch_cuda()
    _capi.bind_device()
    constrained = constrained_dofs_of(bc)
    Kd, Md = device_copy(K), device_copy(M)
    fm = FreeMap(Kd.n_rows, constrained)
    K_red = fm.reduce_csr(Kd)
    M_red = fm.reduce_csr(Md)
    if M_red.nnz != K_red.nnz:
        raise ValueError("K and M must share one sparsity pattern (assemble both with this package)")
    n = K_red.n_rows
    f_scale = np.ascontiguousarray(f_scale, dtype=np.float64)
    n_steps = int(f_scale.size)
    fshape_red = fm.reduce_vec(_to_dev(f_shape))
    f64 = dict(dtype=torch.float64, device="cuda")
    f0 = torch.empty(n, **f64)
    call("psfem_axpby", n, float(f_scale[0]), ptr(fshape_red), 0.0, ptr(None), ptr(f0), stream_ptr())
    a0, _ = pcg_jacobi(M_red, f0, rtol=rtol, maxit=maxit)            # main.py:344 (u0 = 0)
    n_store = (n_steps - 1) // store_every + 1
    U = torch.empty((n_store, n), **f64)
    V = torch.empty((n_store, n), **f64)
    A = torch.empty((n_store, n), **f64)
    En = torch.empty((n_steps, 3), **f64) if energies else None
    if n > SL_MIN_ROWS:
        for A_ in (K_red, M_red):
            A_.sym_lattice()
    s, sM = K_red.struct(), M_red.struct()
    sz = C.c_size_t(0)
    call("psfem_newmark_workspace", n, s.n_blocks, C.byref(sz))
    ws = workspace(sz.value)
    info = (C.c_double * 4)()
    rc = lib.psfem_explicit_run(C.byref(s), C.byref(sM), ptr(fshape_red), f_scale.ctypes.data_as(C.POINTER(C.c_double)), n_steps,
                                float(dt), ptr(a0), int(store_every), ptr(U), ptr(V), ptr(A), ptr(En), float(rtol),
                                int(maxit if maxit is not None else 10 * n + 100), info, ptr(ws), sz.value, stream_ptr())
    check(rc)
    torch.cuda.current_stream().synchronize()
    out = dict(U=U.cpu().numpy().T, V=V.cpu().numpy().T, A=A.cpu().numpy().T, constrained_dofs=constrained,
               info=dict(pcg_iterations=int(info[0]), max_relres=float(info[1]), stabiliser_triggers=int(info[2])))
    if energies:
        e = En.cpu().numpy()
        out.update(kinetic_energy=e[:, 0].copy(), potential_energy=e[:, 1].copy(), total_energy=e[:, 2].copy())
    return out


# ---------------------------------------------------------------------------------------------
# modal (freeModes.py)
# ---------------------------------------------------------------------------------------------
class _ShiftInvertOperator:
    """w = K_hat^-1 y for the Lanczos iteration.  Small banded systems: K_hat is factored ONCE (banded Cholesky,
    psfem_direct_factor) and every application is two substitutions plus one refinement round on a double-double residual
    -- what the reference gets from LAPACK inside eigh / eigsh(sigma=0).  Everything else: a Jacobi-PCG solve per application
    (the fused one-kernel iteration on symmetric-lattice operands)."""

    def __init__(self, Kh: DeviceCSR, pcg_rtol):
        torch = _capi.torch_cuda()
        self.Kh, self.pcg_rtol = Kh, pcg_rtol
        self.pcg_iterations = 0
        self.direct = False
        max_rows, max_hb = direct_limits()
        if 0 < Kh.n_rows <= max_rows:
            hb = half_bandwidth(Kh)
            if hb <= max_hb:
                sz = C.c_size_t(0)
                call("psfem_direct_workspace", Kh.n_rows, hb, C.byref(sz))
                self.ws, self.ws_bytes, self.hb = workspace(sz.value), sz.value, hb
                s_ = Kh.struct()
                rc = lib.psfem_direct_factor(C.byref(s_), hb, ptr(self.ws), sz.value, stream_ptr())
                if rc == _capi.OK:
                    self.direct = True
                elif rc != _capi.ENOCONV:                     # ENOCONV: not positive definite -> let PCG report it
                    check(rc)
        self.out = torch.empty(Kh.n_rows, dtype=torch.float64, device="cuda")

    def __call__(self, y):
        if self.direct:
            torch = _capi.torch_cuda()
            x = torch.empty_like(y)
            info = (C.c_double * 5)()
            s_ = self.Kh.struct()
            call("psfem_direct_substitute", C.byref(s_), ptr(y), ptr(x), self.hb, 1, info, ptr(self.ws), self.ws_bytes, stream_ptr())
            return x
        x, inf = pcg_jacobi(self.Kh, y, rtol=self.pcg_rtol)
        self.pcg_iterations += inf["iterations"]
        return x


def modal(K, M, bc, k=20, tol=1e-10, max_lanczos=None, pcg_rtol=1e-14, max_restarts=200):
    """First k natural frequencies and mode shapes.

    freeModes.py:53-77: reduce, scale K and M by their largest |diagonal| entry, generalized symmetric eigenproblem, clip
    at 0, f = sqrt(lambda * k_max/m_max)/(2 pi), modes rebuilt to full length as rows.  The dense LAPACK `eigh` (:74, all
    modes, O(n^3)) is replaced by a THICK-RESTART shift-invert Lanczos iteration (sigma = 0) in the M-inner product with
    full reorthogonalisation: the basis is bounded (`max_lanczos`, default max(2k + 20, 60) vectors -- 2 x 61 vectors of
    n doubles instead of an unbounded Krylov basis), after every cycle the best k + 10 Ritz vectors are kept and the
    iteration continues from them.  Operator applications: device SpMV with M_hat and a solve with K_hat (factored once
    for small banded systems, Jacobi-PCG otherwise).  Eigenvectors are M_hat-orthonormal like LAPACK's (sign arbitrary).
    Returns (frequencies (k,), modes (k, total_dofs), constrained_dofs, info); warns when the Ritz values have not
    converged to `tol` within `max_restarts` cycles."""
    import warnings
    torch = _capi.torch_cuda()
    _capi.bind_device()
    constrained = constrained_dofs_of(bc)
    Kd, Md = device_copy(K), device_copy(M)
    fm = FreeMap(Kd.n_rows, constrained)
    K_red = fm.reduce_csr(Kd)
    M_red = fm.reduce_csr(Md)
    n = K_red.n_rows
    k = min(k, n)
    f64 = dict(dtype=torch.float64, device="cuda")
    blas = Blas()
    torch.cuda.current_stream().synchronize()
    k_max = float(torch.max(torch.abs(K_red.diagonal())).item())   # freeModes.py:58-59
    m_max = float(torch.max(torch.abs(M_red.diagonal())).item())
    scaling = k_max / m_max                                        # :60
    Kh = K_red.with_vals(_scaled(K_red.vals, 1.0 / k_max))         # :62-63
    Mh = M_red.with_vals(_scaled(M_red.vals, 1.0 / m_max))
    if n > SL_MIN_ROWS:
        Mh.sym_lattice()
    op = _ShiftInvertOperator(Kh, pcg_rtol)
    m_max_basis = int(min(n, max_lanczos or max(2 * k + 20, 60)))
    keep = min(k + 10, max(m_max_basis - 2, k))                    # Ritz vectors kept at a restart
    Vb = torch.zeros((m_max_basis + 1, n), **f64)                  # basis (rows), M-orthonormal
    MV = torch.zeros((m_max_basis + 1, n), **f64)                  # M_hat @ basis
    coef = torch.empty(m_max_basis + 1, **f64)
    H = np.zeros((m_max_basis + 1, m_max_basis + 1))               # projected operator V^T M (K^-1 M) V, built column by column
    rng = np.random.default_rng(12345)
    v = torch.from_numpy(rng.standard_normal(n)).cuda()
    v = op(Mh.matvec(v))                                           # start in range(K^-1 M): no components in the null space of M
    mv = Mh.matvec(v)
    nrm = np.sqrt(blas.dot(v, mv))
    _scaled(v, 1.0 / nrm, Vb[0])
    _scaled(mv, 1.0 / nrm, MV[0])
    theta, Ssel, resid = None, None, None
    converged = False
    j, steps, restarts = 0, 0, 0
    while True:
        # ---- one Lanczos step: w = K^-1 M v_j, orthogonalised against v_0..v_j in the M inner product (twice) ----
        w = op(MV[j])
        col = np.zeros(j + 1)
        for _ in range(2):
            blas.multi_dot(MV, j + 1, w, coef)
            col += coef[: j + 1].cpu().numpy()                     # one host read per pass: the column of H (and the loop needs beta)
            blas.multi_axpy(Vb, j + 1, coef, w)
        H[: j + 1, j] = col
        mw = Mh.matvec(w)
        b = float(np.sqrt(max(blas.dot(w, mw), 0.0)))
        steps += 1
        jn = j + 1                                                  # vectors in the basis
        full = jn == m_max_basis or b < 1e-300
        if jn >= k and (jn % 5 == 0 or full):
            Hs = 0.5 * (H[:jn, :jn] + H[:jn, :jn].T)
            th, S = np.linalg.eigh(Hs)
            order = np.argsort(-th)
            th, S = th[order], S[:, order]
            if np.all(th[:k] > 0):
                resid = np.abs(b * S[-1, :k]) / np.abs(th[:k])
                theta, Ssel = th[:k], S[:, :k]
                if np.all(resid < tol) or b < 1e-300:
                    converged = True
                    break
            if full:
                if restarts >= max_restarts:
                    break
                # ---- thick restart: keep the best `keep` Ritz vectors, continue from the next Lanczos vector ----
                l = min(keep, jn - 1)
                Y = torch.zeros((l, n), **f64)
                MY = torch.zeros((l, n), **f64)
                Sd = torch.from_numpy(np.ascontiguousarray(-S[:, :l].T)).cuda()      # negated: multi_axpy subtracts
                for i in range(l):
                    call("psfem_multi_axpy", jn, n, ptr(Vb), ptr(Sd[i]), ptr(Y[i]), stream_ptr())
                    call("psfem_multi_axpy", jn, n, ptr(MV), ptr(Sd[i]), ptr(MY[i]), stream_ptr())
                Vb[:l].copy_(Y)
                MV[:l].copy_(MY)
                _scaled(w, 1.0 / b, Vb[l])
                _scaled(mw, 1.0 / b, MV[l])
                H[:] = 0.0
                H[np.arange(l), np.arange(l)] = th[:l]
                H[l, :l] = b * S[-1, :l]                                   # K^-1 M y_i = theta_i y_i + (beta s_mi) v_next
                del Y, MY, Sd
                j = l
                restarts += 1
                continue
        if b < 1e-300:
            break
        H[jn, j] = b
        _scaled(w, 1.0 / b, Vb[jn])
        _scaled(mw, 1.0 / b, MV[jn])
        j = jn
    if theta is None:
        raise _capi.ConvergenceError("modal: no positive Ritz values (K or M not positive definite on the free DOFs?)")
    if not converged:
        warnings.warn(f"modal: Ritz values not converged to {tol:g} after {steps} Lanczos steps and {restarts} restarts "
                      f"(largest relative residual {float(np.max(resid)):.2e})")
    lam = 1.0 / theta                                              # eigenvalues of (K_hat, M_hat), ascending; theta > 0 checked above
    lam = np.maximum(lam, 0)                                       # freeModes.py:75
    freqs = np.sqrt(lam * scaling) / (2 * np.pi)                   # :76
    jn = Ssel.shape[0]
    S_dev = torch.from_numpy(np.ascontiguousarray(-Ssel.T)).cuda()  # (k, jn), negated: multi_axpy subtracts
    modes_red = torch.zeros((k, n), **f64)
    for i in range(k):
        call("psfem_multi_axpy", jn, n, ptr(Vb), ptr(S_dev[i]), ptr(modes_red[i]), stream_ptr())
    modes_full = fm.expand(modes_red.contiguous(), nvec=k) if k > 1 else fm.expand(modes_red[0]).reshape(1, -1)
    torch.cuda.current_stream().synchronize()
    info = dict(lanczos_steps=steps, restarts=restarts, basis_vectors=m_max_basis, pcg_iterations=op.pcg_iterations,
                direct_factorisation=op.direct, converged=bool(converged), k_max=k_max, m_max=m_max,
                max_residual=float(np.max(resid)) if resid is not None else None)
    return freqs, modes_full.cpu().numpy().reshape(k, -1), constrained, info
