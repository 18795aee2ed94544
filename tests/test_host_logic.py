"""Host-side (index / load) logic of the facade against the reference fixtures -- runs on CPU."""
import numpy as np
import pytest

BC = [(1, [0, 0]), (13, [None, 0]), (7, [5.0, None]), (1, [0, None]), (4, [None, None])]


def test_boundary_edge_forces_merge(psfem, golden):
    P = psfem.Polygones
    d = golden("ref_bc_loads.npz")
    for dr in ("top", "bottom", "left", "right"):
        for mt in ("coarse", "fine"):
            b = P.boundary(dr, [0, None], 4, 2, mesh_type=mt)
            assert [k for k, _ in b] == list(d[f"boundary_{dr}_{mt}"])
            assert all(disp == [0, None] for _, disp in b)
            n_ = 15 if mt == "coarse" else 45
            Fz = np.zeros(2 * n_)
            Fz[3] = 1.0
            out = P.edgeForces(Fz, [3.0, -1e6], dr, 2.0, 4, 2, mesh_type=mt)
            assert out is Fz                                    # mutated in place and returned (Polygones.py:1125-1149)
            assert np.array_equal(out, d[f"edge_{dr}_{mt}"])
    merged = P.merge_boudary_conditions(P.boundary("left", [0, 0], 3), P.boundary("bottom", [0, 0], 3))
    assert [k for k, _ in merged] == list(d["merge"])


def test_dense_bc_paths_bit_exact(psfem, golden):
    """Dense inputs (what the reference's own scripts pass) go through the same index logic."""
    P = psfem.Polygones
    d = golden("ref_bc_loads.npz")
    K = d["bc_K"]
    F = np.arange(K.shape[0], dtype=float)
    K_red, F_red, cd = P.apply_boundary_conditions(K, F, BC)
    assert cd == list(d["bc_constrained"])                      # bc order, duplicates kept
    assert np.array_equal(K_red, d["bc_K_red"]) and np.array_equal(F_red, d["bc_F_red"])
    assert np.array_equal(P.reduce(F, cd), d["bc_reduce_vec"])
    assert np.array_equal(P.reduce(K, cd), d["bc_reduce_mat"])
    assert np.array_equal(P.reconstruct_full_vector(F_red, cd, K.shape[0]), d["bc_full_vec"])
    assert np.array_equal(P.reconstruct_full_vectors(np.stack([F_red, 2 * F_red], axis=1), cd, K.shape[0]), d["bc_full_vecs"])
    assert np.array_equal(P.reconstruct_full_matrix(K_red, cd, K.shape[0]), d["bc_full_mat"])
    import pytest
    with pytest.raises(ValueError, match="either a vector"):
        P.reduce(np.zeros((2, 2, 2)), cd)


def test_displacement_and_element_table(psfem, golden):
    P = psfem.Polygones
    d = golden("ref_bc_loads.npz")
    import psfem_oracle as O
    nodes, conn = O.mesh(4, 2, 4, 2)
    U = np.linspace(-1, 1, 2 * len(nodes))
    assert np.array_equal(P.displacement(nodes, U, scale=2.5), d["displacement"])
    tab = psfem.ElementTable(conn)
    assert len(tab) == len(conn) and list(tab)[:3] == [0, 1, 2]
    assert tab[1] == [int(v) for v in conn[1]] and len(tab[0]) == 3
    assert dict(tab.items())[5] == tab[5]
    import pytest
    with pytest.raises(KeyError):
        tab[len(conn)]
    with pytest.raises(ValueError, match="Integration method"):
        P.global_stiffness_matrix(nodes, tab, 1.0, 0.3, 1.0, integration="bogus")


def test_ramp_load_matches_main_py(psfem, golden):
    d = golden("ref_newmark.npz")
    shape, scale, times = psfem.ramp_load(63, 20)
    assert scale.size == int(d["nm_t3_nbPoints"]) == 10001
    assert scale[0] == 0.0 and scale[-1] == -1e6 * 20 / 63
    T0 = 10001 // 3
    assert scale[T0] == -(3 * 1e6 * times[T0] / 10.0) * 20 / 63


def test_banded_helpers_on_csr_and_dense(psfem):
    """convert_to_banded / calculate_bandwidth (Polygones.py:959-987, :1205-1216) on the structural CSR this package
    returns and on dense input, against the reference's own double loops restated here."""
    import scipy.linalg
    import scipy.sparse as sparse
    import psfem_oracle as O
    P = psfem.Polygones
    nodes, conn = O.mesh(4.0, 1.0, 6, 2)
    K = O.global_stiffness_matrix(nodes, conn, 210e9, 0.3, 0.1)              # structural CSR with explicit zeros
    Kd = K.toarray()
    n = Kd.shape[0]
    lo = up = 0
    for i in range(n):
        for j in range(n):
            if Kd[i, j] != 0:
                if j < i:
                    lo = max(lo, i - j)
                else:
                    up = max(up, j - i)
    for M in (K, Kd):
        assert P.calculate_bandwidth(M) == (lo, up)
        assert P.calculate_bandwidth_optimized(M) == (up, up)
        ab = P.convert_to_banded(M, lo, up)
        ref = np.zeros((lo + up + 1, n))
        for i in range(n):
            for j in range(max(0, i - lo), min(n, i + up + 1)):
                ref[up + i - j, j] = Kd[i, j]
        assert np.array_equal(ab, ref) and np.array_equal(P.convert_to_banded_optimized(M, lo, up), ref)
    # the flow of 2Dexemple.py:29-37: reduce, band, solve_banded
    free = np.arange(6, n)
    Kr = sparse.csr_matrix(K)[free][:, free]
    lo, up = P.calculate_bandwidth(Kr)
    F = np.linspace(1.0, 2.0, len(free))
    U = scipy.linalg.solve_banded((lo, up), P.convert_to_banded(Kr, lo, up), F)
    assert np.abs(Kr @ U - F).max() < 1e-9 * np.abs(F).max()
    Mn, mx = P.normalize_matrix(K)
    assert mx == abs(K).max() and abs(abs(Mn).max() - 1) < 1e-15


def test_t3_host_helpers(psfem):
    P = psfem.Polygones
    nodes = np.array([[0, 0], [1, 0], [2, 0], [2, 1], [1, 1], [0, 1]], dtype=float)
    elements = {0: [0, 1, 4], 1: [4, 1, 2], 2: [2, 3, 4], 3: [4, 5, 0]}                # 2Dexemple.py:11-17
    N = P.shapeFunction(nodes, elements, 1)
    assert np.allclose(N(1, 0), [0, 1, 0]) and np.allclose(N(1, 1), [1, 0, 0]) and np.allclose(N(2, 0), [0, 0, 1])
    assert abs(sum(N(1.3, 0.2)) - 1) < 1e-15                                           # partition of unity
    B = P.strainDisplacement(nodes, elements, 1)
    L = P.shapeFunctionCoeff(nodes, elements, 1)
    assert B.shape == (3, 6) and B[0, 0] == L[0][1] and B[1, 1] == L[0][2] and B[2, 0] == L[0][2] and B[2, 1] == L[0][1]
    assert np.abs(B @ np.array([1, 0, 1, 0, 1, 0.0])).max() < 1e-15                     # a rigid translation has no strain



@pytest.mark.gpu
def test_sparse_symmetry_check_has_allclose_semantics(psfem):
    """diagnostics._is_symmetric == np.allclose(A, A.T, rtol=1e-5, atol=1e-8) (main.py:65-66) without densifying
    (device kernel psfem_csr_symmetry)."""
    import scipy.sparse as sparse
    from psfem_b200.diagnostics import _is_symmetric
    A = sparse.random(60, 60, 0.1, random_state=1, format="csr")
    S = (A + A.T).tocsr()
    cases = [S]
    for (i, j, d) in ((3, 7, 1e-7), (3, 7, 1e-9), (2, 9, 5.0), (11, 12, -3e-6)):
        B = S.tolil()
        B[i, j] = B[i, j] + d
        cases.append(B.tocsr())
    for B in cases:
        D = B.toarray()
        assert _is_symmetric(B) == np.allclose(D, D.T, rtol=1e-5, atol=1e-8)


def test_shape_function_api_vs_reference(psfem, golden):
    """shape_functions / get_shape_derivatives / create_shape_functions (Polygones.py:441-532) against values of the
    unmodified reference's sympy expressions at three reference points (tests/golden/make_golden.py b_matrix_cases)."""
    P = psfem.Polygones
    d = golden("ref_b_matrix.npz")
    pts = d["points"]
    for name, (et, mt) in {"t3": ("tri", "coarse"), "q4": ("quad", "coarse"), "t6": ("tri", "fine"), "q9": ("quad", "fine")}.items():
        ref = d[f"{name}_shape_vals"]                               # (points, [N, dN/dxi, dN/deta], npe)
        groups = P.create_shape_functions(et, mt)
        for k, (xi, eta) in enumerate(pts):
            got = np.array([[float(f(xi, eta)) for f in grp] for grp in groups])
            assert np.abs(got - ref[k]).max() <= 1e-14
        # vectorised evaluation, like the lambdified functions of the reference
        assert np.allclose(groups[0][0](pts[:, 0], pts[:, 1]), ref[:, 0, 0], rtol=0, atol=1e-14)
        N, xi_s, eta_s = P.shape_functions(et, mt)                  # sympy expressions in the symbols xi, eta
        dxi, deta = P.get_shape_derivatives(et, mt)
        assert len(N) == ref.shape[2] and str(xi_s) == "xi" and str(eta_s) == "eta"
        for k, (xi, eta) in enumerate(pts):
            sub = {"xi": xi, "eta": eta}                            # the reference substitutes by NAME (Polygones.py:676)
            got = np.array([[float(f.subs(sub)) for f in grp] for grp in (N, dxi, deta)])
            assert np.abs(got - ref[k]).max() <= 1e-14
        assert abs(sum(float(n.subs({"xi": 0.3, "eta": 0.2})) for n in N) - 1.0) < 1e-14     # partition of unity


def test_device_backed_csr_never_goes_stale(psfem):
    """The matrices the assembly returns remember their device copy; any host-side edit must invalidate it
    (a penalty `K[i, i] = 1e20` used to be ignored by the drivers).  Pure host logic."""
    import pytest
    import scipy.sparse as sparse
    from psfem_b200.Polygones import DeviceBackedCSR

    def fresh():
        A = DeviceBackedCSR((np.arange(1.0, 8.0), np.array([0, 1, 0, 1, 2, 1, 2]), np.array([0, 2, 5, 7])), shape=(3, 3))
        A.has_canonical_format = True
        return A._psfem_attach("device copy")
    A = fresh()
    assert A._psfem_valid() and isinstance(A, sparse.csr_matrix) and sparse.issparse(A)
    assert np.array_equal(A @ np.ones(3), [3.0, 12.0, 13.0]) and A._psfem_valid()        # reading keeps the cache
    assert A[np.ix_([0, 2], [0, 2])].shape == (2, 2) and (A * 2)._psfem_dev is None and A._psfem_valid()
    with pytest.raises(ValueError, match="read-only"):
        A.data[0] = 5.0                                                                   # direct writes fail loudly
    A[0, 0] = 1e20                                                                        # the penalty idiom
    assert not A._psfem_valid() and A._psfem_dev is None and A[0, 0] == 1e20
    for edit in (lambda M: M.setdiag(3.0), lambda M: M.__imul__(2.0), lambda M: M.__itruediv__(2.0),
                 lambda M: M.eliminate_zeros(), lambda M: M.sort_indices(), lambda M: M.sum_duplicates()):
        A = fresh()
        edit(A)
        assert not A._psfem_valid() and A.data.flags.writeable
    A = fresh()
    A.data.flags.writeable = True                                                         # the caller unlocks the values
    A.data[1] = 7.0
    assert not A._psfem_valid()
    A = fresh()
    A.data = A.data.copy()                                                                # or swaps the array
    assert not A._psfem_valid()
    A = fresh()
    assert A.copy()._psfem_dev is None and A.copy().data.flags.writeable and A._psfem_valid()


def test_continuation_from_true_residual_control_flow(psfem, monkeypatch):
    """device.continue_from_true_residual (the host loop around the classic PCG entry points): stops when the evaluated
    residual passes, when a cycle does not halve it (rounding floor), after `polish` cycles, and never runs for an
    unconverged solve or with a negative `polish`; iterations of the cycles are added and reported."""
    from psfem_b200.device import continue_from_true_residual

    def scenario(residuals, polish, converged=True, rc_seq=None):
        res, runs = list(residuals), []

        def run():
            runs.append(1)
            return (rc_seq.pop(0) if rc_seq else 0), 7

        out = dict(iterations=100, relres=9e-9, converged=converged)
        continue_from_true_residual(run, lambda: res.pop(0), 1e-8, out, polish=polish)
        return out, len(runs)

    out, runs = scenario([5e-9], 3)                              # already below rtol: evaluated, nothing run
    assert runs == 0 and out["relres_true"] == 5e-9 and out["polish_cycles"] == 0 and out["iterations"] == 100
    out, runs = scenario([4e-7, 2e-8, 6e-9], 3)                  # two cycles bring it below rtol
    assert runs == 2 and out["relres_true"] == 6e-9 and out["polish_cycles"] == 2 and out["iterations"] == 114 and out["polish_iterations"] == 14
    out, runs = scenario([4e-7, 2e-8, 1.5e-8, 1.4e-8], 5)        # floor: the second cycle does not halve the residual
    assert runs == 2 and out["relres_true"] == 1.5e-8 and out["polish_cycles"] == 2
    out, runs = scenario([4e-7, 1e-7, 4e-8, 1.5e-8, 1e-9], 3)    # cycle limit
    assert runs == 3 and out["relres_true"] == 1.5e-8 and out["polish_cycles"] == 3
    out, runs = scenario([4e-7], 0)                              # evaluate only
    assert runs == 0 and out["relres_true"] == 4e-7
    out, runs = scenario([4e-7], -1)                             # switched off
    assert runs == 0 and "relres_true" not in out
    out, runs = scenario([4e-7], 3, converged=False)             # an unconverged solve is left alone
    assert runs == 0 and "relres_true" not in out
    out, runs = scenario([4e-7, 3e-7], 3, rc_seq=[psfem._capi.ENOCONV])   # a cycle that fails ends the loop; the solve stays converged
    assert runs == 1 and out["converged"] and out["polish_cycles"] == 1
    monkeypatch.setenv("PSFEM_PCG_POLISH", "1")                  # the default comes from the environment
    res = [4e-7, 1e-7, 1e-9]
    out = dict(iterations=10, relres=9e-9, converged=True)
    continue_from_true_residual(lambda: (0, 1), lambda: res.pop(0), 1e-8, out)
    assert out["polish_cycles"] == 1 and out["relres_true"] == 1e-7
