"""Host-side (index / load) logic of the facade against the reference fixtures -- runs on CPU."""
import numpy as np

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



def test_sparse_symmetry_check_has_allclose_semantics(psfem):
    """diagnostics._is_symmetric == np.allclose(A, A.T, rtol=1e-5, atol=1e-8) (main.py:65-66) without densifying."""
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
