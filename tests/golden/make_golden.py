"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What it does
  * imports /root/reference/Polygones.py with empty matplotlib stub modules (matplotlib is not
    installed) and memoises the two symbolic helpers (`shape_functions`,
    `get_shape_derivatives`) from the outside -- output is bit-identical, it is only faster;
  * converts the reference's three reproducible golden stiffness files to CSR .npz;
  * runs the reference functions on small meshes of all four element types (structured,
    jittered and hand-written-dict meshes) and stores dense K, M, BC indexing, loads;
  * executes the reference driver scripts (`TestBeam/Statictest.py` arithmetic, `main.py`,
    `freeModes.py`) with `runpy` under a MagicMock matplotlib, on parameter-edited temp copies
    where BASELINE.json's configs ask for another element type (the scripts' own usage model:
    "copy and paste mesh parameters"), and stores U, energies and frequencies.
Nothing here is imported by the product.
"""
import functools
import io
import os
import re
import runpy
import shutil
import sys
import tempfile
import types
from unittest.mock import MagicMock

import numpy as np
import scipy.sparse as sparse

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.widgets"):
        sys.modules[n] = MagicMock(name=n)
    sys.path.insert(0, REF)
    import Polygones  # noqa
    Polygones.shape_functions = functools.lru_cache(None)(Polygones.shape_functions)
    Polygones.get_shape_derivatives = functools.lru_cache(None)(Polygones.get_shape_derivatives)
    return Polygones


def conn_of(elements):
    return np.array([elements[k] for k in sorted(elements)], dtype=np.int64)


def golden_files():
    """The reference's own golden stiffness files -> CSR of their non-zeros (values exact as parsed)."""
    import pandas as pd
    K1 = pd.read_csv(f"{REF}/TestBeam/stiffness_matrix.txt", header=None).to_numpy()
    K2 = pd.read_csv(f"{REF}/TestBeam/stiffness_matrix.csv", header=None).to_numpy()
    K3 = pd.read_csv(f"{REF}/comparison/K.csv", header=0).to_numpy()
    out = {}
    for name, K in (("beam738", K1), ("sparse210", K2), ("square882", K3)):
        S = sparse.csr_matrix(K)
        out[f"{name}_indptr"] = S.indptr
        out[f"{name}_indices"] = S.indices
        out[f"{name}_data"] = S.data
        out[f"{name}_shape"] = np.array(K.shape)
    np.savez_compressed(f"{OUT}/ref_golden_files.npz", **out)
    print("golden files:", {k: v.shape for k, v in out.items() if k.endswith("data")})


def small_meshes(P):
    """K, M (dense) of the reference on small meshes of every element type."""
    E, nu, t, rho = 210e9, 0.3, 0.1, 7850.0
    rng = np.random.default_rng(0)
    out = {}
    cases = {
        "t3": dict(L=3, l=2.0, N=4, M=3, offsetX=0.5, offsetY=1, element_type="tri", mesh_type="coarse"),
        "q4": dict(L=5.0, l=3, N=5, M=3, offsetX=0, offsetY=1.5, element_type="quad", mesh_type="coarse"),
        "t6": dict(L=3, l=2, N=3, M=2, offsetX=0, offsetY=1, element_type="tri", mesh_type="fine"),
        "q9": dict(L=2.0, l=1.5, N=2, M=2, offsetX=0.25, offsetY=0, element_type="quad", mesh_type="fine"),
        "t3sq": dict(L=1, l=1, N=3, element_type="tri", mesh_type="coarse"),     # M=0 -> M=N
    }
    for name, kw in cases.items():
        for variant in ("reg", "jit"):
            nodes, elements = P.mesh(**kw)
            if variant == "jit":
                n_, m_ = (kw["N"] + 1, (kw.get("M", 0) or kw["N"]) + 1) if kw["mesh_type"] == "coarse" else \
                         (2 * kw["N"] + 1, 2 * (kw.get("M", 0) or kw["N"]) + 1)
                hx = kw["L"] / (n_ - 1)
                hy = kw["l"] / (m_ - 1)
                nodes = nodes + rng.uniform(-0.12, 0.12, nodes.shape) * np.array([hx, hy])
            et, mt = kw["element_type"], kw["mesh_type"]
            K = P.global_stiffness_matrix(nodes, elements, E, nu, t, et, mt)
            M = P.global_mass_matrix(nodes, elements, rho, t, et, mt)
            key = f"{name}_{variant}"
            out[f"{key}_nodes"] = nodes
            out[f"{key}_conn"] = conn_of(elements)
            out[f"{key}_K"] = K
            out[f"{key}_M"] = M
            out[f"{key}_Ke0"] = P.element_stiffness_matrix(nodes, elements, 0, E, nu, t, et, mt)
            out[f"{key}_Me1"] = P.element_mass_matrix(nodes, elements, 1, rho, t, et, mt)
            print(key, K.shape, "nnz dense", np.count_nonzero(K))
    # hand-written dict mesh of 2Dexemple.py:11-21 (incl. its material)
    nodes = np.array([[0, 0], [1, 0], [2, 0], [2, 1], [1, 1], [0, 1]], dtype=float)
    elements = {0: [0, 1, 4], 1: [4, 1, 2], 2: [2, 3, 4], 3: [4, 5, 0]}
    out["ex2d_nodes"] = nodes
    out["ex2d_conn"] = conn_of(elements)
    out["ex2d_K"] = P.global_stiffness_matrix(nodes, elements, 200e9, 0.3, 0.1)
    # a clockwise T3 (negative signed area -> -Ke, -Me)
    elements_cw = {0: [0, 4, 1], 1: [4, 1, 2]}
    out["cw_conn"] = conn_of(elements_cw)
    out["cw_K"] = P.global_stiffness_matrix(nodes, elements_cw, 200e9, 0.3, 0.1)
    out["cw_M"] = P.global_mass_matrix(nodes, elements_cw, 7850.0, 0.1)
    np.savez_compressed(f"{OUT}/ref_small_meshes.npz", **out)


def bc_and_loads(P):
    """apply_boundary_conditions / reduce / reconstruct / boundary / edgeForces outputs."""
    out = {}
    nodes, elements = P.mesh(4, 2, 4, 2)
    K = P.global_stiffness_matrix(nodes, elements, 210e9, 0.3, 0.1)
    F = np.arange(2 * len(nodes), dtype=float)
    bc = [(1, [0, 0]), (13, [None, 0]), (7, [5.0, None]), (1, [0, None]), (4, [None, None])]
    K_red, F_red, cd = P.apply_boundary_conditions(K, F, bc)
    out["bc_K"] = K
    out["bc_K_red"] = K_red
    out["bc_F_red"] = F_red
    out["bc_constrained"] = np.array(cd)
    out["bc_reduce_vec"] = P.reduce(F, cd)
    out["bc_reduce_mat"] = P.reduce(K, cd)
    out["bc_full_vec"] = P.reconstruct_full_vector(F_red, cd, 2 * len(nodes))
    out["bc_full_vecs"] = P.reconstruct_full_vectors(np.stack([F_red, 2 * F_red], axis=1), cd, 2 * len(nodes))
    out["bc_full_mat"] = P.reconstruct_full_matrix(K_red, cd, 2 * len(nodes))
    for d in ("top", "bottom", "left", "right"):
        for mt in ("coarse", "fine"):
            b = P.boundary(d, [0, None], 4, 2, mesh_type=mt)
            out[f"boundary_{d}_{mt}"] = np.array([k for k, _ in b])
            n_ = (5 * 3) if mt == "coarse" else (9 * 5)
            Fz = np.zeros(2 * n_)
            Fz[3] = 1.0
            out[f"edge_{d}_{mt}"] = P.edgeForces(Fz, [3.0, -1e6], d, 2.0, 4, 2, mesh_type=mt)
    out["merge"] = np.array([k for k, _ in P.merge_boudary_conditions(P.boundary("left", [0, 0], 3), P.boundary("bottom", [0, 0], 3))])
    U = np.linspace(-1, 1, 2 * len(nodes))
    out["displacement"] = P.displacement(nodes, U, scale=2.5)
    np.savez_compressed(f"{OUT}/ref_bc_loads.npz", **out)


def static_cases(P):
    """TestBeam/Statictest.py:40-78 arithmetic for G=2 (tri, quad) + README cantilever C1."""
    out = {}
    E, nu, t, l, L, Pload = 210e9, 0.3, 0.1, 2, 20, 1e4
    for elt in ("tri", "quad"):
        for G in (2, 4):
            N = 10 * G
            nodes, elements = P.mesh(L, l, N, G, element_type=elt, mesh_type="coarse", offsetY=l / 2)
            K = P.global_stiffness_matrix(nodes, elements, E, nu, t, elt, "coarse", integration="gauss")
            F = np.zeros(2 * len(nodes))
            for i in range(len(nodes)):
                F[2 * i + 1] = -Pload * L / len(nodes)
            node1, node2 = int(G / 2), int((len(nodes) - 1) - G / 2)
            bc = [(node1, [0, 0]), (node2, [None, 0])]
            K_red, F_red, cd = P.apply_boundary_conditions(K, F, bc)
            U_red = np.linalg.solve(K_red, F_red)
            U = np.zeros(2 * len(nodes))
            U[np.setdiff1d(np.arange(len(U)), cd)] = U_red
            out[f"static_{elt}_G{G}_U"] = U
            out[f"static_{elt}_G{G}_maxU"] = np.max(np.abs(U))
            print("static", elt, G, np.max(np.abs(U)))
    # C1: README cantilever (README.md usage; h undefined there -> 0.1)
    nodes, elements = P.mesh(L=10, l=1, N=40)
    K = P.global_stiffness_matrix(nodes, elements, 210e9, 0.3, 0.1)
    M = P.global_mass_matrix(nodes, elements, 7850, 0.1)
    F = np.zeros(2 * len(nodes))
    F = P.edgeForces(F, [0, -1e6], "right", 1, 40)
    bc = P.boundary("left", [0, 0], 40)
    K_red, F_red, cd = P.apply_boundary_conditions(K, F, bc)
    U_red = np.linalg.solve(K_red, F_red)
    U = P.reconstruct_full_vector(U_red, cd, 2 * len(nodes))
    out["c1_U"] = U
    out["c1_constrained"] = np.array(cd)
    out["c1_F"] = F
    Ks = sparse.csr_matrix(K)
    Ms = sparse.csr_matrix(M)
    out["c1_K_nnz_dense"] = Ks.nnz
    out["c1_K_data"], out["c1_K_indices"], out["c1_K_indptr"] = Ks.data, Ks.indices, Ks.indptr
    out["c1_M_data"], out["c1_M_indices"], out["c1_M_indptr"] = Ms.data, Ms.indices, Ms.indptr
    print("c1 tip", U[2 * (40 * 41 + 20) + 1], np.max(np.abs(U)))
    np.savez_compressed(f"{OUT}/ref_static.npz", **out)


def _edit_params(src, **params):
    for k, v in params.items():
        src, n = re.subn(rf"^{k}\s*=\s*[^#\n]*", f"{k} = {v!r} ", src, count=1, flags=re.M)
        assert n == 1, k
    return src


def _run_script(P, name, params, compute_params):
    """Write K/M with the ComputeMatrices.py:30-37 recipe for `compute_params`, then runpy a
    (parameter-edited) temp copy of the reference script `name` in that directory."""
    import pandas as pd
    tmp = tempfile.mkdtemp(prefix="psfem_golden_")
    cwd = os.getcwd()
    try:
        cp = compute_params
        nodes, elements = P.mesh(cp["L"], cp["l"], cp["N"], cp["G"], element_type=cp["elt"], mesh_type=cp["met"], offsetY=cp["l"] / 2)
        K = P.global_stiffness_matrix(nodes, elements, cp["E"], cp["nu"], cp["t"], cp["elt"], cp["met"], integration="gauss")
        M = P.global_mass_matrix(nodes, elements, cp["pho"], cp["t"], cp["elt"], cp["met"], integration="gauss")
        pd.DataFrame(K).to_csv(f"{tmp}/stiffness_matrix.txt", sep="\t", index=False, header=False)
        pd.DataFrame(M).to_csv(f"{tmp}/mass_matrix.txt", sep="\t", index=False, header=False)
        src = open(f"{REF}/{name}").read()
        src = _edit_params(src, **params)
        path = f"{tmp}/{name}"
        open(path, "w").write(src)
        os.chdir(tmp)
        stdout = sys.stdout
        sys.stdout = io.StringIO()
        try:
            g = runpy.run_path(path, run_name="__main__")
        finally:
            sys.stdout = stdout
        return g, K, M
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


def newmark_cases(P):
    out = {}
    base = dict(L=20, l=2, E=210e9, nu=0.3, pho=7850)
    for tag, elt, met in (("t3", "tri", "coarse"), ("q9", "quad", "fine")):
        cp = dict(base, N=20, G=2, t=0.1, elt=elt, met=met)
        g, K, M = _run_script(P, "main.py", dict(elt=elt, met=met), cp)
        U, V, A = g["U_reduced"], g["V_reduced"], g["A_reduced"]
        sel = np.arange(0, U.shape[1], 250)
        out[f"nm_{tag}_steps"] = sel
        out[f"nm_{tag}_U"] = U[:, sel]
        out[f"nm_{tag}_V"] = V[:, sel]
        out[f"nm_{tag}_A"] = A[:, sel]
        out[f"nm_{tag}_kin"] = g["kinetic_energy"]
        out[f"nm_{tag}_pot"] = g["potential_energy"]
        out[f"nm_{tag}_tot"] = g["total_energy"]
        out[f"nm_{tag}_maxU_t"] = np.max(np.abs(U), axis=0)
        out[f"nm_{tag}_constrained"] = np.array(g["constrained_dofs"])
        out[f"nm_{tag}_dt"] = g["dt"]
        out[f"nm_{tag}_nbPoints"] = g["nbPoints"]
        print("newmark", tag, U.shape, "E[10000]", g["total_energy"][10000], "max|U|", np.max(np.abs(U)))
    np.savez_compressed(f"{OUT}/ref_newmark.npz", **out)


def explicit_case(P):
    """main.py with `use_newmark = False` (the modified explicit Euler branch, :334-372; dead in the shipped script).
    dt = 1e-3 is far above its stability limit: the series explodes by a factor ~370 per step until the stabiliser
    (:365-370) takes over at step 11, and rounding differences take over the digits after ~45 steps; the first 41
    columns pin the recurrence and the stabiliser (an independent restatement reproduces them to 1e-14)."""
    base = dict(L=20, l=2, E=210e9, nu=0.3, pho=7850)
    cp = dict(base, N=20, G=2, t=0.1, elt="tri", met="coarse")
    g, K, M = _run_script(P, "main.py", dict(elt="tri", met="coarse", use_newmark=False), cp)
    n = 41
    out = dict(ex_U=g["U_reduced"][:, :n], ex_V=g["V_reduced"][:, :n], ex_A=g["A_reduced"][:, :n],
               ex_constrained=np.array(g["constrained_dofs"]), ex_dt=g["dt"], ex_F=g["ForcesVectors_reduced"][:, :n],
               ex_kin=g["kinetic_energy"][:n], ex_pot=g["potential_energy"][:n])
    print("explicit", g["U_reduced"].shape, "max|U[:, :41]|", np.max(np.abs(out["ex_U"])))
    np.savez_compressed(f"{OUT}/ref_explicit.npz", **out)


def modal_cases(P):
    out = {}
    base = dict(L=20, l=2, E=210e9, nu=0.3, pho=7850, t=0.01)
    for tag, N, G, elt, met in (("t3_80x8", 80, 8, "tri", "coarse"), ("t6_20x2", 20, 2, "tri", "fine"),
                                ("t6_40x4", 40, 4, "tri", "fine")):
        cp = dict(base, N=N, G=G, elt=elt, met=met)
        # the script ends in animate_plate_oscillation() calls (UI); under the matplotlib mock they
        # cannot unpack plt.subplots(), so the harness no-ops that one UI function.
        P.animate_plate_oscillation = lambda *a, **k: None
        g, K, M = _run_script(P, "freeModes.py", dict(N=N, G=G, elt=elt, met=met), cp)
        f = g["frequencies_dense"]
        out[f"modal_{tag}_freq"] = f[:40]
        out[f"modal_{tag}_fmax"] = f[-1]     # conditioning of the reference's own dense eigh
        out[f"modal_{tag}_modes"] = g["eigenvectors_reconstructed"][:20]
        out[f"modal_{tag}_constrained"] = np.array(g["constrained_dofs"])
        print("modal", tag, f[:5])
    np.savez_compressed(f"{OUT}/ref_modal.npz", **out)


def body_force_cases(P):
    """Polygones.body_forces (Polygones.py:1151-1200; no caller in the reference) on the small meshes."""
    t = 0.1
    rng = np.random.default_rng(1)
    out = {}
    cases = {
        "t3": dict(L=3, l=2.0, N=4, M=3, offsetX=0.5, offsetY=1, element_type="tri", mesh_type="coarse"),
        "q4": dict(L=5.0, l=3, N=5, M=3, offsetX=0, offsetY=1.5, element_type="quad", mesh_type="coarse"),
        "t6": dict(L=3, l=2, N=3, M=2, offsetX=0, offsetY=1, element_type="tri", mesh_type="fine"),
        "q9": dict(L=2.0, l=1.5, N=2, M=2, offsetX=0.25, offsetY=0, element_type="quad", mesh_type="fine"),
    }
    for name, kw in cases.items():
        nodes, elements = P.mesh(**kw)
        n_, m_ = (kw["N"] + 1, kw["M"] + 1) if kw["mesh_type"] == "coarse" else (2 * kw["N"] + 1, 2 * kw["M"] + 1)
        nodes = nodes + rng.uniform(-0.1, 0.1, nodes.shape) * np.array([kw["L"] / (n_ - 1), kw["l"] / (m_ - 1)])
        F0 = rng.standard_normal(2 * len(nodes))
        F = P.body_forces(nodes, elements, F0.copy(), (1.5, -9.81), t, kw["element_type"], kw["mesh_type"])
        out[f"{name}_nodes"], out[f"{name}_conn"], out[f"{name}_F0"], out[f"{name}_F"] = nodes, conn_of(elements), F0, F
        print("body_forces", name, np.abs(F - F0).max())
    np.savez_compressed(f"{OUT}/ref_body_forces.npz", **out)


def b_matrix_cases(P):
    """Polygones.compute_B_matrix_at_point (Polygones.py:657-708) on jittered small meshes of all four element kinds:
    (B, J, detJ) of two elements at three reference points each, plus the shape functions / derivatives evaluated at
    those points (shape_functions :441, get_shape_derivatives :504, create_shape_functions :520)."""
    rng = np.random.default_rng(2)
    out = {}
    cases = {
        "t3": dict(L=3, l=2.0, N=4, M=3, offsetX=0.5, offsetY=1, element_type="tri", mesh_type="coarse"),
        "q4": dict(L=5.0, l=3, N=5, M=3, offsetX=0, offsetY=1.5, element_type="quad", mesh_type="coarse"),
        "t6": dict(L=3, l=2, N=3, M=2, offsetX=0, offsetY=1, element_type="tri", mesh_type="fine"),
        "q9": dict(L=2.0, l=1.5, N=2, M=2, offsetX=0.25, offsetY=0, element_type="quad", mesh_type="fine"),
    }
    pts = [(0.2, 0.3), (-0.55, 0.1), (1 / 6, 2 / 3)]
    for name, kw in cases.items():
        nodes, elements = P.mesh(**kw)
        n_, m_ = (kw["N"] + 1, kw["M"] + 1) if kw["mesh_type"] == "coarse" else (2 * kw["N"] + 1, 2 * kw["M"] + 1)
        nodes = nodes + rng.uniform(-0.1, 0.1, nodes.shape) * np.array([kw["L"] / (n_ - 1), kw["l"] / (m_ - 1)])
        et, mt = kw["element_type"], kw["mesh_type"]
        Bs, Js, ds = [], [], []
        for e in (0, len(elements) - 1):
            for xi, eta in pts:
                B, J, detJ = P.compute_B_matrix_at_point(nodes, elements, e, xi, eta, et, mt)
                Bs.append(B); Js.append(J); ds.append(detJ)
        N, sx, se = P.shape_functions(et, mt)
        d1, d2 = P.get_shape_derivatives(et, mt)
        vals = np.array([[[float(f.subs({sx: xi, se: eta})) for f in grp] for grp in (N, d1, d2)] for xi, eta in pts])
        out[f"{name}_nodes"], out[f"{name}_conn"] = nodes, conn_of(elements)
        out[f"{name}_B"], out[f"{name}_J"], out[f"{name}_detJ"] = np.array(Bs), np.array(Js), np.array(ds)
        out[f"{name}_shape_vals"] = vals                     # (points, [N, dN/dxi, dN/deta], npe)
        print("b_matrix", name, np.array(Bs).shape, ds[:2])
    out["points"] = np.array(pts)
    np.savez_compressed(f"{OUT}/ref_b_matrix.npz", **out)


if __name__ == "__main__":
    P = import_reference()
    what = sys.argv[1:] or ["files", "small", "bc", "static", "newmark", "explicit", "modal", "body", "bmat"]
    if "bmat" in what:
        b_matrix_cases(P)
    if "files" in what:
        golden_files()
    if "small" in what:
        small_meshes(P)
    if "bc" in what:
        bc_and_loads(P)
    if "static" in what:
        static_cases(P)
    if "newmark" in what:
        newmark_cases(P)
    if "explicit" in what:
        explicit_case(P)
    if "modal" in what:
        modal_cases(P)
    if "body" in what:
        body_force_cases(P)
