#!/usr/bin/env python
"""bench.py -- headline benchmark of the plane-stress FEM hot path on B200.

Metric (BASELINE.json): "Q4 elements assembled/s + CG DOF*iter/s, 33.6M-DOF plate, 1/2/4/8 B200".
Workload (configs[3] at N=1): synthetic 4096x4096 Q4 plate (16 785 409 nodes, 33 570 818 DOF),
left edge clamped, right edge loaded -- mesh(4.096, 4.096, 4096, 4096, element_type='quad').
For N>1 every rank owns a strip of 4096 cell columns of a (4096*N)x4096 plate (weak scaling;
N=4 is BASELINE's configs[4], the 16384x4096 strip), halo columns exchanged over NCCL each SpMV,
CG dot products all-reduced.

One STEP = one numeric assembly of K over all elements of the rank (one launch of the lattice kernel) +
`--cg-iters` Jacobi-PCG iterations on the clamped system (three kernels each; for N>1 the reductions and the
halo exchange are fused into those kernels over peer memory, PSFEM_DIST_TRANSPORT=nccl selects NCCL instead).  `value` = CG DOF*iterations/s (whole job); the assembly rate is
in the `assembly` object.  Device timing with CUDA events on the launching stream, max over ranks.
All inputs (> 7 GB of CSR per rank) are far larger than the 126 MB L2, so no explicit L2 flush.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--cells 4096]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Q4 elements assembled/s + CG DOF*iter/s, 33.6M-DOF plate, 1/2/4/8 B200"
UNIT = "DOF*iter/s"
E, NU, T, RHO = 210e9, 0.3, 0.1, 7850.0


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi in loop mode.  The process needs about a second to produce its first row, so it is started (and
    waited for) BEFORE the warm-up; rows are time-stamped and only those taken while the benchmark loop (warm-up +
    timed steps, the same load) was running are reported."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t_load0 = None

    def start(self, wait_s=5.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < wait_s:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_load(self):
        self.t_load0 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        time.sleep(0.06)                      # let the row that covers the end of the loop arrive
        self.proc.terminate()
        rows = [r for t, r in self.rows if self.t_load0 is None or t >= self.t_load0]
        rows = rows or [r for _, r in self.rows[-1:]]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": "warm-up + timed steps (same load)",
                "window_s": t_end - (self.t_load0 or t_end)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (NumPy/SciPy restatement of the reference) on host cores
# ------------------------------------------------------------------------------------------------
def cpu_run(steps, warmup, asm_cells=512, cg_cells=1024, cg_iters=20):
    """Bounded sample of the same workload on the host: Q4 assembly (Ke + coo->csr, structural
    pattern) of an asm_cells^2 plate and cg_iters Jacobi-PCG iterations on a clamped cg_cells^2 plate."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import psfem_oracle as O
    cores = os.cpu_count() or 1
    h = 1e-3
    nodes, conn = O.mesh(asm_cells * h, asm_cells * h, asm_cells, asm_cells, element_type="quad")
    n2, c2 = O.mesh(cg_cells * h, cg_cells * h, cg_cells, cg_cells, element_type="quad")
    K2 = O.global_stiffness_matrix(n2, c2, E, NU, T, "quad")
    F = O.edge_forces(np.zeros(K2.shape[0]), [0, -1e6], "right", cg_cells * h, cg_cells)
    K_red, F_red, _ = O.apply_boundary_conditions(K2, F, O.boundary("left", [0, 0], cg_cells))
    n_free = K_red.shape[0]
    t_asm, t_cg, used = [], [], 1
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        Ke = O.element_stiffness_threads(nodes, conn, E, NU, T, "quad", "coarse", cores)
        O.assemble(Ke, conn, len(nodes))
        ta = time.perf_counter() - t0
        best = None
        for nt in sorted({1, cores}):
            _, _, sec = O.jacobi_pcg_threads(K_red, F_red, cg_iters, nt)
            if best is None or sec < best[0]:
                best = (sec, nt)
        if s >= warmup:
            t_asm.append(ta)
            t_cg.append(best[0])
            used = best[1]
    asm_rate = asm_cells * asm_cells * len(t_asm) / sum(t_asm)
    cg_rate = n_free * cg_iters * len(t_cg) / sum(t_cg)
    sample = (f"per step: Q4 assembly (NumPy Ke on {cores} threads + scipy coo->csr) of a {asm_cells}x{asm_cells} plate "
              f"and {cg_iters} Jacobi-PCG iterations (SciPy CSR mat-vec, best of 1/{cores} threads) on a clamped "
              f"{cg_cells}x{cg_cells} plate ({n_free} DOF)")
    return dict(cg_rate=cg_rate, asm_rate=asm_rate, cores=max(used, cores), cg_threads=used, sample=sample,
                ms_per_step=1e3 * (sum(t_asm) + sum(t_cg)) / len(t_cg))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_run(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["cg_rate"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus, args.cells), "sample": r["sample"]},
        "assembly": {"value": r["asm_rate"], "unit": "elements/s"},
        "cpu_baseline": {"value": r["cg_rate"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                         "assembly_elements_per_s": r["asm_rate"]},
        "e2e": {"value": r["cg_rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the unmodified reference (pure Python, dense (2n)^2 storage, ~2 Q4 elements/s) cannot run this workload; "
                "this arm times oracle/psfem_oracle.py, the NumPy/SciPy restatement pinned to the reference's goldens",
    }
    print(json.dumps(line), flush=True)


def workload_name(n_gpus, cells, scaling="weak", total_cols=0):
    if total_cols:
        return (f"synthetic {total_cols}x{cells} Q4 strip, left edge clamped, right edge loaded: numeric assembly of K + "
                f"Jacobi-PCG iterations; split into {n_gpus} strips of ~{total_cols // n_gpus} cell columns")
    if scaling == "strong":
        return (f"synthetic {cells}x{cells} Q4 plate, left edge clamped, right edge loaded: numeric assembly of K + "
                f"Jacobi-PCG iterations; the plate is split into {n_gpus} strips of ~{cells // n_gpus} cell columns")
    return (f"synthetic {cells * n_gpus}x{cells} Q4 plate, left edge clamped, right edge loaded: numeric assembly of K + "
            f"Jacobi-PCG iterations; {cells} cell columns per GPU (strip partition)")


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=4096, help="cells per side of the per-GPU plate")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: `--cells` cell columns PER GPU (default, the driver's contract); strong: the cells x cells plate split over all GPUs")
    ap.add_argument("--total-cols", type=int, default=0,
                    help="cell columns of the WHOLE plate, split over all GPUs (rows = --cells): --total-cols 16384 is BASELINE's "
                         "16384x4096 strip (configs[4]) on any number of GPUs; reported as strong scaling")
    ap.add_argument("--cg-iters", type=int, default=25, help="PCG iterations per step")
    ap.add_argument("--e2e-iters", type=int, default=200, help="PCG iterations of one end-to-end static_solve call")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--full-solve", action="store_true", help="also solve the plate to 1e-8 and report iterations/seconds "
                    "(on by default for the plain single-GPU run: 24 500 iterations, about 22 s)")
    ap.add_argument("--no-full-solve", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import psfem_b200
    from psfem_b200 import _capi
    from psfem_b200.dist import StripProblem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, peak_src = peaks()

    cells = args.cells
    if args.total_cols:
        args.scaling = "strong"
    cols_total = args.total_cols or (cells * world if args.scaling == "weak" else cells)
    prob = StripProblem(L=cols_total * 1e-3, l=cells * 1e-3, N=cols_total, M=cells, rank=rank, world=world,
                        E=E, nu=NU, t=T)
    prob.setup()                       # mesh strip, symbolic pattern, first assembly, BC reduction, load vector
    n_free_global = prob.n_free_global
    nel_local = prob.nel_owned
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()                                        # nvidia-smi is up before the first warm-up step
    asm_ms, cg_ms, spmv_ms = [], [], []
    launches = 0
    barrier()
    sampler.mark_load()
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            barrier()
            t_wall0 = time.perf_counter()
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        prob.assemble()                                    # numeric assembly of K (one marching-warp lattice kernel)
        e1.record()
        prob.cg_iterations(args.cg_iters)                  # 3 kernels per iteration (+ halo / all-reduce when world > 1)
        e2.record()
        prob.spmv_only(args.cg_iters)                      # the dominant kernel alone, for its roofline line
        e3.record()
        torch.cuda.synchronize()
        if s >= args.warmup:
            asm_ms.append(e0.elapsed_time(e1))
            cg_ms.append(e1.elapsed_time(e2))
            spmv_ms.append(e2.elapsed_time(e3))
            launches += 1 + 3 * args.cg_iters + args.cg_iters   # assembly + PCG kernels + the SpMV-only leg
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_asm = maxr(sum(asm_ms)) / 1e3
    t_cg = maxr(sum(cg_ms)) / 1e3
    t_spmv = maxr(sum(spmv_ms)) / 1e3
    K_ = args.steps
    cg_rate = n_free_global * args.cg_iters * K_ / t_cg
    asm_rate = prob.nel_global * K_ / t_asm

    # roofline of the dominant kernel (SpMV+dot of the PCG iteration), per rank
    nnz, n = prob.K_red.nnz, prob.n_owned_free
    csr_bytes = 12 * nnz + 4 * (n + 1) + 16 * n                          # CSR SpMV
    nbp = prob.K_red.plan()[3]
    node_block = nbp is not None and not os.environ.get("PSFEM_NO_NODEBLOCK")
    idx_bytes = 2 if (node_block and nbp[5] is not None) else 4
    # node-block SpMV: 8 B value + one node id (16-bit offset or 32-bit) per 2x2 block, node_ptr, x, y
    spmv_bytes = (8 * nnz + idx_bytes * (nnz // 4) + 4 * (n // 2 + 1) + 16 * n) if node_block else csr_bytes
    sym_lat = bool(getattr(prob, "sym_lattice", False)) and not os.environ.get("PSFEM_NO_SYMLATTICE")
    if sym_lat:
        # symmetric-lattice storage: 19 doubles of matrix per node (diagonal + 4 forward 2x2 blocks), x, y
        spmv_bytes = (19 * 8 + 16 + 16) * (n // 2)
    spmv_t = t_spmv / (K_ * args.cg_iters)
    spmv_gbs = spmv_bytes / spmv_t / 1e9
    iter_bytes = spmv_bytes + (4 + 6) * 8 * n                            # update: q, r, dinv, r' ; p-update: p, x, r, dinv, p', x'
    iter_gbs = iter_bytes / (t_cg / (K_ * args.cg_iters)) / 1e9
    asm_bytes = 320 * nel_local                                           # 16 conn + 16 coords + 288 CSR values
    asm_gbs = asm_bytes / (t_asm / K_) / 1e9
    asm_traffic = None
    try:
        asm_traffic = json.load(open(os.path.join(ROOT, "profiles", "asm_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass

    line = {
        "metric": METRIC, "value": cg_rate, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": args.warmup,
        "ms_per_step": 1e3 * (t_asm + t_cg) / K_, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(world, cells, args.scaling, args.total_cols), "dof": int(prob.n_dof_global), "free_dof": int(n_free_global),
                   "elements": int(prob.nel_global), "nnz_per_gpu": int(nnz), "cg_iters_per_step": args.cg_iters,
                   "l2": "inputs (>7 GB CSR per GPU) exceed the 126 MB L2; no flush",
                   "parallelism": f"strip{world}", "transport": prob.transport if world > 1 else "single GPU"},
        "assembly": {"value": asm_rate, "unit": "elements/s", "ms": 1e3 * t_asm / K_,
                     "roofline": {"bound": "hbm", "achieved": asm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": asm_gbs / hbm_peak,
                                  "bytes_per_element": 320, "kernel": "lattice_q4_kernel<true> (marching warps, TMA bulk stores)",
                                  "traffic": asm_traffic}},
        "roofline": {"kernel": ("spmv_sl_kernel<true> (PCG SpMV + dot on symmetric-lattice storage: marching warps, 152 B of matrix per node, "
                                "bit-identical to the CSR kernels)" if sym_lat else
                                f"spmv_nb_kernel<true> (PCG SpMV + dot on the 2x2 node-block structure, TMA-staged, {8 + idx_bytes / 4:g} B/non-zero)" if node_block
                                else "spmv2_kernel<true> (PCG SpMV + dot, CSR, TMA-staged, 12 B/non-zero)"),
                     "bound": "hbm", "achieved": spmv_gbs, "peak": hbm_peak,
                     "unit": "GB/s", "frac": spmv_gbs / hbm_peak, "traffic": None, "peak_source": peak_src,
                     "bytes_per_launch": spmv_bytes, "ms_per_launch": 1e3 * spmv_t,
                     "csr_equivalent": {"bytes_per_launch": csr_bytes, "gbs": csr_bytes / spmv_t / 1e9,
                                        "note": "SURVEY 8d counts the CSR formulation, 12 B per non-zero"}},
        "pcg_iteration": {"achieved": iter_gbs, "unit": "GB/s", "frac": iter_gbs / hbm_peak, "bytes_per_dof_iter": iter_bytes / n,
                          "ms": 1e3 * t_cg / (K_ * args.cg_iters)},
        "gpu_launches": launches, "clocks": clocks, "timed_region_wall_s": wall,
    }
    traffic_file = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            key = "sym_lattice" if sym_lat else ("node_block" if node_block else "csr")
            line["roofline"]["traffic"] = (tj.get(key) or {}).get("dram_bytes_per_launch")
        except Exception:
            pass

    # anti-shortcut variant: interior nodes jittered by U(-0.2h, 0.2h), every element distinct -- same kernel
    try:
        g = torch.Generator(device="cuda")
        g.manual_seed(0)
        nodes0 = prob.nodes
        m_ = prob.part.m
        jit = (torch.rand(nodes0.shape, generator=g, device="cuda", dtype=torch.float64) - 0.5) * 0.4e-3
        ii = torch.arange(nodes0.shape[0], device="cuda")
        interior = ((ii % m_) > 0) & ((ii % m_) < m_ - 1) & (ii >= m_) & (ii < nodes0.shape[0] - m_)
        prob.nodes = nodes0 + jit * interior[:, None]
        ts = []
        for _ in range(4):
            a, b = ev(), ev()
            a.record(); prob.assemble(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        prob.nodes = nodes0
        prob.assemble()                                    # restore K of the regular plate
        line["assembly"]["jittered_mesh_ms"] = maxr(min(ts[1:]))
    except Exception as e:                                 # never lose the bench line over the side measurement
        line["assembly"]["jittered_mesh_ms"] = f"skipped: {e}"
    if world == 1 and not (args.no_cpu or args.no_e2e or args.no_full_solve):
        args.full_solve = True                             # the default run also carries the north-star solve to 1e-8
    if args.full_solve:                                    # before the e2e legs, which release the device-resident state
        line["full_solve"] = prob.full_solve(1e-8)
    # ---- end to end through the public API with HOST buffers ------------------------------------
    if not args.no_e2e and world == 1:
        line["e2e"] = e2e_leg(psfem_b200, prob, args, cells)
    elif world > 1:
        line["e2e"] = prob.e2e_distributed(args.e2e_iters, min(K_, 2))
    if rank == 0 and not args.no_cpu and world == 1:
        r = cpu_run(1, 0)
        line["cpu_baseline"] = {"value": r["cg_rate"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                "assembly_elements_per_s": r["asm_rate"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_leg(psfem_b200, prob, args, cells):
    """The call a user of the reference makes, host buffers in and out, copies inside the timed region:
       K = Polygones.global_stiffness_matrix(nodes, elements, E, nu, t, 'quad')      (H2D mesh, D2H CSR)
       U = static_solve(K_host, F_host, bc, maxit=e2e_iters)                        (H2D CSR + F, D2H U)."""
    import torch
    import scipy.sparse as sparse
    P = psfem_b200.Polygones
    prob.release()                                     # free the device-resident benchmark state first
    torch.cuda.empty_cache()
    h = 1e-3
    nodes, elements = P.mesh(cells * h, cells * h, cells, cells, element_type="quad")
    F = P.edgeForces(np.zeros(2 * len(nodes)), [0, -1e6], "right", cells * h, cells, cells)
    bc = P.boundary("left", [0, 0], cells, cells)
    steps = min(args.steps, 2)
    t_asm, t_cg = 0.0, 0.0
    h2d = d2h = 0
    for s in range(steps + 1):
        elements._plan_keep = elements._plan                   # symbolic plan is reused like in the device loop
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = P.global_stiffness_matrix(nodes, elements, E, NU, T, "quad")
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        Kh = sparse.csr_matrix((K.data, K.indices, K.indptr), shape=K.shape, copy=False)   # plain host CSR: no cached device copy
        del K
        t2 = time.perf_counter()
        try:
            U, U_red, cd, info = psfem_b200.static_solve(Kh, F, bc, rtol=1e-8, maxit=args.e2e_iters)
            its = info["iterations"]
        except psfem_b200.ConvergenceError:
            its = args.e2e_iters
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if s >= 1:
            t_asm += t1 - t0
            t_cg += t3 - t2
            h2d = nodes.nbytes + Kh.data.nbytes + Kh.indices.nbytes + Kh.indptr.nbytes + F.nbytes
            d2h = Kh.data.nbytes + Kh.indices.nbytes + Kh.indptr.nbytes + F.nbytes
        n_free = Kh.shape[0] - len(set(2 * k for k, _ in bc)) * 2
        del Kh
    return {"value": n_free * args.e2e_iters * steps / t_cg, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "pcg_iters_per_call": args.e2e_iters,
            "assembly_elements_per_s": cells * cells * steps / t_asm,
            "calls": "Polygones.global_stiffness_matrix(host nodes, elements) -> host scipy CSR; "
                     "static_solve(host CSR, host F, bc, maxit) -> host U (nodes, F: the caller's pageable arrays; the CSR arrays are the "
                     "page-locked buffers global_stiffness_matrix returned)"}


if __name__ == "__main__":
    main()
