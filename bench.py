#!/usr/bin/env python
"""bench.py -- headline benchmark of the plane-stress FEM hot path on B200.

Metric (BASELINE.json): "Q4 elements assembled/s + CG DOF*iter/s, 33.6M-DOF plate, 1/2/4/8 B200".
Workload (configs[3] at N=1): synthetic 4096x4096 Q4 plate (16 785 409 nodes, 33 570 818 DOF), left edge clamped,
right edge loaded -- mesh(4.096, 4.096, 4096, 4096, element_type='quad').  For N>1 every rank owns a strip of 4096
cell columns of a (4096*N)x4096 plate (weak scaling; N=4 is BASELINE's configs[4], the 16384x4096 strip); the halo
columns and the CG dot products travel over NVLink peer memory inside the PCG kernels (PSFEM_DIST_TRANSPORT=nccl
selects NCCL between the kernels instead).

One STEP = one numeric assembly of K over all elements of the rank + `--cg-iters` Jacobi-PCG iterations on the clamped
system.  `value` = CG DOF*iterations/s (whole job); the assembly rate is in the `assembly` object.  Device timing with
CUDA events on the launching stream, max over ranks.  All inputs (> 2.5 GB of matrix per rank) are far larger than the
126 MB L2, so no explicit L2 flush.

Besides the contract keys the line carries
  roofline / assembly.roofline   dominant kernels against the measured HBM peak (algorithmic bytes)
  e2e                            the same metric through the public API with HOST buffers (copies inside the timed region)
  cpu_baseline                   the oracle port on the host cores, SAME plate (bounded sample), BLAS pinned to one thread
  full_solve                     (N=1) the north-star solve of the plate to 1e-8
  extras                         (N=1) mass / K+M / T3 assembly, symbolic phase, Newmark step at C4, configs 2 and 3
  dist_check / strong            (N>1) multi-GPU correctness checks run before timing; strong scaling of C4 and C5

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--cells 4096]
"""
import argparse
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
ASM_BYTES_PER_Q4 = 304        # 16 B coordinates + 288 B CSR values; the lattice kernel reads no connectivity (SURVEY 8d counts 320 with it)

# the UNMODIFIED reference, timed in the build container (8 cores; /root/reference does not exist on the GPU box):
# BASELINE.md section 2 / section 3 (iii), "for the record"
REFERENCE_RECORDED = {
    "where": "build container, 8 host cores, reference imported unmodified (matplotlib stubbed)",
    "c1_T3_global_stiffness_matrix_s": 0.19, "c1_T3_elements_per_s": 16800.0,
    "q4_16x16_global_stiffness_matrix_elements_per_s": 2.1,
    "main_py_T3_10001_steps_s": 4.0, "freeModes_py_T3_80x8_s": 3.0,
}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def plate_counts(cols, rows):
    """(nodes, dof, free dof with the left edge clamped, elements, structural nnz) of a cols x rows Q4 plate."""
    n, m = cols + 1, rows + 1
    return n * m, 2 * n * m, 2 * (n - 1) * m, cols * rows, 4 * (3 * n - 2) * (3 * m - 2)


def workload_name(n_gpus, cells, scaling="weak", total_cols=0):
    if total_cols:
        return (f"synthetic {total_cols}x{cells} Q4 strip, left edge clamped, right edge loaded: numeric assembly of K + "
                f"Jacobi-PCG iterations; split into {n_gpus} strips of ~{total_cols // n_gpus} cell columns")
    if scaling == "strong":
        return (f"synthetic {cells}x{cells} Q4 plate, left edge clamped, right edge loaded: numeric assembly of K + "
                f"Jacobi-PCG iterations; the plate is split into {n_gpus} strips of ~{cells // n_gpus} cell columns")
    return (f"synthetic {cells * n_gpus}x{cells} Q4 plate, left edge clamped, right edge loaded: numeric assembly of K + "
            f"Jacobi-PCG iterations; {cells} cell columns per GPU (strip partition)")


def config_of(args, world):
    """The `config` object of BOTH arms (pure arithmetic, no GPU): the reference arm runs on this arm's config."""
    cols = args.total_cols or (args.cells * world if args.scaling == "weak" else args.cells)
    _, dof, free, nel, _ = plate_counts(cols, args.cells)
    return {"workload": workload_name(world, args.cells, args.scaling, args.total_cols), "dof": int(dof), "free_dof": int(free),
            "elements": int(nel), "cg_iters_per_step": args.cg_iters,
            "l2": "inputs (> 2.5 GB of matrix per GPU) exceed the 126 MB L2; no flush",
            "parallelism": f"strip{world}",
            "transport": (os.environ.get("PSFEM_DIST_TRANSPORT", "peer") if world > 1 else "single GPU")}


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
# reference arm / cpu baseline: the oracle (NumPy/SciPy restatement of the reference) on the host cores
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuPlate:
    """The SAME plate as the GPU arm on the host (one GPU's share: cells x cells Q4 cells, left edge clamped, right edge
    loaded), built once with the oracle in threaded slabs (untimed set-up; the one-shot COO of 16.7 M elements would
    need 17 GB).  A step = a bounded sample of the workload with every host thread busy and BLAS / OpenMP pinned to ONE
    thread per pool thread (threadpoolctl), so that the same code prints the same rate under `python` and under
    `torchrun` (which exports OMP_NUM_THREADS=1): round 1's N=1 leg lost 5x to oversubscription.
      assembly  NumPy Ke + scipy coo->csr of `asm_cols` cell columns (x cells rows) of that plate, threaded slabs
      CG        `cg_iters` Jacobi-PCG iterations (SciPy CSR mat-vec over row blocks on all threads) on the whole plate"""

    def __init__(self, cells, asm_cols=128, cg_iters=5):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import psfem_oracle as O
        from threadpoolctl import threadpool_limits
        self.O, self.limits = O, threadpool_limits
        self.cells, self.cores = cells, host_threads()
        self.asm_cols, self.cg_iters = min(asm_cols, cells), cg_iters
        h = 1e-3
        t0 = time.perf_counter()
        with threadpool_limits(1):
            self.nodes, self.conn = O.mesh(cells * h, cells * h, cells, cells, element_type="quad")
            self.K_red = O.assemble_chunked(self.nodes, self.conn, cells, cells, E, NU, T, self.cores, drop_first_node_column=True)
            F = O.edge_forces(np.zeros(2 * len(self.nodes)), [0, -1e6], "right", cells * h, cells, cells)
            self.F_red = F[2 * (cells + 1):]
            self.split = O.row_blocks(self.K_red, self.cores)       # row blocks of the threaded mat-vec: set-up, once
        self.n_free = self.K_red.shape[0]
        self.setup_s = time.perf_counter() - t0

    def step(self):
        O = self.O
        with self.limits(1):
            t0 = time.perf_counter()
            O.assemble_chunked(self.nodes, self.conn, self.cells, self.cells, E, NU, T, self.cores, col0=0, col1=self.asm_cols)
            t_asm = time.perf_counter() - t0
            _, _, t_cg = O.jacobi_pcg_threads(self.K_red, self.F_red, self.cg_iters, self.cores, split=self.split)
        return t_asm, t_cg

    def run(self, steps, warmup):
        ta, tc = [], []
        for s in range(warmup + steps):
            a, c = self.step()
            if s >= warmup:
                ta.append(a)
                tc.append(c)
        # rows of asm_cols node columns are produced from (asm_cols) cell columns (+ nothing before column 0)
        asm_rate = self.asm_cols * self.cells * len(ta) / sum(ta)
        cg_rate = self.n_free * self.cg_iters * len(tc) / sum(tc)
        sample = (f"per step, on {self.cores} threads (BLAS/OpenMP pinned to 1 per thread): Q4 assembly (NumPy Ke + scipy coo->csr, "
                  f"threaded slabs) of {self.asm_cols} cell columns of the {self.cells}x{self.cells} plate and {self.cg_iters} Jacobi-PCG "
                  f"iterations (SciPy CSR mat-vec over row blocks) on the whole clamped plate ({self.n_free} DOF)")
        return dict(cg_rate=cg_rate, asm_rate=asm_rate, cores=self.cores, sample=sample,
                    ms_per_step=1e3 * (sum(ta) + sum(tc)) / len(tc), setup_s=self.setup_s)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    r = CpuPlate(args.cells).run(args.steps, args.warmup)
    base = {"value": r["cg_rate"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
            "assembly_elements_per_s": r["asm_rate"], "setup_s": r["setup_s"], "reference_unmodified": REFERENCE_RECORDED}
    line = {
        "impl": "reference", "metric": METRIC, "value": r["cg_rate"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args, world),
        "assembly": {"value": r["asm_rate"], "unit": "elements/s"},
        "cpu_baseline": base,
        "e2e": {"value": r["cg_rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the unmodified reference (pure Python, dense (2n)^2 storage, ~2 Q4 elements/s) cannot run this workload; this arm "
                "times oracle/psfem_oracle.py, the NumPy/SciPy restatement pinned to the reference's goldens, on one GPU's share of "
                "the plate (the per-DOF rate does not depend on the number of strips)",
    }
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--no-extras", action="store_true", help="skip the N=1 side records (M / K+M / T3 assembly, Newmark step, configs 2-3)")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the dist check and the strong-scaling records")
    ap.add_argument("--full-solve", action="store_true", help="also solve the plate to 1e-8 and report iterations/seconds "
                    "(on by default for the plain single-GPU run: 24 500 iterations, about 22 s)")
    ap.add_argument("--no-full-solve", action="store_true")
    args = ap.parse_args()
    if args.total_cols:
        args.scaling = "strong"
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import psfem_b200
    from psfem_b200.dist import StripProblem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=150))   # a rank that dies must not hold the others (and the GPUs) for NCCL's default 10 minutes
    hbm_peak, peak_src = peaks()
    cells = args.cells

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    line_extra = {}
    # ---- N>1: correctness of the strip solver on THIS box before anything is timed ----------------------------
    if world > 1 and not args.no_strong:
        try:
            from psfem_b200.dist_check import run_checks
            chk = run_checks(world, rank)
            line_extra["dist_check"] = "ok"
            line_extra["dist_check_detail"] = {"what": "peer == NCCL iterates, owned rows bit-identical to the single-GPU matrix (Q4, T3, T6, Q9), "
                                                       "restart converges, rollers + pinned node through dist.static_solve == single-GPU static_solve",
                                               **{k: (float(v) if isinstance(v, float) else v) for k, v in chk.items()}}
        except Exception as e:                                  # noqa: BLE001 -- report, never hide
            line_extra["dist_check"] = f"FAILED: {type(e).__name__}: {e}"[:400]

    cols_total = args.total_cols or (cells * world if args.scaling == "weak" else cells)
    prob = StripProblem(L=cols_total * 1e-3, l=cells * 1e-3, N=cols_total, M=cells, rank=rank, world=world, E=E, nu=NU, t=T)
    t0 = time.perf_counter()
    prob.setup()                       # mesh strip, symbolic pattern, first assembly, BC reduction, load vector
    setup_s = time.perf_counter() - t0
    n_free_global = prob.n_free_global
    nel_local = prob.nel_owned
    ev = lambda: torch.cuda.Event(enable_timing=True)

    sampler = ClockSampler(local)
    sampler.start()                                        # nvidia-smi is up before the first warm-up step
    asm_ms, cg_ms, spmv_ms = [], [], []
    launches = 0
    kpi = prob.kernels_per_iteration
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
        prob.cg_iterations(args.cg_iters)                  # the PCG kernels (+ halo / reductions fused in when world > 1)
        e2.record()
        prob.spmv_only(args.cg_iters)                      # the dominant kernel alone, for its roofline line
        e3.record()
        torch.cuda.synchronize()
        if s >= args.warmup:
            asm_ms.append(e0.elapsed_time(e1))
            cg_ms.append(e1.elapsed_time(e2))
            spmv_ms.append(e2.elapsed_time(e3))
            launches += 1 + kpi * args.cg_iters + args.cg_iters   # assembly + PCG kernels + the SpMV-only leg
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()

    t_asm = maxr(sum(asm_ms)) / 1e3
    t_cg = maxr(sum(cg_ms)) / 1e3
    t_spmv = maxr(sum(spmv_ms)) / 1e3
    K_ = args.steps
    cg_rate = n_free_global * args.cg_iters * K_ / t_cg
    asm_rate = prob.nel_global * K_ / t_asm

    # roofline of the dominant kernel (SpMV + dot of the PCG iteration), per rank
    nnz, n = prob.K_red.nnz, prob.n_owned_free
    csr_bytes = 12 * nnz + 4 * (n + 1) + 16 * n                          # CSR SpMV
    nbp = prob.K_red.plan()[3]
    node_block = nbp is not None and not os.environ.get("PSFEM_NO_NODEBLOCK")
    idx_bytes = 2 if (node_block and nbp[5] is not None) else 4
    spmv_bytes = (8 * nnz + idx_bytes * (nnz // 4) + 4 * (n // 2 + 1) + 16 * n) if node_block else csr_bytes
    sym_lat = bool(getattr(prob, "sym_lattice", False)) and not os.environ.get("PSFEM_NO_SYMLATTICE")
    if sym_lat:
        spmv_bytes = (19 * 8 + 16 + 16) * (n // 2)                        # 19 doubles of matrix per node + x + y
    spmv_t = t_spmv / (K_ * args.cg_iters)
    spmv_gbs = spmv_bytes / spmv_t / 1e9
    iter_bytes = prob.iteration_bytes(spmv_bytes)
    iter_gbs = iter_bytes / (t_cg / (K_ * args.cg_iters)) / 1e9
    asm_bytes = ASM_BYTES_PER_Q4 * nel_local
    asm_gbs = asm_bytes / (t_asm / K_) / 1e9
    same_workload = cells == 4096 and world == 1                          # committed ncu traffic figures are for this configuration only

    def traffic(name, key=None):
        if not same_workload:
            return None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            return (tj.get(key) or {}).get("dram_bytes_per_launch") if key else tj.get("dram_bytes_per_launch")
        except Exception:
            return None

    iter_t = t_cg / (K_ * args.cg_iters)
    if kpi == 1:
        # the whole PCG iteration is ONE kernel (psfem_cg_sl.cuh): it is the dominant kernel of the step (25 launches against 1 assembly)
        roofline = {"kernel": "cg_sl_tma_kernel (one Jacobi-PCG iteration per launch: Chronopoulos-Gear recurrence fused into the marching-warp "
                              "symmetric-lattice SpMV, operands staged by TMA bulk copies + mbarriers; 152 B matrix + 128 B vectors + 16 B x per node)",
                    "bound": "hbm", "achieved": iter_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": iter_gbs / hbm_peak,
                    "traffic": traffic("spmv_traffic.json", "cg_fused"), "peak_source": peak_src,
                    "bytes_per_launch": iter_bytes, "ms_per_launch": 1e3 * iter_t,
                    "csr_equivalent": {"bytes_per_launch": csr_bytes + 13 * 8 * n - 16 * n, "note": "SURVEY 8d: 323.9 B per DOF*iteration for CSR SpMV + 3 vector kernels"}}
    else:
        roofline = {"kernel": prob.dominant_kernel(sym_lat, node_block, idx_bytes), "bound": "hbm", "achieved": spmv_gbs, "peak": hbm_peak,
                    "unit": "GB/s", "frac": spmv_gbs / hbm_peak,
                    "traffic": traffic("spmv_traffic.json", "sym_lattice" if sym_lat else ("node_block" if node_block else "csr")),
                    "peak_source": peak_src, "bytes_per_launch": spmv_bytes, "ms_per_launch": 1e3 * spmv_t}
    cfg = config_of(args, world)                # identical to the reference arm's config (nothing solver-specific in it)
    line = {
        "metric": METRIC, "value": cg_rate, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": args.warmup,
        "ms_per_step": 1e3 * (t_asm + t_cg) / K_, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "assembly": {"value": asm_rate, "unit": "elements/s", "ms": 1e3 * t_asm / K_,
                     "roofline": {"bound": "hbm", "achieved": asm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": asm_gbs / hbm_peak,
                                  "bytes_per_element": ASM_BYTES_PER_Q4,
                                  "bytes_note": "16 B coordinates + 288 B CSR values; connectivity is not read (SURVEY 8d's 320 B counts it)",
                                  "kernel": "lattice_q4_kernel<true> (marching warps, TMA bulk stores)",
                                  "traffic": traffic("asm_traffic.json")}},
        "roofline": roofline,
        "spmv": {"kernel": prob.dominant_kernel(sym_lat, node_block, idx_bytes), "achieved": spmv_gbs, "unit": "GB/s", "frac": spmv_gbs / hbm_peak,
                 "bytes_per_launch": spmv_bytes, "ms_per_launch": 1e3 * spmv_t,
                 "traffic": traffic("spmv_traffic.json", "sym_lattice" if sym_lat else ("node_block" if node_block else "csr")),
                 "note": "the SpMV + dot kernel alone (the operator of every other driver: Newmark right-hand sides, energies, Lanczos); "
                         "inside the PCG it is part of the fused iteration kernel when the operand is a symmetric lattice matrix",
                 "csr_equivalent": {"bytes_per_launch": csr_bytes, "gbs": csr_bytes / spmv_t / 1e9,
                                    "note": "SURVEY 8d counts the CSR formulation, 12 B per non-zero"}},
        "pcg_iteration": {"achieved": iter_gbs, "unit": "GB/s", "frac": iter_gbs / hbm_peak, "bytes_per_dof_iter": iter_bytes / n,
                          "ms": 1e3 * t_cg / (K_ * args.cg_iters), "kernels": kpi,
                          "survey_8d_csr_bytes_per_dof_iter": 323.9},
        "solver": {"pcg": prob.pcg_variant, "nnz_per_gpu": int(nnz), "transport": prob.transport if world > 1 else "single GPU"},
        "gpu_launches": launches, "clocks": clocks, "timed_region_wall_s": wall, "setup_s": setup_s,
    }
    line.update(line_extra)

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
        del jit, ii, interior
    except Exception as e:                                 # never lose the bench line over the side measurement
        line["assembly"]["jittered_mesh_ms"] = f"skipped: {e}"
    if world == 1 and not (args.no_cpu or args.no_e2e or args.no_full_solve):
        args.full_solve = True                             # the default run also carries the north-star solve to 1e-8
    if args.full_solve:                                    # before the e2e legs, which release the device-resident state
        line["full_solve"] = prob.full_solve(1e-8)
        try:
            line["full_solve"]["true_relres"] = true_residual(StripProblem, prob, cols_total, cells, world, rank)
            line["full_solve"]["true_relres_note"] = ("||F - K U|| / ||F|| of the returned solution, recomputed from scratch"
                                                      + (": the strips' solution gathered on rank 0 and multiplied with the matrix of the WHOLE plate, "
                                                         "assembled on one GPU in the lattice-only set-up" if world > 1 else ""))
        except Exception as e:                             # noqa: BLE001
            line["full_solve"]["true_relres"] = f"skipped: {type(e).__name__}: {e}"[:300]
    if world == 1 and not args.no_extras:
        try:
            line["extras"] = extras_leg(psfem_b200, prob, cells, hbm_peak)
        except Exception as e:                             # noqa: BLE001
            line["extras"] = f"skipped: {type(e).__name__}: {e}"[:300]
    # ---- N>1: strong scaling of C4 and C5, each against the single-GPU iteration measured in this same run ----
    if world > 1 and not args.no_strong:
        prob.release()
        torch.cuda.empty_cache()
        try:
            line["strong"] = strong_leg(StripProblem, world, rank, cells, args.cg_iters, barrier, maxr)
        except Exception as e:                             # noqa: BLE001
            line["strong"] = f"FAILED: {type(e).__name__}: {e}"[:400]
    # ---- end to end through the public API with HOST buffers ------------------------------------
    if not args.no_e2e:
        if world == 1:
            line["e2e"] = e2e_leg(psfem_b200, prob, args, cells)
        else:
            line["e2e"] = e2e_distributed(psfem_b200, prob, args, cols_total, cells, world, rank, barrier, maxr)
    if rank == 0 and not args.no_cpu and world == 1:
        r = CpuPlate(cells).run(1, 0)
        line["cpu_baseline"] = {"value": r["cg_rate"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                "assembly_elements_per_s": r["asm_rate"], "setup_s": r["setup_s"],
                                "reference_unmodified": REFERENCE_RECORDED}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def true_residual(StripProblem, prob, cols_total, cells, world, rank):
    """Independent check of a finished solve: the residual of the returned displacement vector, not of the recurrence.
    Several GPUs: rank 0 gets the gathered solution, assembles the WHOLE plate on its own GPU (lattice-only set-up: the
    134 M-DOF strip needs 10 GB there) and multiplies."""
    import torch
    if world == 1:
        prob.finish_x()
        x = prob.st["x"]
        r = prob.b_own - prob.K_red.matvec(x)
        return float(torch.linalg.vector_norm(r).item() / torch.linalg.vector_norm(prob.b_own).item())
    U = prob.gather_solution(dst=0)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    if rank == 0:
        one = StripProblem(L=cols_total * 1e-3, l=cells * 1e-3, N=cols_total, M=cells, rank=0, world=1, E=E, nu=NU, t=T).setup(lattice_only=True)
        x = torch.from_numpy(U[2 * (cells + 1):]).cuda()          # the clamped first node column is not part of the system
        r = one.b_own - one.K_red.matvec(x)
        out[0] = torch.linalg.vector_norm(r) / torch.linalg.vector_norm(one.b_own)
        one.release()
    import torch.distributed as dist
    dist.broadcast(out, src=0)
    return float(out.item())


def strong_leg(StripProblem, world, rank, cells, cg_iters, barrier, maxr):
    """Strong scaling inside the driver's weak run: the C4 plate (cells x cells) and BASELINE's C5 strip (4*cells x cells)
    split over all ranks, PCG iteration time against ONE GPU running the whole C4 plate, measured here on every rank
    (independent replicas, max over ranks).  C5 does not fit one GPU's int32 indices (2.4e9 non-zeros), its reference
    is 4 x the single-GPU C4 iteration (the same per-DOF rate)."""
    import torch
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def time_iters(prob, reps=4, iters=50):
        ts = []
        for _ in range(reps):
            barrier()
            a, b = ev(), ev()
            a.record(); prob.cg_iterations(iters); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / iters)
        return maxr(min(ts[1:]))

    out = {}
    one = StripProblem(L=cells * 1e-3, l=cells * 1e-3, N=cells, M=cells, rank=0, world=1, E=E, nu=NU, t=T).setup()
    t1 = time_iters(one)
    n_free_c4 = one.n_free_global
    one.release()
    torch.cuda.empty_cache()
    out["single_gpu_c4_ms_per_iter"] = t1
    t1_c5 = None
    try:     # the C5 strip fits ONE GPU in the solve-only set-up (no CSR): its single-GPU iteration is measured, not extrapolated
        big = StripProblem(L=4 * cells * 1e-3, l=cells * 1e-3, N=4 * cells, M=cells, rank=0, world=1, E=E, nu=NU, t=T).setup(lattice_only=True)
        t1_c5 = time_iters(big, reps=3, iters=20)
        big.release()
        del big
        torch.cuda.empty_cache()
        out["single_gpu_c5_ms_per_iter"] = t1_c5
    except Exception as e:                                 # noqa: BLE001
        out["single_gpu_c5_ms_per_iter"] = f"skipped: {type(e).__name__}: {e}"[:200]
    for name, cols in (("c4", cells), ("c5", 4 * cells)):
        if cols // world < 2:
            continue
        p = StripProblem(L=cols * 1e-3, l=cells * 1e-3, N=cols, M=cells, rank=rank, world=world, E=E, nu=NU, t=T).setup()
        tn = time_iters(p)
        ref = t1 * (cols // cells)
        if name == "c5" and t1_c5:
            ref = t1_c5
        out[name] = {"plate": f"{cols}x{cells}", "free_dof": int(p.n_free_global), "ms_per_iter": tn,
                     "dof_iter_per_s": p.n_free_global / (tn * 1e-3), "speedup_vs_one_gpu": ref / tn,
                     "one_gpu_reference_ms": ref, "transport": p.transport, "pcg": p.pcg_variant,
                     "cols_per_gpu": cols // world}
        p.release()
        torch.cuda.empty_cache()
    out["note"] = ("speedup = (single-GPU C4 iteration measured in this run x plate size / C4 size) / N-GPU iteration; device-timed, "
                   "max over ranks, best of 3 batches of 50 iterations")
    return out


def extras_leg(psfem_b200, prob, cells, hbm_peak):
    """Driver-visible numbers for the rest of the path at C4 (VERDICT r1 missing #5): mass assembly, K+M in one pass,
    T3 assembly, symbolic phase, a Newmark-beta step (main.py:304-332 recurrence) -- and BASELINE configs 2 and 3 as wall
    time next to the unmodified reference's recorded run."""
    import torch
    from psfem_b200 import Polygones as P
    from psfem_b200.device import Assembler, device_mesh
    ev = lambda: torch.cuda.Event(enable_timing=True)
    f64 = dict(dtype=torch.float64, device="cuda")

    def best(fn, reps=4):
        ts = []
        for _ in range(reps):
            a, b = ev(), ev()
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return min(ts[1:])

    out = {}
    nel = prob.nel_owned
    asm = prob.asm
    import ctypes as C
    from psfem_b200._capi import call, ptr, stream_ptr, workspace
    tf = C.c_double(0.0)
    ws_ = workspace(256)
    call("psfem_fp64_peak", C.byref(tf), ptr(ws_), 256, stream_ptr())
    out["fp64_peak_tflops_measured"] = tf.value
    M_vals = torch.empty(asm.nnz, **f64)
    K2 = torch.empty(asm.nnz, **f64)
    ms = best(lambda: asm.assemble(prob.nodes, rho=RHO, t=T, what=2, M_out=M_vals))
    k_ms = best(lambda: asm.assemble(prob.nodes, E=E, nu=NU, t=T, what=1, K_out=prob.K_vals))
    # SASS of lattice_q4_kernel<true,4,2,false> (profiles/r2_sass.md): 219 DFMA + 179 DMUL + 143 DADD per marching step of a
    # lane = 760 flop per element; one halo lane in 32 and one halo column per 64-column segment recompute elements (x 1.05)
    flop_per_el = (2 * 219 + 179 + 143) * 1.05
    out["assembly_K_isolated"] = {"ms": k_ms, "elements_per_s": nel / k_ms * 1e3, "hbm_frac": ASM_BYTES_PER_Q4 * nel / (k_ms * 1e-3) / 1e9 / hbm_peak,
                                  "fp64_tflops": flop_per_el * nel / (k_ms * 1e-3) / 1e12,
                                  "fp64_frac_of_measured_peak": flop_per_el * nel / (k_ms * 1e-3) / 1e12 / max(tf.value, 1e-9),
                                  "note": "back-to-back launches outside the PCG loop; co-bound by HBM and the fp64 pipe (ncu r1: 64 % of the pipe)"}
    out["assembly_M"] = {"ms": ms, "elements_per_s": nel / ms * 1e3, "bytes_per_element": ASM_BYTES_PER_Q4,
                         "hbm_frac": ASM_BYTES_PER_Q4 * nel / (ms * 1e-3) / 1e9 / hbm_peak}
    ms = best(lambda: asm.assemble(prob.nodes, E=E, nu=NU, rho=RHO, t=T, what=3, K_out=K2, M_out=M_vals))
    bkm = 16 + 2 * 288
    out["assembly_KM"] = {"ms": ms, "elements_per_s": nel / ms * 1e3, "bytes_per_element": bkm,
                          "hbm_frac": bkm * nel / (ms * 1e-3) / 1e9 / hbm_peak}
    del K2
    # K straight into the solver's symmetric-lattice operand (what a solve-only flow runs instead of the CSR assembly)
    from psfem_b200.device import assemble_lattice_sl
    n_cols, m_rows = prob.part.e1 - prob.part.e0, prob.part.m
    ms = best(lambda: assemble_lattice_sl(prob.nodes, n_cols, m_rows, 1, n_cols - 1, E=E, nu=NU, t=T, what=1))
    bsl = 16 + 19 * 8
    out["assembly_SL"] = {"ms": ms, "elements_per_s": nel / ms * 1e3, "bytes_per_element": bsl,
                          "hbm_frac": bsl * nel / (ms * 1e-3) / 1e9 / hbm_peak,
                          "what": "psfem_assemble_lattice_sl: the clamped plate's K as 19 doubles per node, no CSR values / indices / reduction"}
    # symbolic phase of the lattice (topology check + closed-form pattern), once per mesh
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a2 = Assembler(prob.conn, prob.part.n_ext_nodes, asm.etype)
    torch.cuda.synchronize()
    out["symbolic_ms"] = 1e3 * (time.perf_counter() - t0)
    out["symbolic_note"] = "lattice detection on the connectivity + row_ptr / col_idx / node_ptr written in closed form (2.5 GB), host-timed"
    del a2
    # T3 on the same nodes (33.5 M triangles)
    _, conn3 = device_mesh(cells * 1e-3, cells * 1e-3, cells, cells, 0, 0, "tri", "coarse")
    a3 = Assembler(conn3, prob.part.n_ext_nodes, psfem_b200._capi.T3)
    K3 = torch.empty(a3.nnz, **f64)
    ms = best(lambda: a3.assemble(prob.nodes, E=E, nu=NU, t=T, what=1, K_out=K3))
    b3 = 8 + 14 * 8
    out["assembly_T3"] = {"ms": ms, "elements_per_s": a3.nel / ms * 1e3, "bytes_per_element": b3,
                          "hbm_frac": b3 * a3.nel / (ms * 1e-3) / 1e9 / hbm_peak}
    del a3, K3, conn3
    torch.cuda.empty_cache()
    # Newmark-beta step at C4: dt ~ h / wave speed (K_eff = K + M/(beta dt^2) well conditioned), warm-started PCG to 1e-8
    try:
        K_ext = asm.pattern_csr(prob.K_vals)
        Md = K_ext.with_vals(M_vals)
        dt, steps = 2e-7, 9
        bc = P.boundary("left", [0, 0], cells, cells)
        n_nodes = prob.part.n_ext_nodes
        shape, scale, _ = psfem_b200.ramp_load(n_nodes, cells * 1e-3, P=1e6, T=10000 * dt, dt=dt)
        nm = psfem_b200.newmark(K_ext, Md, bc, shape, scale[:steps], dt=dt, store_every=steps, energies=True, rtol=1e-8, maxit=300)
        its = nm["info"]["pcg_iterations"]
        out["newmark_step"] = {"ms_per_step": nm["info"]["run_ms"] / (steps - 1), "pcg_iterations_per_step": its / (steps - 1),
                               "dt": dt, "steps": steps - 1, "energies": True, "max_relres": nm["info"]["max_relres"],
                               "what": "predictor, RHS F - K(u_p + beta_R v_p) - alpha_R M v_p (2 SpMV), PCG solve with K_eff, corrector, "
                                       "stabiliser, 2 energy SpMV+dot (main.py:304-332, :377-385) on the clamped 4096x4096 plate"}
        del nm, Md, K_ext
    except Exception as e:                                 # noqa: BLE001
        out["newmark_step"] = f"skipped: {type(e).__name__}: {e}"[:300]
    del M_vals
    torch.cuda.empty_cache()
    # BASELINE's C5 strip (16384 x 4096, 134 M DOF) on ONE GPU: possible because the solve-only set-up never builds the CSR
    # side (2.4e9 non-zeros would not fit int32 indices); 10.2 GB of matrix
    try:
        from psfem_b200.dist import StripProblem
        prob.release()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        t0 = time.perf_counter()
        big = StripProblem(L=4 * cells * 1e-3, l=cells * 1e-3, N=4 * cells, M=cells, rank=0, world=1, E=E, nu=NU, t=T).setup(lattice_only=True)
        torch.cuda.synchronize()
        t_setup = time.perf_counter() - t0
        ms = best(lambda: big.cg_iterations(20), reps=3) / 20
        out["c5_single_gpu_lattice_only"] = {"plate": f"{4 * cells}x{cells}", "free_dof": int(big.n_free_global), "ms_per_iter": ms,
                                             "dof_iter_per_s": big.n_free_global / (ms * 1e-3), "setup_s": t_setup,
                                             "device_memory_gb": torch.cuda.max_memory_allocated() / 1e9}
        big.release()
        del big
        torch.cuda.empty_cache()
    except Exception as e:                                 # noqa: BLE001
        out["c5_single_gpu_lattice_only"] = f"skipped: {type(e).__name__}: {e}"[:300]
    # BASELINE configs 2 and 3 end to end (mesh -> K, M -> driver), wall clock, second of two runs
    try:
        L, l, N, G = 20, 2, 20, 2
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            nodes, elements = P.mesh(L, l, N, G, element_type="quad", mesh_type="fine", offsetY=l / 2)
            K = P.global_stiffness_matrix(nodes, elements, E, NU, 0.1, "quad", "fine")
            M = P.global_mass_matrix(nodes, elements, RHO, 0.1, "quad", "fine")
            bc = [(int(G / 2), [0, 0]), (int((len(nodes) - 1) - G / 2), [None, 0])]
            shape, scale, _ = psfem_b200.ramp_load(len(nodes), L)
            o = psfem_b200.newmark(K, M, bc, shape, scale, dt=1e-3)
            torch.cuda.synchronize(); t2 = time.perf_counter() - t0
        out["config2_newmark_q9_10001_steps"] = {"wall_s": t2, "kernel_ms": o["info"]["run_ms"], "free_dof": int(o["U"].shape[0]),
                                                 "reference_main_py_T3_123dof_s": REFERENCE_RECORDED["main_py_T3_10001_steps_s"],
                                                 "note": "the reference needs ~2.6 s per Q9 ELEMENT to assemble this mesh (sympy); its 4.0 s is main.py as shipped (T3, 123 DOF)"}
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            nodes, elements = P.mesh(20, 2, 80, 8, element_type="tri", mesh_type="fine", offsetY=1)
            K = P.global_stiffness_matrix(nodes, elements, E, NU, 0.01, "tri", "fine")
            M = P.global_mass_matrix(nodes, elements, RHO, 0.01, "tri", "fine")
            bc = [(4, [0, 0]), (int((len(nodes) - 1) - 4), [None, 0])]
            f, modes, cd, inf = psfem_b200.modal(K, M, bc, k=20)
            torch.cuda.synchronize(); t3 = time.perf_counter() - t0
        out["config3_modal_t6_80x8_20_modes"] = {"wall_s": t3, "free_dof": int(K.shape[0] - len(cd)), "lanczos_steps": inf["lanczos_steps"],
                                                 "f1_hz": float(f[0]), "converged": bool(inf["converged"]),
                                                 "reference_freeModes_py_T3_1455dof_s": REFERENCE_RECORDED["freeModes_py_T3_80x8_s"]}
    except Exception as e:                                 # noqa: BLE001
        out["configs_2_3"] = f"skipped: {type(e).__name__}: {e}"[:300]
    return out


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
    its_done = 0
    for s in range(steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = P.global_stiffness_matrix(nodes, elements, E, NU, T, "quad")
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        Kh = sparse.csr_matrix((K.data, K.indices, K.indptr), shape=K.shape, copy=False)   # plain host CSR: no cached device copy
        del K
        t2 = time.perf_counter()
        U, U_red, cd, info = psfem_b200.static_solve(Kh, F, bc, rtol=1e-8, maxit=args.e2e_iters, method="pcg", raise_on_fail=False)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if s >= 1:
            t_asm += t1 - t0
            t_cg += t3 - t2
            its_done += info["iterations"]
            h2d = nodes.nbytes + Kh.data.nbytes + Kh.indices.nbytes + Kh.indptr.nbytes + F.nbytes
            d2h = Kh.data.nbytes + Kh.indices.nbytes + Kh.indptr.nbytes + U.nbytes + U_red.nbytes
        n_free = U_red.shape[0]
        del Kh, U, U_red
    return {"value": n_free * its_done / t_cg, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "pcg_iters_per_call": its_done // steps,
            "assembly_elements_per_s": cells * cells * steps / t_asm,
            "calls": "Polygones.global_stiffness_matrix(host nodes, elements) -> host scipy CSR; "
                     "static_solve(host CSR, host F, bc, maxit, raise_on_fail=False) -> host U and U_reduced (nodes, F: the caller's "
                     "pageable arrays; the CSR arrays are the page-locked buffers global_stiffness_matrix returned)"}


def e2e_distributed(psfem_b200, prob, args, cols_total, cells, world, rank, barrier, maxr):
    """End to end at N>1 through the public multi-GPU call `psfem_b200.dist.static_solve`: host coordinates and host
    load in (each rank uploads the rows of its strip from pinned memory), device mesh connectivity, assembly, boundary
    conditions, `e2e_iters` PCG iterations, and the displacement vector gathered to the host of rank 0 -- all inside the
    timed region."""
    import torch
    from psfem_b200 import dist as pdist
    from psfem_b200.device import device_mesh
    if hasattr(prob, "asm"):
        prob.release()
    torch.cuda.empty_cache()
    p = prob.part
    h = 1e-3
    # host inputs of THIS rank's strip, page-locked (a user holds the global arrays; each rank touches only its rows)
    nd, _ = device_mesh(cols_total * h, cells * h, cols_total, cells, 0, 0, "quad", "coarse", col0=p.e0, ncols=p.e1 - p.e0,
                        cell0=p.cell0, ncells=p.cell1 - p.cell0)
    nodes_strip = nd.cpu().pin_memory().numpy()
    del nd

    def load(a, b):                                       # Polygones.edgeForces(zeros, [0, -1e6], 'right', ...) restricted to [a, b)
        Fh = np.zeros(b - a)
        m = p.m
        k0 = 2 * (p.n - 1) * m
        lo_, hi_ = max(a, k0), min(b, k0 + 2 * m)
        if hi_ > lo_:
            idx = np.arange(lo_, hi_)
            Fh[idx[idx % 2 == 1] - a] = -1e6 * (cells * h) / (m - 1)
        return Fh

    steps = min(args.steps, 2)
    times, infos = [], []
    for s in range(steps + 1):
        barrier()
        t0 = time.perf_counter()
        U, info = pdist.static_solve(cols_total * h, cells * h, cols_total, cells, F=load, E=E, nu=NU, t=T, nodes=_StripNodes(nodes_strip, p),
                                     rtol=1e-8, maxit=args.e2e_iters, check_every=args.e2e_iters, raise_on_fail=False)
        barrier()
        times.append(time.perf_counter() - t0)
        infos.append(info)
        del U
    t = maxr(sum(times[1:]))
    its = infos[-1]["iterations"]
    return {"value": prob.n_free_global * its * steps / t, "unit": UNIT, "h2d_bytes_per_step": int(infos[-1]["h2d_bytes"]),
            "d2h_bytes_per_step": int(infos[-1]["d2h_bytes"]), "steps": steps, "pcg_iters_per_call": int(its),
            "calls": "per rank: psfem_b200.dist.static_solve(L, l, N, M, F=host load, nodes=host coordinates, maxit) -> strip upload from "
                     "pinned host memory, device connectivity, pattern + assembly + BC, PCG iterations, D2H of the owned displacements, "
                     "gather of the global U on rank 0"}


class _StripNodes:
    """Stands for the global (n*m, 2) host coordinate array inside a rank that holds only the rows of its strip."""

    def __init__(self, strip, part):
        self._strip, self._part = strip, part
        self.shape = (part.n * part.m, 2)
        self.dtype = strip.dtype

    def __getitem__(self, sl):
        p = self._part
        if not (isinstance(sl, slice) and sl.start == p.e0 * p.m and sl.stop == p.e1 * p.m):
            raise IndexError("this rank holds the rows of its own strip only")
        return self._strip

    def __array__(self, *a, **k):
        raise TypeError("a strip window cannot be materialised as the global array")


if __name__ == "__main__":
    main()
