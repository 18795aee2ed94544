"""Strip solve over all ranks with the true-residual evaluation / continuation, checked from scratch (bench.true_residual: the
gathered solution under the matrix of the whole plate on rank 0).
usage: torchrun --nproc-per-node N scripts/dist_polish_probe.py [cols] [cells]
env: PROBE_POLISH (list, default -1,3)  PROBE_LATTICE (1)  PROBE_MAXIT (cut the solve short: no residual check)
     PROBE_REPEAT (set-up -> restart -> solve that many times)  PROBE_STAGES (staged warm-up with state reads)"""
import os
import sys
sys.path.insert(0, '.')
import datetime
import time
import torch
import torch.distributed as dist
import psfem_b200  # noqa: F401
from psfem_b200.dist import StripProblem
import bench

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=150))
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
maxit = int(os.environ.get("PROBE_MAXIT", "0"))


def stage(prob, name, fn):
    t0 = time.perf_counter()
    try:
        fn()
        torch.cuda.synchronize()
        msg = f"{prob.residual()} iters {prob.st['iters']}"
    except Exception as e:                                  # noqa: BLE001
        msg = f"FAILED {e}"
    print(f"rank {rank}: {name}: {msg} ({time.perf_counter() - t0:.2f} s)", flush=True)


for rep in range(int(os.environ.get("PROBE_REPEAT", "1"))):
    prob = StripProblem(L=cols * 1e-3, l=cells * 1e-3, N=cols, M=cells, rank=rank, world=world, E=bench.E, nu=bench.NU, t=bench.T)
    prob.setup(lattice_only=os.environ.get("PROBE_LATTICE", "1") == "1")
    if os.environ.get("PROBE_STAGES"):
        stage(prob, "after set-up", lambda: None)
        stage(prob, "50 iterations", lambda: prob.cg_iterations(50))
        stage(prob, "1000 iterations", lambda: [prob.cg_iterations(100) for _ in range(10)])
        stage(prob, "restart", prob.start)
        stage(prob, "100 iterations", lambda: prob.cg_iterations(100))
        stage(prob, "2000 iterations with state reads", lambda: [(prob.cg_iterations(100), prob.residual()) for _ in range(20)])
    for polish in [int(v) for v in os.environ.get('PROBE_POLISH', '-1,3').split(',')]:
        try:
            prob.start()
            torch.cuda.synchronize()
            info = prob.solve(1e-8, polish=polish, **(dict(maxit=maxit) if maxit else {}))
        except Exception as e:                              # noqa: BLE001
            print(f"rank {rank}: repetition {rep}, polish={polish}: FAILED after {prob.st['iters']} iterations (seq {prob.seq}): {e}", flush=True)
            raise
        tr = float("nan") if maxit else bench.true_residual(StripProblem, prob, cols, cells, world, rank)
        if rank == 0:
            print(f"{cols}x{cells} on {world} GPUs, repetition {rep}, polish={polish}: {info}  independent true relres {tr:.3e}", flush=True)
    prob.release()
    del prob
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
