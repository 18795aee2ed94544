"""Strip solve over all ranks with the true-residual evaluation / continuation, checked from scratch (bench.true_residual: the
gathered solution under the matrix of the whole plate on rank 0).
usage: torchrun --nproc-per-node N scripts/dist_polish_probe.py [cols] [cells]"""
import os
import sys
sys.path.insert(0, '.')
import datetime
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
prob = StripProblem(L=cols * 1e-3, l=cells * 1e-3, N=cols, M=cells, rank=rank, world=world, E=bench.E, nu=bench.NU, t=bench.T).setup(lattice_only=True)
for polish in (-1, 3):
    prob.start()
    torch.cuda.synchronize()
    info = prob.solve(1e-8, polish=polish)
    tr = bench.true_residual(StripProblem, prob, cols, cells, world, rank)
    if rank == 0:
        print(f"{cols}x{cells} on {world} GPUs, polish={polish}: {info}  independent true relres {tr:.3e}", flush=True)
prob.release()
dist.barrier()
dist.destroy_process_group()
