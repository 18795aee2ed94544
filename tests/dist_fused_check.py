"""Multi-GPU check of the strip-partitioned PCG (run under torchrun, one rank per GPU; see
test_gpu_parity.py::test_fused_strip_pcg_multi_gpu).  The checks themselves live in the package
(psfem_b200.dist_check.run_checks) so that bench.py --gpus N runs them too before it times anything."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import psfem_b200  # noqa: F401
    from psfem_b200.dist_check import run_checks

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=150))
    out = run_checks(world, rank)
    if rank == 0:
        print(f"dist_fused_check OK: world={world} iters={out['iters']} relres={out['relres']:.3e} "
              f"full-solve iterations={out['full_solve_iterations']}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
