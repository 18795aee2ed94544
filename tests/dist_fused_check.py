"""Multi-GPU check of the strip-partitioned PCG (run under torchrun, one rank per GPU; see
test_gpu_parity.py::test_fused_strip_pcg_multi_gpu):
  1. the peer-memory transport (collectives fused into the PCG kernels) and the NCCL transport give the same
     iterates (bit-identical for 2 ranks: the scalar sums have two addends);
  2. both match the single-GPU solver on the owned rows, and the assembled, reduced matrix rows a rank owns are
     the single-GPU rows bit for bit;
  3. a restarted solve (sequence numbers keep growing) converges to the requested residual.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import psfem_b200  # noqa: F401
    from psfem_b200.dist import StripProblem

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, M, iters = 96 * world, 80, 40
    kw = dict(L=N * 1e-2, l=M * 1e-2, N=N, M=M, E=210e9, nu=0.3, t=0.1)
    res = {}
    for transport in ("peer", "nccl"):
        prob = StripProblem(rank=rank, world=world, transport=transport, **kw).setup()
        assert prob.transport == transport, f"transport {transport} not available: fell back to {prob.transport}"
        assert (prob.fabric is not None) == (transport == "peer")
        if transport == "peer":                                    # assembled + reduced rows this rank owns
            rp = prob.K_red.row_ptr.cpu().numpy().astype(np.int64)
            res["K_rows"] = (rp - rp[0], prob.K_red.vals[int(rp[0]):int(rp[-1])].cpu().numpy().copy(),
                             prob.K_red.col_idx[int(rp[0]):int(rp[-1])].cpu().numpy().astype(np.int64) - prob.plan["own"][0])
        prob.cg_iterations(iters)
        r, b = prob.residual()
        res[transport] = (prob.st["x"].cpu().numpy().copy(), r, b)
        if transport == "peer":                                    # restart on the same boards, solve to 1e-9
            info = prob.full_solve(1e-9)
            assert info["relres"] <= 1e-9, info
            res["solve_iters"] = info["iterations"]
        lo_hi = (prob.part.dof_shift, prob.plan["own"])
        prob.release()
    xp, rp, bp = res["peer"]
    xn, rn, bn = res["nccl"]
    scale = np.abs(xn).max()
    if world == 2:
        assert np.array_equal(xp, xn), np.abs(xp - xn).max() / scale
        assert rp == rn and bp == bn
    else:
        assert np.abs(xp - xn).max() <= 1e-9 * scale
        assert abs(rp - rn) <= 1e-6 * rn
    # single-GPU solver on the whole plate (every rank repeats it; small)
    one = StripProblem(rank=0, world=1, **kw).setup()
    one.cg_iterations(iters)
    x1 = one.st["x"].cpu().numpy()
    r1, b1 = one.residual()
    # owned free DOFs of this rank inside the global free vector: rank 0 starts at 0, sizes are gathered
    sizes = [None] * world
    dist.all_gather_object(sizes, len(xp))
    off = sum(sizes[:rank])
    assert sum(sizes) == len(x1)
    # the strip's owned rows of K are the single-GPU rows BIT FOR BIT (pattern, values; columns up to the strip offset)
    rp1 = one.K_red.row_ptr.cpu().numpy().astype(np.int64)
    a, b_ = int(rp1[off]), int(rp1[off + len(xp)])
    rel_ptr, vals, cols = res["K_rows"]
    assert np.array_equal(rel_ptr, rp1[off:off + len(xp) + 1] - a)
    assert np.array_equal(vals, one.K_red.vals[a:b_].cpu().numpy())
    assert np.array_equal(cols + off, one.K_red.col_idx[a:b_].cpu().numpy().astype(np.int64))
    assert np.abs(xp - x1[off:off + len(xp)]).max() <= 1e-8 * np.abs(x1).max()
    assert abs(rp - r1) <= 1e-5 * r1 and abs(bp - b1) <= 1e-12 * b1
    dist.barrier()
    if rank == 0:
        print(f"dist_fused_check OK: world={world} iters={iters} relres={rp / bp:.3e} full-solve iterations={res['solve_iters']}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
