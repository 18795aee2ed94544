"""torchrun script: per-kernel times of the fused strip PCG (psfem_dist_pcg_profile) on every rank."""
import ctypes as C, os, sys
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import psfem_b200
from psfem_b200._capi import call, stream_ptr
from psfem_b200.dist import StripProblem
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prob = StripProblem(L=cells * world * 1e-3, l=cells * 1e-3, N=cells * world, M=cells, rank=rank, world=world, transport="peer").setup()
prob.cg_iterations(10)
for rep in range(3):
    h = (C.c_double * 3)()
    dist.barrier(); torch.cuda.synchronize()
    call("psfem_dist_pcg_profile", C.byref(prob.A), C.byref(prob.ctx), 25, prob.seq + 1, h, stream_ptr())
    prob.seq += 25
    print(f"rank {rank} cells {cells}: spmv+dot {h[0]:.4f} update {h[1]:.4f} pupdate {h[2]:.4f} sum {sum(h):.4f} ms", flush=True)
# the single-GPU vector kernels on the same buffers, in this multi-GPU process (no peer waits): what the kernels cost alone
st = prob.st
lo, hi = prob.plan["own"]
rz = torch.ones(2, dtype=torch.float64, device="cuda"); pq = torch.ones(1, dtype=torch.float64, device="cuda"); out2 = torch.zeros(2, dtype=torch.float64, device="cuda")
for name, fn in (("pcg_update (r,q,dinv -> r)", lambda: prob.ops.update(st["r"], st["q"], st["dinv"], rz[0:1], pq, out2)),
                 ("pcg_pupdate (p,x,r,dinv -> p,x)", lambda: prob.ops.pupdate(st["p_ext"][lo:hi], st["x"], st["r"], st["dinv"], rz[0:1], rz[1:2], pq)),
                 ("spmv+dot alone", lambda: prob.ops.spmv_dot(st["p_ext"], lo, st["q"], pq))):
    ts = []
    for rep in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"rank {rank} alone {name}: min {min(ts[2:]):.4f} median {sorted(ts[2:])[5]:.4f} ms", flush=True)
# streaming bandwidth of the peer-visible allocation against ordinary device memory (same kernel, same size)
from psfem_b200._capi import ptr
n = prob.n_owned_free
lo, hi = prob.plan["own"]
a_peer = prob.st["p_ext"][lo:hi]
a_loc = torch.empty(n, dtype=torch.float64, device="cuda").copy_(a_peer)
out = torch.empty(n, dtype=torch.float64, device="cuda")
for name, src, dst in (("local->local", a_loc, out), ("peer->local", a_peer, out), ("local->peer", a_loc, a_peer)):
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call("psfem_axpby", n, 1.0, ptr(src), 0.0, ptr(None), ptr(dst), stream_ptr())
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"rank {rank} copy {name}: {min(ts):.4f} ms = {16 * n / min(ts) / 1e6:.0f} GB/s", flush=True)
prob.release()
dist.destroy_process_group()
