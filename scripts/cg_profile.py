"""A few fused PCG iterations at benchmark scale for ncu (scripts/cg_probe.py is the timing probe)."""
import os, sys
sys.path.insert(0, '.')
import torch
import psfem_b200
from psfem_b200.dist import StripProblem
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
os.environ["PSFEM_PCG"] = sys.argv[2] if len(sys.argv) > 2 else "fused"
prob = StripProblem(L=cells * 1e-3, l=cells * 1e-3, N=cells, M=cells, rank=0, world=1).setup()
prob.cg_iterations(12)
torch.cuda.synchronize()
print("ok", prob.residual())
