"""How small can the true residual of an fp64 displacement vector be?  Solve the clamped plate to rtol 1e-8, then measure
|| K (U o delta) || / || F || for a random relative perturbation delta of half an ulp of every entry: the residual that
ROUNDING U to fp64 alone produces.  usage: python scripts/residual_floor_probe.py [cells] [cols]"""
import sys
sys.path.insert(0, '.')
import torch
import psfem_b200
from psfem_b200.dist import StripProblem
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cols = int(sys.argv[2]) if len(sys.argv) > 2 else cells
prob = StripProblem(L=cols * 1e-3, l=cells * 1e-3, N=cols, M=cells, rank=0, world=1).setup(lattice_only=True)
info = prob.full_solve(1e-8)
prob.finish_x()
x = prob.st["x"]
bn = torch.linalg.vector_norm(prob.b_own).item()
true_r = torch.linalg.vector_norm(prob.b_own - prob.K_red.matvec(x)).item() / bn
g = torch.Generator(device="cuda"); g.manual_seed(1)
delta = (torch.rand(x.shape, generator=g, device="cuda", dtype=torch.float64) - 0.5) * 2.0 ** -52
floor = torch.linalg.vector_norm(prob.K_red.matvec(x * delta)).item() / bn
print(info)
print(f"{cols}x{cells}: {info['iterations']} iterations, recurrence relres {info['relres']:.3e}, true relres {true_r:.3e}, "
      f"||K (U o delta)|| / ||F|| for |delta| <= 2^-53: {floor:.3e}, max|U| {x.abs().max().item():.3e}")
