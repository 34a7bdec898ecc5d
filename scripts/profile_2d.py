"""For ncu: BASELINE configs[2] (2048^2 from 200k oriented points) — assemble, a few Jacobi-PCG iterations with the 2D TMA
kernel, two multigrid-PCG iterations (2D epilogue kernel)."""
import sys

import torch

sys.path.insert(0, ".")
import field_interpolation_b200 as fi
from field_interpolation_b200 import workloads as W

sizes = [2048, 2048]
cloud = W.circles_2d(200_000, seed=0)
pos = torch.from_numpy(W.to_lattice(cloud["unit_pos"], sizes)).cuda()
nrm = torch.from_numpy(cloud["normals"]).cuda()
f = fi.sdf_from_points(sizes, fi.Weights(), pos, nrm)
out = torch.empty(sizes[0] * sizes[1], device="cuda")
_, st = f.solve(fi.solve_options(fi.FI_F32, 12, 1e-30, check_every=12), out=out)
print(st)
_, st = f.solve(fi.solve_options(fi.FI_F64, 2, 1e-30, preconditioner=fi.FI_PRECOND_MULTIGRID), out=out)
print(st)
