"""For ncu (--profile-from-start off): two multigrid-PCG iterations on the 3D SDF workload, after the hierarchy and the
V-cycle graph have been built by an untimed first call."""
import sys

import torch

sys.path.insert(0, ".")
import field_interpolation_b200 as fi
from field_interpolation_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prec = fi.FI_F64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else fi.FI_F32
cloud = W.sphere_torus_3d(1_000_000, seed=0)
pos = torch.from_numpy(W.to_lattice(cloud["unit_pos"], [n] * 3)).cuda()
nrm = torch.from_numpy(cloud["normals"]).cuda()
f = fi.sdf_from_points([n] * 3, fi.Weights(), pos, nrm)
out = torch.empty(n**3, device="cuda")
opt = fi.solve_options(prec, iters, 1e-30, preconditioner=fi.FI_PRECOND_MULTIGRID)
_, st = f.solve(opt, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
_, st = f.solve(opt, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(st)
