"""One short pass of the hot path for ncu: assemble the 3D SDF system and run a few PCG iterations."""
import sys

import torch

sys.path.insert(0, ".")
import field_interpolation_b200 as fi
from field_interpolation_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
prec = fi.FI_F64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else fi.FI_F32
cloud = W.sphere_torus_3d(1_000_000, seed=0)
pos = torch.from_numpy(W.to_lattice(cloud["unit_pos"], [n] * 3)).cuda()
nrm = torch.from_numpy(cloud["normals"]).cuda()
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1  # > 1: the later steps are warm (FI_B200_TRACE=1 prints the phases of each)
out = torch.empty(n**3, device="cuda")
for rep in range(reps):
    f = fi.sdf_from_points([n] * 3, fi.Weights(), pos, nrm)
    _, st = f.solve(fi.solve_options(prec, iters, 1e-30, check_every=min(iters, 100)), out=out)
    f.close()
    print("step", rep, st, flush=True)
