"""Time to a 1e-6 TRUE relative residual with the multigrid-preconditioned CG on the BASELINE configs that fit one GPU,
from host arrays (assembly + hierarchy + solve), for FI_F32 (fp32 outer CG, widened to fp64 when it stalls) and FI_F64."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import field_interpolation_b200 as fi
from field_interpolation_b200 import workloads as W

which = sys.argv[1:] or ["C3", "C4", "512"]
for name in which:
    if name == "C3":
        sizes, cloud = [2048, 2048], W.circles_2d(200_000, seed=0)
    elif name == "C4":
        sizes, cloud = [256] * 3, W.sphere_torus_3d(1_000_000, seed=0)
    else:
        sizes, cloud = [512] * 3, W.sphere_torus_3d(1_000_000, seed=0)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    for prec, pname in ((fi.FI_F32, "f32"), (fi.FI_F64, "f64")):
        for rep in range(2):
            t0 = time.perf_counter()
            f = fi.sdf_from_points(sizes, fi.Weights(), pos, cloud["normals"])
            x, st = f.solve(fi.solve_options(prec, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID))
            dt = time.perf_counter() - t0
            f.close()
            print(json.dumps(dict(config=name, precision=pname, rep=rep, seconds=round(dt, 4), **st)), flush=True)
