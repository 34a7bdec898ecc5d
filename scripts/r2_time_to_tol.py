"""Time to a 1e-6 TRUE relative residual with the multigrid-preconditioned CG on the BASELINE configs that fit one GPU,
from host arrays (assembly + hierarchy + solve), for FI_F32 (fp32 outer CG, widened to fp64 when it stalls) and FI_F64."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import field_interpolation_b200 as fi
from field_interpolation_b200 import workloads as W

import os

which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["C3", "C4", "512"]
# --tail=0,4096,...: repeat every case with FI_B200_MG_TAIL_CELLS set to each value (the hierarchy reads it per solve)
tails = next((a.split("=", 1)[1].split(",") for a in sys.argv[1:] if a.startswith("--tail=")), [None])
for name in which:
    if name == "C2":
        sizes, cloud = [512, 512], W.circles_2d(10_000, seed=0)
    elif name == "C3":
        sizes, cloud = [2048, 2048], W.circles_2d(200_000, seed=0)
    elif name == "C4":
        sizes, cloud = [256] * 3, W.sphere_torus_3d(1_000_000, seed=0)
    else:
        sizes, cloud = [512] * 3, W.sphere_torus_3d(1_000_000, seed=0)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    for tail in tails:
        if tail is not None:
            os.environ["FI_B200_MG_TAIL_CELLS"] = tail
        for prec, pname in ((fi.FI_F32, "f32"), (fi.FI_F64, "f64")):
            for rep in range(3 if tail is not None else 2):
                t0 = time.perf_counter()
                f = fi.sdf_from_points(sizes, fi.Weights(), pos, cloud["normals"])
                x, st = f.solve(fi.solve_options(prec, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID))
                dt = time.perf_counter() - t0
                f.close()
                print(json.dumps(dict(config=name, precision=pname, tail_cells=tail, rep=rep, seconds=round(dt, 4), **st)), flush=True)
