"""Exploration (not part of the product): PCG / cascade convergence of the 3D SDF workload on one GPU."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import field_interpolation_b200 as fi
from field_interpolation_b200 import workloads as W

sizes_list = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 128, 256]
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
cloud = W.sphere_torus_3d(npts, seed=0)
up = torch.from_numpy(cloud["unit_pos"]).cuda()
nr = torch.from_numpy(cloud["normals"]).cuda()
for n in sizes_list:
    sizes = [n, n, n]
    for prec, name in ((fi.FI_F32, "f32"), (fi.FI_F64, "f64"), (fi.FI_MIXED, "mixed")):
        for factor, coarse_tol in ((2, 1e-6), (2, 1e-3)):
            torch.cuda.synchronize()
            t = time.time()
            x, st = fi.sdf_solve_cascade(sizes, fi.Weights(), up, nr, options=fi.solve_options(prec, 60000, 1e-6, check_every=64),
                                         factor=factor, coarsest_size=16, coarse_tolerance=coarse_tol)
            torch.cuda.synchronize()
            dt = time.time() - t
            fin = st["finest"]
            print(json.dumps({"n": n, "prec": name, "factor": factor, "coarse_tol": coarse_tol, "wall_s": round(dt, 3),
                              "level_iters": st["level_iterations"], "level_ms": [round(m, 1) for m in st["level_ms"]],
                              "init_res": [float(f"{r:.3g}") for r in st["level_initial_residual"]],
                              "relres": fin["relative_residual"], "true_res": fin["true_residual"], "conv": fin["converged"],
                              "cell_iters_per_s": st["cell_iterations"] / (st["total_ms"] * 1e-3)}), flush=True)
    # zero-guess baseline at this size (fp32)
    pos = up * (n - 1.0)
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, nr)
    x, st = f.solve(fi.solve_options(fi.FI_F32, 60000, 1e-6, check_every=64), guess=None, out=torch.empty(n**3, device="cuda"))
    print(json.dumps({"n": n, "zero_guess_f32": st}), flush=True)
    del f
