"""Exploration: multigrid-preconditioned CG vs Jacobi-PCG on the 3D SDF workload (iterations, time, residuals)."""
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
variants = [(3, 12.0), (2, 12.0), (4, 12.0), (3, 30.0), (3, 6.0)] if len(sys.argv) <= 3 else [tuple(float(x) for x in v.split(":")) for v in sys.argv[3].split(",")]
cloud = W.sphere_torus_3d(npts, seed=0)
up = torch.from_numpy(cloud["unit_pos"]).cuda()
nr = torch.from_numpy(cloud["normals"]).cuda()
for n in sizes_list:
    f = fi.sdf_from_points([n] * 3, fi.Weights(), up * (n - 1.0), nr)
    out = torch.empty(n**3, device="cuda")
    for prec, pname in ((fi.FI_F32, "f32"), (fi.FI_F64, "f64")):
        for nu, ratio in variants:
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                x, st = f.solve(fi.solve_options(prec, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID, mg_smoothing_steps=int(nu), mg_cheb_ratio=ratio), out=out)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            print(json.dumps({"n": n, "prec": pname, "nu": nu, "ratio": ratio, "iters": st["iterations"], "wall_s_warm": round(dt, 4), "solve_ms": round(st["solve_ms"], 2),
                              "ms_per_iter": round(st["solve_ms"] / max(1, st["iterations"]), 3), "relres": st["relative_residual"], "true_res": st["true_residual"],
                              "conv": st["converged"]}), flush=True)
    f.close()
