"""Multigrid tuning sweep (B200): iterations and device time of the fp64-outer MG-PCG to a 1e-6 true residual for
smoothing steps on the finest level / on the coarse levels, V- vs W-cycle, and the Chebyshev interval ratio."""
import itertools
import json
import os
import sys

import torch

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
    pos = torch.from_numpy(W.to_lattice(cloud["unit_pos"], sizes)).cuda()
    nrm = torch.from_numpy(cloud["normals"]).cuda()
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, nrm)
    out = torch.empty(int(torch.tensor(sizes).prod()), device="cuda")
    combos = [(3, 0, 1, 99, 12.0), (2, 4, 1, 99, 12.0), (1, 4, 1, 99, 12.0)]
    combos += [(nu, nuc, 2, wl, 12.0) for nu, nuc, wl in itertools.product((2, 3, 4), (5, 6, 8), (1, 2, 3))]
    combos += [(3, 6, 2, 2, r) for r in (8.0, 20.0)] + [(3, 4, 2, 2, 12.0), (3, 10, 2, 2, 12.0), (3, 12, 2, 1, 12.0), (4, 6, 2, 99, 12.0)]
    for nu, nuc, gamma, wl, ratio in combos:
        os.environ["FI_B200_MG_NU_COARSE"] = str(nuc)
        os.environ["FI_B200_MG_GAMMA"] = str(gamma)
        os.environ["FI_B200_MG_WLEVELS"] = str(wl)
        opt = fi.solve_options(fi.FI_F64, 200, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID, mg_smoothing_steps=nu, mg_cheb_ratio=ratio)
        try:
            _, st = f.solve(opt, out=out)  # builds the hierarchy
            _, st = f.solve(opt, out=out)
            print(json.dumps(dict(config=name, nu=nu, nu_coarse=nuc, gamma=gamma, wlevels=wl, ratio=ratio, iterations=st["iterations"], solve_ms=round(st["solve_ms"], 2),
                                  ms_per_iteration=round(st["solve_ms"] / max(1, st["iterations"]), 3), converged=st["converged"], cycle_at_end=st["outer_sweeps"], true_residual=st["true_residual"])), flush=True)
        except fi.FiError as e:
            print(json.dumps(dict(config=name, nu=nu, nu_coarse=nuc, gamma=gamma, wlevels=wl, ratio=ratio, error=str(e))), flush=True)
    f.close()
