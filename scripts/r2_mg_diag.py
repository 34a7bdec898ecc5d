"""Round-2 diagnosis of the fp32 multigrid-PCG failure on BASELINE config C3 (2048^2, 200k oriented points).

Prints one JSON line per experiment: the fp32 solve (FI_F32 + FI_PRECOND_MULTIGRID) against the fp64-outer solve of
the same system, for several iteration caps / smoothing steps, twice each (run-to-run spread)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import field_interpolation_b200 as fi  # noqa: E402
from field_interpolation_b200 import workloads as W  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "C3"
    if which == "C3":
        sizes = [2048, 2048]
        cloud = W.circles_2d(200_000, seed=0)
    elif which == "C4":
        sizes = [256, 256, 256]
        cloud = W.sphere_torus_3d(1_000_000, seed=0)
    else:
        sizes = [512, 512, 512]
        cloud = W.sphere_torus_3d(1_000_000, seed=0)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, cloud["normals"])
    tag = dict(which=which, data_term=os.environ.get("FI_B200_DATA_TERM", "default"))
    t0 = time.time()
    x64, st64 = f.solve(fi.solve_options(fi.FI_F64, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID))
    print(json.dumps(dict(tag, exp="f64_mg", wall=time.time() - t0, **st64)), flush=True)
    x64b, st64b = f.solve(fi.solve_options(fi.FI_F64, 500, 1e-9, preconditioner=fi.FI_PRECOND_MULTIGRID))
    print(json.dumps(dict(tag, exp="f64_mg_1e-9", rel_vs_1e6=rel(x64, x64b), **st64b)), flush=True)
    for nu in (0, 4, 6):
        for cap in (5, 10, 20, 40, 80, 500):
            for rep in range(2 if cap == 500 else 1):
                x32, st32 = f.solve(fi.solve_options(fi.FI_F32, cap, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID, mg_smoothing_steps=nu))
                print(json.dumps(dict(tag, exp="f32_mg", nu=nu, cap=cap, rep=rep, rel_vs_f64=rel(x32, x64b), xmin=float(x32.min()),
                                      xmax=float(x32.max()), **st32)), flush=True)
    xm, stm = f.solve(fi.solve_options(fi.FI_MIXED, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID))
    print(json.dumps(dict(tag, exp="mixed_mg", rel_vs_f64=rel(xm, x64b), **stm)), flush=True)
    # Jacobi fp32 from the multigrid solution: does plain fp32 PCG hold the fixed point?
    xj, stj = f.solve(fi.solve_options(fi.FI_F32, 200, 1e-6), guess=x64b)
    print(json.dumps(dict(tag, exp="f32_jacobi_from_x64", rel_vs_f64=rel(xj, x64b), **stj)), flush=True)
    f.close()


if __name__ == "__main__":
    main()
