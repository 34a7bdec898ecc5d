"""One GPU: the slab-sharded multigrid path on a 1-rank communicator against the plain single-GPU multigrid solve, at
bench sizes (timing of the unfused sharded V-cycle next to the fused single-GPU one; same iteration counts expected).

    python scripts/slab_mg_one_rank.py 256,512 [gather_cells]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import field_interpolation_b200 as fi
from field_interpolation_b200 import dist as fid
from field_interpolation_b200 import workloads as W


class OneRank:
    @staticmethod
    def get_backend():
        return "gloo"

    @staticmethod
    def broadcast(t, src=0):
        return None


ns = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256").split(",")]
if len(sys.argv) > 2:
    os.environ["FI_B200_MG_GATHER_CELLS"] = sys.argv[2]
for n in ns:
    sizes = [n, n, n]
    cloud = W.sphere_torus_3d(1_000_000, seed=0)
    d_pos = torch.from_numpy(W.to_lattice(cloud["unit_pos"], sizes)).cuda()
    d_nrm = torch.from_numpy(cloud["normals"]).cuda()
    weights = fi.Weights()
    runner = fid.SlabRunner(sizes, weights, 0, 1, OneRank)
    plan = fid.slab_mg_plan(sizes, 1, 2, int(os.environ.get("FI_B200_MG_GATHER_CELLS", "0")))
    for prec, name in ((fi.FI_F32, "f32"), (fi.FI_F64, "f64")):
        opt = fi.solve_options(prec, 300, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID)
        out = torch.zeros(runner.local_cells, device="cuda")
        for rep in range(2):
            st = runner.step(d_pos, d_nrm, opt, out)
        f = fi.sdf_from_points(sizes, weights, d_pos, d_nrm)
        ref = torch.zeros(n ** 3, device="cuda")
        for rep in range(2):
            _, st1 = f.solve(opt, out=ref)
        f.close()
        err = float((out - ref).norm() / ref.norm())
        print(json.dumps({"n": n, "prec": name, "sharded_levels": plan["sharded_levels"], "iters_slab": st["iterations"], "iters_plain": st1["iterations"],
                          "solve_ms_slab": st["solve_ms"], "solve_ms_plain": st1["solve_ms"], "setup_ms_slab": st["setup_ms"], "setup_ms_plain": st1["setup_ms"],
                          "true_slab": st["true_residual"], "true_plain": st1["true_residual"], "rel_diff": err}), flush=True)
        del out, ref
    runner.close()
    fi.trim_memory() if hasattr(fi, "trim_memory") else None
