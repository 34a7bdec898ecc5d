"""Per-kernel device time of the PCG iteration on an nx x ny x nz lattice (what one rank of a slab-sharded solve
computes, without the communication)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import field_interpolation_b200 as fi
from field_interpolation_b200 import workloads as W

shapes = [[int(a) for a in s.split("x")] for s in sys.argv[1].split(",")]
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
for sizes in shapes:
    cloud = W.sphere_torus_3d(npts, seed=0)
    pos = torch.from_numpy(W.to_lattice(cloud["unit_pos"], sizes)).cuda()
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, torch.from_numpy(cloud["normals"]).cuda())
    for prec, name, B in ((fi.FI_F32, "f32", 4),):
        opt = fi.solve_options(prec, 0, 1e-6)
        f.time_iterations(20, opt)
        t = f.time_iterations(iters, opt)
        cells = sizes[0] * sizes[1] * sizes[2]
        per = {k: round(v / iters, 4) for k, v in t.items() if k.endswith("_ms")}
        print(json.dumps({"sizes": sizes, "prec": name, **per, "ideal_iter_ms_48B": round(48 * cells / 6.5367e12 * 1e3, 4),
                          "Gcell_iters_per_s": round(cells / (per["iteration_ms"] * 1e-3) / 1e9, 2)}), flush=True)
    f.close()
