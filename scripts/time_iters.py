"""Exploration: per-iteration and per-apply device time of the PCG kernels on the 3D SDF workload."""
import json
import sys

import torch

sys.path.insert(0, ".")
import field_interpolation_b200 as fi
from field_interpolation_b200 import workloads as W

sizes_list = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [256, 512]
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
cloud = W.sphere_torus_3d(npts, seed=0)
up = torch.from_numpy(cloud["unit_pos"]).cuda()
nr = torch.from_numpy(cloud["normals"]).cuda()
for n in sizes_list:
    f = fi.sdf_from_points([n] * 3, fi.Weights(), up * (n - 1.0), nr)
    for prec, name, B in ((fi.FI_F32, "f32", 4), (fi.FI_F64, "f64", 8)):
        for fast in (True, False):
            if not fast and n > 256:
                continue
            opt = fi.solve_options(prec, 0, 1e-6, use_fast_stencil=fast)
            f.time_iterations(20, opt)
            t = f.time_iterations(iters, opt)
            cells = n**3
            per = {k: round(v / iters, 4) for k, v in t.items() if k.endswith("_ms")}
            print(json.dumps({"n": n, "prec": name, "fast": fast, "fused": t["fused"], **per,
                              "iter_GBs_52B": round(13 * B * cells / (per["iteration_ms"] * 1e-3) / 1e9, 1),
                              "apply_GBs": round((5 if t["fused"] else 2) * B * cells / (per["apply_ms"] * 1e-3) / 1e9, 1),
                              "update_GBs": round(7 * B * cells / (per["update_ms"] * 1e-3) / 1e9, 1),
                              "Gcell_iters_per_s": round(cells / (per["iteration_ms"] * 1e-3) / 1e9, 2)}), flush=True)
    del f
