"""Where one bench step spends its time (FI_B200_TRACE=1 for the library's own phase lines)."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import field_interpolation_b200 as fi
from field_interpolation_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
cloud = W.sphere_torus_3d(1_000_000, seed=0)
d_pos = torch.from_numpy(W.to_lattice(cloud["unit_pos"], [n] * 3)).cuda()
d_nrm = torch.from_numpy(cloud["normals"]).cuda()
runner = None
if world > 1:
    import torch.distributed as dist
    from field_interpolation_b200 import dist as fid
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    runner = fid.SlabRunner([n] * 3, fi.Weights(), rank, world, dist)
out = torch.empty(n**3 if runner is None else runner.local_cells, device="cuda")
for iters in (100, 100, 100, 400, 800):
    opt = fi.solve_options(fi.FI_F32, iters, 1e-30, check_every=100)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if runner is None:
        f = fi.sdf_from_points([n] * 3, fi.Weights(), d_pos, d_nrm)
        t1 = time.perf_counter()
        _, st = f.solve(opt, out=out)
        t2 = time.perf_counter()
        f.close()
    else:
        t1 = t0
        st = runner.step(d_pos, d_nrm, opt, out)
        t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if rank == 0:
        print(f"iters {iters}: total {1e3*(t3-t0):.1f} ms = assemble {1e3*(t1-t0):.1f} + solve call {1e3*(t2-t1):.1f} + close {1e3*(t3-t2):.1f};"
              f" lib setup_ms {st['setup_ms']:.1f} solve_ms {st['solve_ms']:.1f}  per-iter {st['solve_ms']/iters:.4f}", flush=True)
