"""Multi-GPU parity check (run under torchrun on >= 2 GPUs): the z-slab sharded solve against the single-GPU
solve of the same system, iterate for iterate.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/slab_check.py [--mg-only]

Second half: multigrid-preconditioned CG through the sharded V-cycle against the single-GPU V-cycle (same iteration
count within rounding, same field).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import field_interpolation_b200 as fi
from field_interpolation_b200 import dist as fid
from field_interpolation_b200 import workloads as W

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
QUICK = "--quick" in sys.argv  # the small cases only (tests/test_gpu_zz_multi_gpu.py)
PCG_CASES = () if "--mg-only" in sys.argv else (([64, 48, 40], 4000, {}), ([128, 64, 37], 20000, dict(model_1=0.3)), ([96, 40, 64], 8000, dict(model_2=0.0, model_4=0.2)),
                            ([256, 256, 256], 1000000, {}))
if QUICK:
    PCG_CASES = PCG_CASES[:3]
for sizes, npts, orders in PCG_CASES:
    cloud = W.sphere_torus_3d(npts, seed=1)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    weights = fi.Weights(**orders)
    d_pos, d_nrm = torch.from_numpy(pos).cuda(), torch.from_numpy(cloud["normals"]).cuda()
    # the largest case also runs on the cost-balanced partition (fi_slab_balanced_cuts); its gather below uses the cuts
    cuts = fid.balanced_cuts(sizes, world, d_pos, 60.0 if QUICK else 0.0, 4) if (sizes[2] >= 256 or (QUICK and sizes[2] == 40)) and "--uniform" not in sys.argv else None
    runner = fid.SlabRunner(sizes, weights, rank, world, dist, cuts=cuts)
    own = (lambda r: (cuts[r], cuts[r + 1])) if cuts else (lambda r: fid.slab_range(sizes[2], world, r))
    for prec, name, tol in ((fi.FI_F32, "f32", 2e-4), (fi.FI_F64, "f64", 1e-10)):
        for its in (1, 2, 25):
            opt = fi.solve_options(prec, its, 1e-30, check_every=8)
            out = torch.zeros(runner.local_cells, device="cuda")
            st = runner.step(d_pos, d_nrm, opt, out)
            parts = [torch.zeros((own(r)[1] - own(r)[0]) * sizes[0] * sizes[1], device="cuda") for r in range(world)]
            dist.all_gather(parts, out)
            full = torch.cat(parts).cpu().numpy()
            if rank == 0:
                f = fi.sdf_from_points(sizes, weights, d_pos, d_nrm)
                ref, st1 = f.solve(opt, out=torch.zeros(int(np.prod(sizes)), device="cuda"))
                ref = ref.cpu().numpy()
                err = float(np.linalg.norm(full - ref) / max(np.linalg.norm(ref), 1e-300))
                good = err <= tol and st["iterations"] == st1["iterations"] == its and abs(st["relative_residual"] - st1["relative_residual"]) <= 1e-3 * st1["relative_residual"] + 1e-12
                ok = ok and good
                print(json.dumps({"sizes": sizes, "cuts": cuts, "prec": name, "its": its, "rel_diff_vs_single": err, "relres_slab": st["relative_residual"],
                                  "relres_single": st1["relative_residual"], "true_slab": st["true_residual"], "true_single": st1["true_residual"], "ok": bool(good)}), flush=True)
                f.close()
    runner.close()

# ---- multigrid-preconditioned CG: the sharded V-cycle against the single-GPU V-cycle (the same linear operator) --------
def gather_own(out, sizes):
    parts = [torch.zeros((fid.slab_range(sizes[2], world, r)[1] - fid.slab_range(sizes[2], world, r)[0]) * sizes[0] * sizes[1], device="cuda")
             for r in range(world)]
    dist.all_gather(parts, out)
    return torch.cat(parts).cpu().numpy()


MG_CASES = (([64, 48, 40], 5000, {}, 1000), ([128, 64, 72], 20000, dict(model_1=0.2), 0), ([128, 128, 128], 200000, {}, 100000),
            ([256, 256, 256], 1000000, {}, 0), ([256, 256, 256], 1000000, {}, 300000))
for sizes, npts, orders, gather in (MG_CASES[:2] if QUICK else MG_CASES):
    if gather:
        os.environ["FI_B200_MG_GATHER_CELLS"] = str(gather)
    else:
        os.environ.pop("FI_B200_MG_GATHER_CELLS", None)
    radius = 2
    try:
        plan = fid.slab_mg_plan(sizes, world, radius, gather)
    except Exception as e:  # slabs too thin for this many ranks: every rank skips alike
        if rank == 0:
            print(json.dumps({"mg": True, "sizes": sizes, "skipped": str(e)}), flush=True)
        continue
    cloud = W.sphere_torus_3d(npts, seed=2)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    weights = fi.Weights(**orders)
    runner = fid.SlabRunner(sizes, weights, rank, world, dist)
    d_pos, d_nrm = torch.from_numpy(pos).cuda(), torch.from_numpy(cloud["normals"]).cuda()
    for prec, name, tol, close in ((fi.FI_F64, "f64", 1e-8, 1e-5), (fi.FI_F32, "f32", 1e-6, 2e-3)):
        opt = fi.solve_options(prec, 300, tol, preconditioner=fi.FI_PRECOND_MULTIGRID)
        out = torch.zeros(runner.local_cells, device="cuda")
        st = runner.step(d_pos, d_nrm, opt, out)
        st = runner.step(d_pos, d_nrm, opt, out)  # warm (allocations, NCCL channels)
        full = gather_own(out, sizes)
        if rank == 0:
            f = fi.sdf_from_points(sizes, weights, d_pos, d_nrm)
            ref, st1 = f.solve(opt, out=torch.zeros(int(np.prod(sizes)), device="cuda"))
            ref = ref.cpu().numpy()
            err = float(np.linalg.norm(full - ref) / max(np.linalg.norm(ref), 1e-300))
            good = bool(st["converged"]) and abs(st["iterations"] - st1["iterations"]) <= 2 and err <= close
            ok = ok and good
            print(json.dumps({"mg": True, "sizes": sizes, "prec": name, "sharded_levels": plan["sharded_levels"], "iters_slab": st["iterations"],
                              "iters_single": st1["iterations"], "rel_diff_vs_single": err, "true_slab": st["true_residual"], "true_single": st1["true_residual"],
                              "solve_ms_slab": st["solve_ms"], "setup_ms_slab": st["setup_ms"], "solve_ms_single": st1["solve_ms"], "ok": good}), flush=True)
            f.close()
    runner.close()
os.environ.pop("FI_B200_MG_GATHER_CELLS", None)

flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.destroy_process_group()
if rank == 0:
    print("SLAB CHECK", "PASSED" if ok else "FAILED")
sys.exit(0 if int(flag.item()) else 1)
