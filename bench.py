#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on the B200 path, and the reference's CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Metric: lattice cells x CG iterations per second on the 3D SDF workload (sdf_from_points from a synthetic
sphere+torus cloud, default Weights), whole job.  A *step* is one pass of the hot path over the cloud:
assemble the normal equations from the points (already resident in HBM) and run `--iters` Jacobi-PCG
iterations on the lattice — a fixed, stated number of iterations per step, because a full solve to 1e-6 is
tens of thousands of iterations (reported separately by --time-to-tol).  `e2e` is the same step through the
public API with HOST buffers: host->device copy of points and normals and device->host copy of the field
inside the timed region.

Under torchrun (--gpus N > 1) the lattice is z-slab sharded over the ranks (strong scaling, the work is the
same lattice); rank 0 prints the line, timing is the max over ranks of CUDA-event time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (lattice size n (n^3), points, description)
    "sdf3d_512_1M": (512, 1_000_000, "3D sdf_from_points 512^3 lattice, 1M sphere+torus samples (north_star target)"),
    "sdf3d_256_1M": (256, 1_000_000, "3D sdf_from_points 256^3 lattice, 1M sphere+torus samples (BASELINE configs[3])"),
    "sdf3d_1024_20M": (1024, 20_000_000, "3D sdf_from_points 1024^3 lattice, 20M samples (BASELINE configs[4])"),
    "sdf3d_128_1M": (128, 1_000_000, "3D sdf_from_points 128^3 lattice, 1M samples (CPU-sized sample)"),
    "sdf3d_64_100k": (64, 100_000, "3D sdf_from_points 64^3 lattice, 100k samples (smoke)"),
}
BYTES_PER_CELL_ITER = {"f32": 52, "f64": 104}  # one Jacobi-PCG iteration, per lattice cell


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, precision):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of this workload (profiles/ncu_traffic.json, written by scripts/ncu_summary.py); None when there is none."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    t = json.load(open(path)).get(f"{workload}:{precision}")
    return None if t is None else t["dram_bytes_per_launch"]


class ClockSampler:
    """SM clock and throttle reasons sampled during the timed region (B200_PROFILING.md recipe) through NVML in a
    background thread — the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints, without
    a second process polling the driver while the kernels are being timed."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, device_index: int, period_s: float = 0.1):
        self.idx, self.period, self.sm, self.reasons, self.stop, self.thread, self.max_mhz, self.err = device_index, period_s, [], set(), False, None, None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception as e:  # no NVML: report that, never guess
            self.err = repr(e)
        return self

    def _loop(self):
        while not self.stop:
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as e:
                self.err = repr(e)
            time.sleep(self.period)

    def __exit__(self, *a):
        self.stop = True
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable: " + str(self.err)]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm)}


# --------------------------------------------------------------------------------------------------------
def cpu_reference_arm(n: int, npts: int, iters: int, steps: int, warmup: int):
    """The reference's CPU path on this box's host cores: assembly by the reference's own code when oracle/_ref
    is present (else the port), then the restated Eigen path (CSC -> AtA -> Jacobi-preconditioned BiCGSTAB,
    float) for a bounded number of iterations.  Single thread: the reference has no threading."""
    from field_interpolation_b200 import workloads as W
    from oracle import oracle as O
    ref = O.reference()
    asm, kind = (ref, "reference") if ref is not None else (O.port(), "port")
    cloud = W.sphere_torus_3d(npts, seed=0)
    sizes = [n, n, n]
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    times, asm_s, ata_s = [], 0.0, 0.0
    for step in range(warmup + steps):
        t0 = time.perf_counter()
        sys_ = asm.sdf_from_points(sizes, O.make_weights(), pos, cloud["normals"]).system()
        t1 = time.perf_counter()
        N = O.port().normal(sys_, n ** 3, "f32")
        t2 = time.perf_counter()
        x, its, err = N.bicgstab(guess=np.zeros(n ** 3, np.float32), max_iter=iters, tol=1e-30)
        t3 = time.perf_counter()
        if step >= warmup:
            times.append((t3 - t0, its))
            asm_s, ata_s = t1 - t0, t2 - t1
        del N, sys_
    total = sum(t for t, _ in times)
    its = sum(i for _, i in times)
    value = (n ** 3) * its / total
    return value, total / len(times), {
        "kind": "port" if kind == "port" else "reference(assembly)+port(Eigen path restated; Eigen not installable offline)",
        "cores": 1, "host_cores": os.cpu_count(),
        "sample": f"{n}^3 lattice, {npts} points, {iters} BiCGSTAB iterations per step after assembly ({asm_s:.2f} s) and "
                  f"CSC+AtA ({ata_s:.2f} s); iterations only: {(n ** 3) * its / max(1e-9, total - len(times) * (asm_s + ata_s)):.3e} cell-iters/s",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sdf3d_512_1M", choices=sorted(WORKLOADS))
    ap.add_argument("--iters", type=int, default=400, help="PCG iterations per step")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--cpu-sample", default="sdf3d_128_1M", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-iters", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--time-to-tol", action="store_true", help="also time the Jacobi-PCG coarse-to-fine cascade to 1e-6 (slow)")
    ap.add_argument("--no-time-to-tol", action="store_true", help="skip the multigrid time-to-1e-6 measurement")
    ap.add_argument("--mg-timeout", type=float, default=240.0, help="N>1: deadline in seconds for the sharded multigrid time-to-1e-6 section")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n, npts, desc = WORKLOADS[args.workload]
    metric, unit = "3D SDF solve: lattice cells x CG iterations per second", "cell-iters/s"

    if args.impl == "reference":
        if rank != 0:
            return
        sn, snpts, sdesc = WORKLOADS[args.cpu_sample]
        value, sec, base = cpu_reference_arm(sn, snpts, args.cpu_iters, max(1, min(args.steps, 2)), 1)
        base["value"] = value
        base["unit"] = unit
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "cpu_sample": sdesc, "precision": "f32"},
            "cpu_baseline": base, "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch
    import field_interpolation_b200 as fi
    from field_interpolation_b200 import workloads as W

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    prec = fi.FI_F32 if args.precision == "f32" else fi.FI_F64
    sizes = [n, n, n]
    cloud = W.sphere_torus_3d(npts, seed=0)
    h_pos = torch.from_numpy(W.to_lattice(cloud["unit_pos"], sizes)).pin_memory()
    h_nrm = torch.from_numpy(cloud["normals"]).pin_memory()
    d_pos, d_nrm = h_pos.cuda(non_blocking=True), h_nrm.cuda(non_blocking=True)
    N = n ** 3
    opt = fi.solve_options(prec, args.iters, 1e-30, check_every=min(args.iters, 100))
    weights = fi.Weights()

    if world > 1:
        from field_interpolation_b200 import dist as fid
        # The occupied cells of the cloud cluster in the middle slabs, and the slowest rank sets the pace of every
        # iteration: partition by cost (lattice planes + data points, fi_slab_balanced_cuts) instead of by plane count.
        # One histogram kernel over the resident points, computed once for the cloud (it fixes the size of every rank's
        # output buffer); FI_B200_BENCH_UNIFORM=1 keeps the uniform partition.
        runner = fid.SlabRunner(sizes, weights, rank, world, dist)
        if os.environ.get("FI_B200_BENCH_UNIFORM") != "1":
            try:
                runner.set_cuts(fid.balanced_cuts(sizes, world, d_pos, 0.0, 8))
            except fi.FiError:  # deterministic (same arguments on every rank): everybody stays on the uniform partition
                runner.set_cuts(None)
    else:
        runner = None

    d_out = torch.empty(N if world == 1 else runner.local_cells, dtype=torch.float32, device="cuda")
    h_out = torch.empty(d_out.numel(), dtype=torch.float32).pin_memory()

    last_stats = {}

    def step_device():
        if runner is not None:
            st = runner.step(d_pos, d_nrm, opt, d_out)
            last_stats.update(st)
            return st
        f = fi.sdf_from_points(sizes, weights, d_pos, d_nrm)
        _, st = f.solve(opt, out=d_out)
        f.close()
        return st

    def step_e2e():
        if runner is not None:
            st = runner.step(h_pos, h_nrm, opt, d_out)
            h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            return st
        f = fi.sdf_from_points(sizes, weights, h_pos.numpy(), h_nrm.numpy())
        _, st = f.solve(opt, out=h_out.numpy())
        f.close()
        return st

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        its = 0
        for _ in range(steps):
            its += fn()["iterations"]
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        # the library runs on its own stream; the bracketing synchronisations make the host wall time the safe
        # upper bound of the device time — report the larger of the two
        ms = max(ms, wall * 1e3)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, its

    for _ in range(args.warmup):
        step_device()
    fi.kernel_launches_reset()
    with ClockSampler(local_rank) as clocks:
        ms, its = timed(step_device, args.steps)
    launches = fi.kernel_launches()
    clock_summary = clocks.summary()
    value = N * its / (ms * 1e-3)

    step_e2e()
    ms_e2e, its_e2e = timed(step_e2e, args.steps)
    e2e_value = N * its_e2e / (ms_e2e * 1e-3)
    h2d = h_pos.numel() * 4 + h_nrm.numel() * 4
    d2h = N * 4  # every rank reads back its owned planes: the whole field over all ranks

    # roofline of the dominant kernel, measured live with CUDA events on the solver stream
    peak, peak_src = peaks()
    roof, extra = None, {}
    if rank == 0 and runner is None:
        f = fi.sdf_from_points(sizes, weights, d_pos, d_nrm)
        f.time_iterations(20, opt)
        t = f.time_iterations(200, opt)
        f.close()
        B = 4 if args.precision == "f32" else 8
        words = 5 if t["fused"] else 2  # fused direction+stencil: read r, M^-1, p_old, write p_new, q; plain stencil: read p, write q
        apply_s, upd_s, it_s = t["apply_ms"] / 200e3, t["update_ms"] / 200e3, t["iteration_ms"] / 200e3
        sten_s, data_s = t["stencil_ms"] / 200e3, t["data_term_ms"] / 200e3
        # dominant kernel = the lattice-sized stencil kernel, timed alone with CUDA events on the solver stream
        # (200 back-to-back launches); the data-term kernel that completes the operator apply is listed beside it
        achieved = words * B * N / sten_s / 1e9
        roof = {"bound": "hbm", "kernel": "stencil3d_tma_kernel<fused direction + stencil + p.q>" if t["fused"] else "stencil kernel",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload, args.precision),
                "algorithmic_bytes_per_cell": words * B, "algorithmic_bytes_per_launch": words * B * N, "peak_source": peak_src,
                "avg_launch_ms": sten_s * 1e3}
        extra = {"update_kernel": {"achieved": 7 * B * N / upd_s / 1e9, "frac": 7 * B * N / upd_s / 1e9 / peak, "avg_launch_ms": upd_s * 1e3,
                                   "algorithmic_bytes_per_cell": 7 * B},
                 "data_term_kernels": {"avg_launch_ms": data_s * 1e3, "note": "apply_blocks_kernel over the occupied cells (+ generic rows)"},
                 "apply_stencil_plus_data_term": {"achieved": words * B * N / apply_s / 1e9, "frac": words * B * N / apply_s / 1e9 / peak,
                                                  "avg_ms": apply_s * 1e3},
                 "iteration": {"achieved_52B_convention": BYTES_PER_CELL_ITER[args.precision] * N / it_s / 1e9,
                               "frac_52B_convention": BYTES_PER_CELL_ITER[args.precision] * N / it_s / 1e9 / peak,
                               "achieved_actual_bytes": (words + 7 + (0 if t["fused"] else 4)) * B * N / it_s / 1e9,
                               "frac_actual_bytes": (words + 7 + (0 if t["fused"] else 4)) * B * N / it_s / 1e9 / peak,
                               "ms_per_iteration": it_s * 1e3, "cell_iters_per_s_iterations_only": N / it_s}}

    if rank == 0 and runner is not None:
        # z-slab sharding: what one iteration moves over NVLink (R boundary planes of r to each neighbour, pushed by the
        # update kernel with peer stores, plus two 8/16-byte mailbox all-reduces), from rank 0's last step
        R = 2  # default Weights: model_2 -> radius-2 star
        it_ms = last_stats.get("solve_ms", 0.0) / max(1, last_stats.get("iterations", 1))
        if it_ms > 0:
            # no single kernel is timed alone on the slabs: the whole iteration in SURVEY §8(d)'s 52 B/cell convention, per GPU
            per_gpu = BYTES_PER_CELL_ITER[args.precision] * N / world / (it_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "whole Jacobi-PCG iteration on one slab (stencil + data term + update, exchanges included)",
                    "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak, "traffic": None,
                    "algorithmic_bytes_per_cell": BYTES_PER_CELL_ITER[args.precision], "peak_source": peak_src, "per": "GPU"}
        halo_bytes = 2 * R * n * n * (4 if args.precision == "f32" else 8)  # an interior rank: both neighbours
        extra = {"slab": {"planes_per_rank": [b - a for a, b in zip(runner.cuts, runner.cuts[1:])] if runner.cuts else n // world,
                          "partition": "cost-balanced (fi_slab_balanced_cuts)" if runner.cuts else "uniform", "setup_ms": last_stats.get("setup_ms"), "solve_ms": last_stats.get("solve_ms"),
                          "ms_per_iteration": it_ms, "halo_bytes_per_iteration_per_interior_rank": halo_bytes,
                          "halo_GBps_per_interior_rank_averaged_over_iteration": halo_bytes / max(it_ms, 1e-9) / 1e6,
                          "path": "peer stores over NVLink inside pcg_update_peer_kernel" if os.environ.get("FI_B200_P2P", "1") != "0" else "ncclSend/ncclRecv"}}

    def emit(ttt, base=None):
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": desc, "lattice": sizes, "points": npts, "weights": "default (model_2=0.5, trilinear value rows, cell-edge gradient rows)",
                       "pcg_iterations_per_step": args.iters, "step": "assemble normal equations from device-resident points + PCG iterations",
                       "l2": "working set (>= 6 lattice vectors) exceeds L2; no flush needed" if N * 4 * 6 > 126e6 else "flush not applied",
                       "parallelism": "single GPU" if world == 1 else f"z-slab x{world}"},
            "roofline": roof, "cpu_baseline": base,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clock_summary, "kernels": extra, "time_to_1e-6": ttt,
        }
        print(json.dumps(line), flush=True)

    ttt = None
    watchdog = None
    if runner is not None and not args.no_time_to_tol:
        # The sharded V-cycle is the one part of this file that exchanges halos through NCCL point-to-point calls between
        # the timed steps' barriers; a rank that fails alone would leave its peers waiting for ever.  The headline numbers
        # above are complete by now, so every rank arms the same deadline: when it passes, rank 0 prints the line without
        # the multigrid figures and all ranks leave.
        def bail():
            if rank == 0:
                emit({"error": f"sharded multigrid time-to-1e-6 did not finish within {args.mg_timeout} s; skipped"})
            sys.stdout.flush()
            os._exit(0)
        watchdog = threading.Timer(args.mg_timeout, bail)
        watchdog.daemon = True
        watchdog.start()
    if runner is not None and not args.no_time_to_tol:
        # metric (ii) on the slabs: every rank passes the whole cloud from HOST arrays and receives its owned planes in host
        # memory; the V-cycle is sharded (fi_slab_mg_plan).  Wall time between barriers, max over ranks.
        try:
            from field_interpolation_b200 import dist as fid
            ttt = {}
            h_own = h_out.numpy()
            for pname, pcode in (("f32", fi.FI_F32), ("f64_outer_f32_vcycle", fi.FI_F64)):
                best = None
                for rep in range(3):  # the first pass also pays allocations and NCCL channel setup
                    barrier()
                    t0 = time.perf_counter()
                    st = runner.step(h_pos.numpy(), h_nrm.numpy(), fi.solve_options(pcode, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID), h_own)
                    barrier()
                    t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    sec = float(t.item())
                    if rep > 0 and (best is None or sec < best["seconds"]):
                        best = {"seconds": sec, "iterations": int(st["iterations"]), "true_residual": st["true_residual"],
                                "recurrence_residual": st["relative_residual"], "converged": bool(st["converged"]),
                                "solve_ms": st["solve_ms"], "setup_ms": st["setup_ms"]}
                ttt[pname] = best
            plan = fid.slab_mg_plan(sizes, world, 2)  # level sizes (the plane ranges it lists are those of the uniform partition)
            ttt["method"] = (f"fi_slab_sdf_solve (host arrays) + multigrid-preconditioned CG, V-cycle z-slab sharded on {plan['sharded_levels']} level(s) "
                             f"({'/'.join('x'.join(map(str, s_)) for s_ in plan['sizes'][:-1])}), replicated from {'x'.join(map(str, plan['sizes'][-1]))}; best of 2 after one warm pass")
        except Exception as e:  # deterministic refusals (lattice not shardable this way) hit every rank alike
            ttt = {"error": repr(e)}
    if rank == 0 and runner is None and not args.no_time_to_tol:
        # metric (ii): host point arrays -> field with |AtA x - Atb| / |Atb| <= 1e-6, everything included
        # (H2D, assembly, multigrid hierarchy, MG-preconditioned CG, D2H).  Not part of `value`.
        ttt = {}
        for pname, pcode in (("f32", fi.FI_F32), ("f64_outer_f32_vcycle", fi.FI_F64)):
            best = None
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                f = fi.sdf_from_points(sizes, weights, h_pos.numpy(), h_nrm.numpy())
                _, st = f.solve(fi.solve_options(pcode, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID), out=h_out.numpy())
                f.close()
                sec = time.perf_counter() - t0
                if best is None or sec < best["seconds"]:
                    best = {"seconds": sec, "iterations": int(st["iterations"]), "true_residual": st["true_residual"],
                            "recurrence_residual": st["relative_residual"], "converged": bool(st["converged"]),
                            "solve_ms": st["solve_ms"], "setup_ms": st["setup_ms"]}
            ttt[pname] = best
        ttt["method"] = "sdf_from_points (host arrays) + multigrid-preconditioned CG (FI_PRECOND_MULTIGRID), best of 2"
        if args.time_to_tol:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x, cst = fi.sdf_solve_cascade(sizes, weights, torch.from_numpy(cloud["unit_pos"]).cuda(), d_nrm,
                                          options=fi.solve_options(fi.FI_MIXED, 400000, 1e-6, check_every=100), factor=2, coarsest_size=16)
            torch.cuda.synchronize()
            ttt["jacobi_cascade"] = {"seconds": time.perf_counter() - t0, "levels": cst["levels"], "level_iterations": cst["level_iterations"],
                                     "true_residual": cst["finest"]["true_residual"], "converged": bool(cst["finest"]["converged"]),
                                     "precision": "f32 PCG + f64 refinement"}

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sn, snpts, _ = WORKLOADS[args.cpu_sample]
        v, sec, base = cpu_reference_arm(sn, snpts, args.cpu_iters, 1, 0)
        base["value"], base["unit"] = v, unit

    if watchdog is not None:
        watchdog.cancel()
    if rank == 0:
        emit(ttt, base)
    if runner is not None and isinstance(ttt, dict) and "error" in ttt:
        sys.stdout.flush()
        os._exit(0)  # a rank that failed alone must not wait for its peers in destroy_process_group
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
