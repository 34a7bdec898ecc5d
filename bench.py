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


def kernel_source_sha():
    """Identifies the build of the dominant kernel: sha256 over the sources that define it."""
    import hashlib
    h = hashlib.sha256()
    for name in ("stencil_tma.cu", "tma.cuh", "peer.cuh", "common.cuh"):  # everything the kernel's code comes from
        with open(os.path.join(ROOT, "field_interpolation_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_traffic(workload, precision):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu capture of
    this workload (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py on a GPU visit).  A capture taken from
    another build of the kernel (source hash differs) is refused: (None, why)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, "no capture committed"
    t = json.load(open(path)).get(f"{workload}:{precision}")
    if t is None:
        return None, "no capture of this workload"
    if t.get("kernel_source_sha") != kernel_source_sha():
        return None, f"stale capture refused: taken from kernel sources {t.get('kernel_source_sha')}, this build is {kernel_source_sha()}"
    return t["dram_bytes_per_launch"], f"ncu capture {t.get('captured', '?')} of this kernel build"


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
def cpu_reference_arm(n: int, npts: int, iters: int, budget_s: float, max_steps: int):
    """The reference's CPU path on this box's host cores: assembly by the reference's own code when oracle/_ref is
    present (else the port), then the restated Eigen path (CSC -> AtA -> Jacobi-preconditioned BiCGSTAB, float) for a
    bounded number of iterations.  Single thread: the reference has no threading.  Runs whole steps (assembly + AtA +
    `iters` iterations) until `budget_s` is used up, at least one, at most max_steps; returns what actually ran."""
    from field_interpolation_b200 import workloads as W
    from oracle import oracle as O
    ref = O.reference()
    asm, kind = (ref, "reference") if ref is not None else (O.port(), "port")
    cloud = W.sphere_torus_3d(npts, seed=0)
    sizes = [n, n, n]
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    times, asm_s, ata_s, it_s = [], 0.0, 0.0, 0.0
    began = time.perf_counter()
    while len(times) < max(1, max_steps):
        t0 = time.perf_counter()
        sys_ = asm.sdf_from_points(sizes, O.make_weights(), pos, cloud["normals"]).system()
        t1 = time.perf_counter()
        N = O.port().normal(sys_, n ** 3, "f32")
        t2 = time.perf_counter()
        x, its, err = N.bicgstab(guess=np.zeros(n ** 3, np.float32), max_iter=iters, tol=1e-30)
        t3 = time.perf_counter()
        times.append((t3 - t0, its))
        asm_s, ata_s, it_s = t1 - t0, t2 - t1, (t3 - t2) / max(1, its)
        del N, sys_
        if time.perf_counter() - began + (t3 - t0) > budget_s:
            break
    total = sum(t for t, _ in times)
    its = sum(i for _, i in times)
    value = (n ** 3) * its / total
    return value, total / len(times), len(times), {
        "kind": "port" if kind == "port" else "reference(assembly)+port(Eigen path restated; Eigen not installable offline)",
        "cores": 1, "host_cores": os.cpu_count(),
        "sample": f"{n}^3 lattice, {npts} points, {len(times)} step(s) of: assembly ({asm_s:.2f} s) + CSC and AtA ({ata_s:.2f} s) + "
                  f"{iters} BiCGSTAB iterations ({it_s:.3f} s each, 2 SpMV per iteration); iterations only: {(n ** 3) / max(it_s, 1e-12):.3e} cell-iters/s",
        "assembly_s": asm_s, "csc_and_ata_s": ata_s, "s_per_bicgstab_iteration": it_s, "steps_run": len(times),
    }


def cpu_time_to_tol(n: int, npts: int, tol: float = 1e-6):
    """Host time of the reference's solve to a relative residual: assembly + AtA + BiCGSTAB (diagonal preconditioner) from a zero
    guess, in DOUBLE — the float path (what solve_sparse_linear_with_guess runs) does not get there: at 64^3 the restated
    float BiCGSTAB ends in NaN after 4,765 iterations (measured)."""
    from field_interpolation_b200 import workloads as W
    from oracle import oracle as O
    ref = O.reference()
    asm = ref if ref is not None else O.port()
    cloud = W.sphere_torus_3d(npts, seed=0)
    sizes = [n, n, n]
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    t0 = time.perf_counter()
    sys_ = asm.sdf_from_points(sizes, O.make_weights(), pos, cloud["normals"]).system()
    N = O.port().normal(sys_, n ** 3, "f64")
    x, its, err = N.bicgstab(guess=np.zeros(n ** 3, np.float64), max_iter=0, tol=tol)
    return {"lattice": sizes, "points": npts, "seconds": time.perf_counter() - t0, "bicgstab_iterations": its, "relative_residual": err,
            "arithmetic": "f64 (restated Eigen BiCGSTAB; the float path ends in NaN at this size)", "cores": 1}


def single_gpu_configs(fi, W, torch, peak):
    """BASELINE.json configs[0..3] on one GPU, each through the public API from HOST arrays: Jacobi-PCG throughput (the
    reference's preconditioner; lattice cells x iterations / s over a fixed 200 iterations, device time) and the time to a
    1e-6 true relative residual with the multigrid-preconditioned CG (host wall time: H2D, assembly, hierarchy, solve, D2H)."""
    out = {}

    def jacobi_and_mg(f_make, sizes, label, precision=fi.FI_F32):
        N = int(np.prod(sizes))
        f = f_make()
        opt = fi.solve_options(precision, 0, 1e-6)
        f.time_iterations(20, opt)
        t = f.time_iterations(200, opt)
        f.close()
        it_s = t["iteration_ms"] / 200e3
        B = 4 if precision == fi.FI_F32 else 8
        sec = {"lattice": sizes, "cells": N,
               "jacobi_pcg": {"ms_per_iteration": it_s * 1e3, "cell_iters_per_s": N / it_s, "fused_direction_stencil_kernel": bool(t["fused"]),
                              "stencil_ms": t["stencil_ms"] / 200, "data_term_ms": t["data_term_ms"] / 200, "update_ms": t["update_ms"] / 200,
                              "GBps_52B_convention": 13 * B * N / it_s / 1e9, "frac_of_hbm_peak_52B_convention": 13 * B * N / it_s / 1e9 / peak,
                              "resident_in_L2": bool(N * B * 6 < 100e6)}}
        best = None
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f = f_make()
            _, st = f.solve(fi.solve_options(fi.FI_F64, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID))
            f.close()
            dt = time.perf_counter() - t0
            if best is None or dt < best["seconds"]:
                best = {"seconds": dt, "iterations": int(st["iterations"]), "true_residual": st["true_residual"], "converged": bool(st["converged"]),
                        "solve_ms": st["solve_ms"], "setup_ms": st["setup_ms"], "arithmetic": "fp64 outer CG, fp32 V-cycle"}
        sec["time_to_1e-6_multigrid"] = best
        out[label] = sec

    # C1: the 1D demo (src/field_1d.cpp:98-114): data rows first, then the model rows; solved "exactly" every frame
    c1 = W.field_1d(100)
    w = fi.Weights()

    def frame():
        f = fi.LatticeField(c1["sizes"])
        for p, v, g in zip(c1["pos"], c1["value"], c1["gradient"]):
            fi.add_value_constraint(f, p, float(v), w.data_pos)
            fi.add_gradient_constraint(f, p, g, w.data_gradient, w.gradient_kernel)
        fi.add_field_constraints(f, w)
        # 100 unknowns: the multigrid hierarchy is a single level whose "coarsest" dense inverse is computed on the device, i.e. a
        # direct solve refined by the fp64 outer CG (Jacobi-PCG needs > 2N iterations on this 1D fourth-order system)
        x, st = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-10, preconditioner=fi.FI_PRECOND_MULTIGRID))
        f.close()
        return st
    frame()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        st = frame()
    out["C1_field_1d_100"] = {"lattice": c1["sizes"], "ms_per_frame_assemble_and_solve": (time.perf_counter() - t0) / reps * 1e3,
                              "pcg_iterations": int(st["iterations"]), "true_residual": st["true_residual"], "converged": bool(st["converged"]),
                              "note": "100 unknowns: launch- and sync-latency bound; the reference's Cholesky solves this on the CPU in microseconds"}

    # C2: 512^2, 10k noisy value points, model_1 + model_2 (src/interpolate_2d.cpp:31-47)
    c2 = W.interpolate_2d(512, 10_000, seed=1)
    p2 = W.to_lattice(c2["unit_pos"], c2["sizes"])
    w2 = fi.Weights(**c2["weights"])

    def make_c2():
        f = fi.LatticeField(c2["sizes"])
        fi.add_field_constraints(f, w2)
        fi.add_points(f, w2.data_pos, w2.value_kernel, 0.0, w2.gradient_kernel, p2, None, None, c2["value"])
        return f
    jacobi_and_mg(make_c2, c2["sizes"], "C2_interpolate2d_512_10k")

    # C3: 2048^2 from 200k oriented points; C4: 256^3 from 1M
    c3 = W.circles_2d(200_000, seed=0)
    p3 = W.to_lattice(c3["unit_pos"], [2048, 2048])
    jacobi_and_mg(lambda: fi.sdf_from_points([2048, 2048], w, p3, c3["normals"]), [2048, 2048], "C3_sdf2d_2048_200k")
    c4 = W.sphere_torus_3d(1_000_000, seed=0)
    p4 = W.to_lattice(c4["unit_pos"], [256, 256, 256])
    jacobi_and_mg(lambda: fi.sdf_from_points([256, 256, 256], w, p4, c4["normals"]), [256, 256, 256], "C4_sdf3d_256_1M")
    # the same step the reference arm times on this configuration (bench.py --impl reference): assembly + 20 iterations
    f = fi.sdf_from_points([256, 256, 256], w, p4, c4["normals"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f2 = fi.sdf_from_points([256, 256, 256], w, p4, c4["normals"])
    _, st = f2.solve(fi.solve_options(fi.FI_F32, 20, 1e-30, check_every=20))
    dt = time.perf_counter() - t0
    f2.close()
    f.close()
    out["C4_sdf3d_256_1M"]["reference_arm_step"] = {"seconds": dt, "iterations": int(st["iterations"]), "cell_iters_per_s": 256 ** 3 * st["iterations"] / dt,
                                                     "step": "host arrays -> assemble -> 20 Jacobi-PCG iterations -> field on the host (what --impl reference times on the CPU, "
                                                             "with BiCGSTAB iterations of 2 SpMV each there)"}
    return out


def slab_parity_and_c5(fi, fid, W, torch, dist, rank, world, sizes, weights, d_pos, d_nrm, runner):
    """N > 1: (i) the sharded 512^3 solve against the single-GPU solve of the same system, every rank comparing the planes it
    owns (max over ranks of the relative L2 difference); (ii) at 8 ranks BASELINE configs[4]: 1024^3 from 20M points."""
    sec = {}
    N = int(np.prod(sizes))
    try:
        opt = fi.solve_options(fi.FI_F32, 100, 1e-30, check_every=50)
        own = torch.zeros(runner.local_cells, device="cuda")
        runner.step(d_pos, d_nrm, opt, own)
        f = fi.sdf_from_points(sizes, weights, d_pos, d_nrm)
        full = torch.zeros(N, device="cuda")
        _, st1 = f.solve(opt, out=full)
        f.close()
        plane = sizes[0] * sizes[1]
        mine = full[runner.z0 * plane:runner.z1 * plane]
        num = torch.linalg.vector_norm((own - mine).double()) ** 2
        den = torch.linalg.vector_norm(mine.double()) ** 2
        t = torch.stack([num, den])
        dist.all_reduce(t)
        sec["slab_vs_single_gpu"] = {"lattice": sizes, "iterations": 100, "precision": "f32", "relative_l2_difference": float(torch.sqrt(t[0] / t[1]).item()),
                                     "note": "100 Jacobi-PCG iterations from zero on every rank's owned planes vs the 1-GPU solve of the same system"}
        del full, own
        fi._lib.lib().fi_trim_memory()
    except Exception as e:
        sec["slab_vs_single_gpu"] = {"error": repr(e)}
    if world == 8:
        try:
            n5, pts5 = 1024, 20_000_000
            cloud = W.sphere_torus_3d(pts5, seed=0)
            s5 = [n5, n5, n5]
            h_pos = W.to_lattice(cloud["unit_pos"], s5)
            h_nrm = cloud["normals"]
            r5 = fid.SlabRunner(s5, weights, rank, world, dist)
            dp, dn = torch.from_numpy(h_pos).cuda(), torch.from_numpy(h_nrm).cuda()
            try:
                r5.set_cuts(fid.balanced_cuts(s5, world, dp, 0.0, 8))
            except fi.FiError:
                r5.set_cuts(None)
            out = torch.zeros(r5.local_cells, device="cuda")
            its = 100
            opt = fi.solve_options(fi.FI_F32, its, 1e-30, check_every=50)
            r5.step(dp, dn, opt, out)  # warm: allocations, NCCL / peer mappings
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            st = r5.step(dp, dn, opt, out)
            torch.cuda.synchronize(); dist.barrier()
            t = torch.tensor([time.perf_counter() - t0, st["solve_ms"] * 1e-3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c5 = {"lattice": s5, "points": pts5, "cells": n5 ** 3, "ranks": world,
                  "jacobi_pcg": {"iterations": its, "step_seconds_assemble_plus_iterations": float(t[0].item()), "ms_per_iteration": float(t[1].item()) * 1e3 / its,
                                 "cell_iters_per_s_iterations_only": n5 ** 3 * its / float(t[1].item()),
                                 "cell_iters_per_s_whole_step": n5 ** 3 * its / float(t[0].item())}}
            h_own = np.empty(r5.local_cells, np.float32)
            best = None
            for rep in range(2):
                torch.cuda.synchronize(); dist.barrier()
                t0 = time.perf_counter()
                st = r5.step(h_pos, h_nrm, fi.solve_options(fi.FI_F64, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID), h_own)
                torch.cuda.synchronize(); dist.barrier()
                tt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                if rep > 0 or best is None:
                    best = {"seconds": float(tt.item()), "iterations": int(st["iterations"]), "true_residual": st["true_residual"], "converged": bool(st["converged"]),
                            "solve_ms": st["solve_ms"], "setup_ms": st["setup_ms"], "arithmetic": "fp64 outer CG, fp32 sharded V-cycle", "from": "host arrays on every rank"}
            c5["time_to_1e-6_multigrid"] = best
            free, total = torch.cuda.mem_get_info()
            m = torch.tensor([float(total - free)], device="cuda", dtype=torch.float64)
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            c5["device_memory_GB_in_use_after_solve_max_over_ranks"] = float(m.item()) / 1e9
            c5["device_memory_note"] = "total - free right after the solves: includes the library's block cache, i.e. the high-water mark of its allocations, and torch's context"
            c5["reference"] = "not representable: 3.2e9 rows overflow the reference's int32 Triplet.row (sparse_linear.hpp:10)"
            r5.close()
            sec["C5_sdf3d_1024_20M"] = c5
        except Exception as e:
            sec["C5_sdf3d_1024_20M"] = {"error": repr(e)}
    return sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sdf3d_512_1M", choices=sorted(WORKLOADS))
    ap.add_argument("--iters", type=int, default=400, help="PCG iterations per step")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--cpu-sample", default=None, choices=sorted(WORKLOADS),
                    help="lattice the CPU arm runs on (default: sdf3d_256_1M = BASELINE configs[3] for --impl reference, sdf3d_128_1M for the cpu_baseline leg)")
    ap.add_argument("--cpu-iters", type=int, default=20, help="BiCGSTAB iterations per CPU step")
    ap.add_argument("--cpu-budget-s", type=float, default=100.0, help="--impl reference: host seconds to spend on whole CPU steps (at least one step runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-BASELINE-config sections (C1..C4 at N=1, C5 and the slab parity check at N=8)")
    ap.add_argument("--time-to-tol", action="store_true", help="also time the Jacobi-PCG coarse-to-fine cascade to 1e-6 (slow)")
    ap.add_argument("--no-time-to-tol", action="store_true", help="skip the multigrid time-to-1e-6 measurement")
    ap.add_argument("--mg-timeout", type=float, default=420.0, help="N>1: deadline in seconds for everything after the headline (sharded multigrid time-to-1e-6, slab parity, C5)")
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
        # The reference's representation (int32 rows, explicit triplets -> CSC -> AtA) cannot hold this arm's 512^3 lattice in
        # minutes or in memory it would be honest to ask for; the bounded sample is BASELINE configs[3] (256^3, the same cloud
        # generator) — the b200 arm reports the same configuration in its `configs.C4_sdf3d_256_1M` section.
        sn, snpts, sdesc = WORKLOADS[args.cpu_sample or "sdf3d_256_1M"]
        value, sec, ran, base = cpu_reference_arm(sn, snpts, args.cpu_iters, args.cpu_budget_s, max(1, args.steps))
        base["value"] = value
        base["unit"] = unit
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": ran,
            "warmup": 0, "requested_steps": args.steps, "requested_warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "cpu_sample": sdesc, "precision": "f32", "pcg_iterations_per_step": args.cpu_iters,
                       "step": "assemble (reference TU) + CSC + AtA + BiCGSTAB iterations, one host thread",
                       "note": "steps/warmup are what actually ran inside --cpu-budget-s; the sample lattice is smaller than the b200 arm's"},
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
        traffic, traffic_note = ncu_traffic(args.workload, args.precision)
        roof = {"bound": "hbm", "kernel": "stencil3d_tma_kernel<fused direction + stencil + p.q>" if t["fused"] else "stencil kernel",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                "algorithmic_bytes_per_cell": words * B, "algorithmic_bytes_per_launch": words * B * N, "peak_source": peak_src,
                "avg_launch_ms": sten_s * 1e3}
        # update kernel: the fused path updates x every second iteration with both terms (16 B/cell, then 32 B/cell in fp32): 6 words on average
        uw = 6 if t["fused"] else 7
        extra = {"update_kernel": {"achieved": uw * B * N / upd_s / 1e9, "frac": uw * B * N / upd_s / 1e9 / peak, "avg_launch_ms": upd_s * 1e3,
                                   "algorithmic_bytes_per_cell": uw * B,
                                   "note": "mean of the two alternating launches of the deferred x update" if t["fused"] else "x updated every iteration"},
                 "data_term_kernels": {"avg_launch_ms": data_s * 1e3, "note": "apply_blocks_kernel: one thread per occupied cell, 8 atomics into q (+ generic rows)"},
                 "apply_stencil_plus_data_term": {"achieved": words * B * N / apply_s / 1e9, "frac": words * B * N / apply_s / 1e9 / peak,
                                                  "avg_ms": apply_s * 1e3},
                 "iteration": {"achieved_52B_convention": BYTES_PER_CELL_ITER[args.precision] * N / it_s / 1e9,
                               "frac_52B_convention": BYTES_PER_CELL_ITER[args.precision] * N / it_s / 1e9 / peak,
                               "achieved_actual_bytes": (words + uw + (0 if t["fused"] else 4)) * B * N / it_s / 1e9,
                               "frac_actual_bytes": (words + uw + (0 if t["fused"] else 4)) * B * N / it_s / 1e9 / peak,
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
            "gpu_launches": int(launches), "clocks": clock_summary, "kernels": extra, "time_to_1e-6": ttt, "configs": configs,
        }
        print(json.dumps(line), flush=True)

    ttt = None
    watchdog = None
    configs = None
    if not args.no_configs and runner is None and rank == 0:
        configs = single_gpu_configs(fi, W, torch, peak)
    if runner is not None and not (args.no_time_to_tol and args.no_configs):
        # What follows (sharded V-cycle, slab-vs-single parity, the 1024^3 configuration) runs collectives between the timed
        # steps' barriers; a rank that fails alone would leave its peers waiting for ever.  The headline numbers above are
        # complete by now, so every rank arms the same deadline: when it passes, rank 0 prints the line with what has been
        # measured so far and all ranks leave.
        def bail():
            if rank == 0:
                emit(ttt if isinstance(ttt, dict) and ttt else {"error": f"the sections after the headline did not finish within {args.mg_timeout} s; skipped"})
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

    if runner is not None and not args.no_configs:
        from field_interpolation_b200 import dist as fid
        configs = slab_parity_and_c5(fi, fid, W, torch, dist, rank, world, sizes, weights, d_pos, d_nrm, runner)

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sn, snpts, _ = WORKLOADS[args.cpu_sample or "sdf3d_128_1M"]
        v, sec, _, base = cpu_reference_arm(sn, snpts, 2 * args.cpu_iters, 0.0, 1)
        base["value"], base["unit"] = v, unit
        # solve against solve, same configuration on both sides: the CPU path to 1e-6 at 64^3 next to the GPU's
        base["time_to_1e-6_64cubed"] = cpu_time_to_tol(64, 100_000)
        g64 = {}
        h64p = W.to_lattice(W.sphere_torus_3d(100_000, seed=0)["unit_pos"], [64, 64, 64])
        h64n = W.sphere_torus_3d(100_000, seed=0)["normals"]
        for pname, popt in (("jacobi_pcg_f64", fi.solve_options(fi.FI_F64, 0, 1e-6)),
                            ("multigrid_pcg_f32", fi.solve_options(fi.FI_F32, 500, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID))):
            best = None
            for rep in range(2):
                t0 = time.perf_counter()
                f64 = fi.sdf_from_points([64, 64, 64], weights, h64p, h64n)
                _, st = f64.solve(popt)
                f64.close()
                sec64 = time.perf_counter() - t0
                if best is None or sec64 < best["seconds"]:
                    best = {"seconds": sec64, "iterations": int(st["iterations"]), "true_residual": st["true_residual"], "converged": bool(st["converged"])}
            g64[pname] = best
        base["time_to_1e-6_64cubed"]["b200_same_config"] = g64

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
