"""CPU: kernel LOGIC of the CUDA sources on the functional emulator of tests/emu/ (the library's own .cu files compiled
with g++; every thread of a block is a fiber, so __syncthreads, warp collectives, the TMA/mbarrier pipeline of
stencil_tma.cu and the reduction protocols run with their real semantics, block after block on the host).

These tests exist because the build container has no GPU: they catch indexing / protocol errors before a B200 box is
available.  They are NOT the parity evidence for the CUDA build (floating-point contraction, memory model and the real
TMA unit are out of their reach) — the `-m gpu` tests are — and nothing in the product ever loads the emulated library.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT, assert_system_bit_exact, bits, golden_names, load_golden
from field_interpolation_b200 import workloads as W
from oracle import oracle as O


@pytest.mark.parametrize("name", golden_names("iso_"))
def test_isosurface_kernels_on_reference_fixtures(emu, name):
    """csrc/isosurface.cu: count -> scan -> emit marching squares, area, Catmull-Rom upsampling against outputs frozen
    from the reference's own code (bit for bit; compiled like the CUDA build without fp contraction)."""
    g = load_golden(name)
    field, up = g["field"], int(g["upsample"])
    lines, area = emu.iso_surface(field, 0.0, want_area=True)
    assert lines.shape == g["lines"].shape and np.array_equal(bits(lines), bits(g["lines"]))
    assert abs(np.float32(area) - g["area"]) <= np.spacing(np.abs(g["area"]))
    big = emu.bicubic_upsample(field, up)
    assert np.array_equal(bits(big), bits(g["upsampled"]))
    il = emu.iso_surface(big, float(g["iso"]))
    assert il.shape == g["iso_lines_up"].shape and np.array_equal(bits(il), bits(g["iso_lines_up"]))


@pytest.mark.parametrize("shape", [(31, 1025), (300, 257)])
def test_isosurface_blocks_spanning_rows(emu, port, shape):
    rng = np.random.default_rng(shape[0])
    a = rng.standard_normal(shape).astype(np.float32)
    a[rng.random(shape) < 0.1] = 0.0
    want = port.iso_surface(a, 0.25)
    got, area = emu.iso_surface(a, 0.25, want_area=True)
    assert got.shape == want.shape and np.array_equal(bits(got), bits(want))
    assert abs(np.float32(area) - np.float32(port.calc_area(want))) <= np.spacing(np.abs(np.float32(port.calc_area(want))))


def test_assembly_and_solves_small_3d(emu, port):
    """Scatter / sort / triplet export, the TMA-staged fused PCG iteration and the multigrid V-cycle on a lattice the
    emulator finishes in seconds: triplet view bit for bit against the port, solved fields against the exact solve."""
    sizes = [32, 8, 12]
    cloud = W.sphere_torus_3d(400, seed=3)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    f = emu.sdf_from_points(sizes, emu.Weights(), pos, cloud["normals"])
    want = port.sdf_from_points(sizes, O.make_weights(), pos, cloud["normals"]).system()
    assert_system_bit_exact(f.eq, want.rows, want.cols, want.vals, want.rhs)
    exact = O.exact_solve(want, f.num_unknowns)
    x, st = f.solve(emu.solve_options(emu.FI_F64, 0, 1e-10, check_every=64))
    assert st["converged"] and np.linalg.norm(x - exact) <= 1e-5 * np.linalg.norm(exact)
    # the three stencil kernels agree on the operator
    v = np.random.default_rng(0).normal(size=f.num_unknowns)
    ys = []
    for mode in (1, 2, 0):
        f.use_fast_stencil(mode)
        ys.append(f.apply(v, emu.FI_F64))
    assert np.allclose(ys[0], ys[2], rtol=0, atol=1e-11 * np.abs(ys[2]).max()) and np.allclose(ys[1], ys[2], rtol=0, atol=1e-11 * np.abs(ys[2]).max())
    f.use_fast_stencil(1)
    xm, stm = f.solve(emu.solve_options(emu.FI_F64, 200, 1e-10, preconditioner=emu.FI_PRECOND_MULTIGRID))
    assert stm["converged"] and stm["iterations"] < st["iterations"] and np.linalg.norm(xm - exact) <= 1e-5 * np.linalg.norm(exact)


# ---- world_size > 1: the z-slab sharded solves over the fake NCCL (processes + shared memory) ---------------------------
def _run_ranks(world, case, gather=None, p2p=False, extra_env=None):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    build_emu.build()
    nccl_dir = build_emu.build_fake_nccl()
    with tempfile.TemporaryDirectory() as work:
        env = dict(os.environ, LD_LIBRARY_PATH=nccl_dir + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        env.pop("FI_B200_MG_GATHER_CELLS", None)
        if p2p:  # emulated CUDA IPC (shared-memory allocations): mailboxes + halo push by peer stores, graph-captured iterations
            env["CUDA_EMU_IPC"] = "1"
            env.pop("FI_B200_P2P", None)
        else:    # ncclSend / ncclRecv + ncclAllReduce inside the iteration
            env["FI_B200_P2P"] = "0"
        env.pop("FI_B200_PEER_FOLD", None)
        env.update(extra_env or {})
        if gather:
            env["FI_B200_MG_GATHER_CELLS"] = str(gather)
        # stderr goes to files: a rank blocked on a full pipe would stall its peers in the next collective
        logs = [open(os.path.join(work, f"stderr_rank{r}.log"), "w") for r in range(world)]
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "emu", "slab_rank.py"), str(r), str(world), work, json.dumps(case)],
                                  env=env, stderr=logs[r]) for r in range(world)]
        try:
            for p in procs:
                p.wait(timeout=600)
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
            for f in logs:
                f.close()
        errs = [open(os.path.join(work, f"stderr_rank{r}.log")).read() for r in range(world)]
        assert [p.returncode for p in procs] == [0] * world, "\n".join(e[-800:] for e in errs)
        stats = json.load(open(os.path.join(work, "stats_rank0.json")))
        for r in range(1, world):  # every rank computed the same partition
            assert json.load(open(os.path.join(work, f"stats_rank{r}.json")))["cuts"] == stats["cuts"]
        fields = {k: np.concatenate([np.load(os.path.join(work, f"{k}_rank{r}.npy")) for r in range(world)]) for k in case["solves"]}
    return fields, stats


SLAB_SOLVES = {"pcg64": {"precision": "f64", "max_iterations": 10, "tolerance": 1e-30},
               "pcg32": {"precision": "f32", "max_iterations": 10, "tolerance": 1e-30},
               "mg64": {"precision": "f64", "max_iterations": 100, "tolerance": 1e-9, "multigrid": True}}


@pytest.mark.parametrize("sizes,weights,gather", [([32, 24, 24], {}, 1000), ([32, 16, 37], {"model_1": 0.3}, None)])
def test_slab_sharded_solves_match_one_rank(sizes, weights, gather):
    """fi_slab_sdf_solve with 2 and 3 emulated ranks against the same call on 1 rank: Jacobi-PCG iterate for iterate
    (the halo planes of r travel by ncclSend/ncclRecv, the fused kernel recomputes the direction on the halo planes), and
    the sharded multigrid V-cycle (halo exchange before every smoother application, restriction ownership, all-gather of
    the restricted residual, replicated tail) — the same linear operator whatever the number of ranks: same iteration
    count, same field.  gather = 1000 forces two sharded levels; nz = 37 gives ragged slabs."""
    case = {"sizes": sizes, "points": 2500, "seed": 5, "weights": weights, "solves": SLAB_SOLVES}
    base, st1 = _run_ranks(1, case, gather)
    assert st1["mg64"]["converged"]
    for world in (2, 3):
        out, st = _run_ranks(world, case, gather)
        assert st["pcg64"]["iterations"] == st["pcg32"]["iterations"] == 10
        assert np.linalg.norm(out["pcg64"] - base["pcg64"]) <= 1e-7 * np.linalg.norm(base["pcg64"])  # fp64 iterates, stored as float
        assert np.linalg.norm(out["pcg32"] - base["pcg32"]) <= 2e-4 * np.linalg.norm(base["pcg32"])
        assert st["mg64"]["converged"] and abs(st["mg64"]["iterations"] - st1["mg64"]["iterations"]) <= 1
        assert abs(st["mg64"]["relative_residual"] - st1["mg64"]["relative_residual"]) <= 0.05 * st1["mg64"]["relative_residual"]
        assert np.linalg.norm(out["mg64"] - base["mg64"]) <= 1e-6 * np.linalg.norm(base["mg64"])


def test_slab_peer_memory_path_matches_one_rank():
    """The default multi-GPU iteration — r in a CUDA-IPC-mapped vector, boundary planes stored straight into the
    neighbours' halos by pcg_update_peer_kernel, scalar sums through the per-rank mailboxes, no NCCL inside the CUDA
    graph — with 2 and 3 emulated ranks (processes sharing memory) against 1 rank: iterate for iterate in fp64, and the
    same iteration count to a tolerance (the device-side stop and the leftover replays of a finished solve included)."""
    solves = {"pcg64": {"precision": "f64", "max_iterations": 12, "tolerance": 1e-30},
              "pcg32": {"precision": "f32", "max_iterations": 12, "tolerance": 1e-30},
              "conv64": {"precision": "f64", "max_iterations": 2000, "tolerance": 1e-3},
              "guess64": {"precision": "f64", "max_iterations": 12, "tolerance": 1e-30, "guess": True}}  # each rank passes its planes of a guess
    case = {"sizes": [32, 16, 24], "points": 2500, "seed": 5, "weights": {}, "point_weights": True, "solves": solves}
    base, st1 = _run_ranks(1, case)
    assert st1["conv64"]["converged"]
    for world in (2, 3):
        out, st = _run_ranks(world, case, p2p=True)
        assert st["pcg64"]["iterations"] == st["pcg32"]["iterations"] == 12
        assert np.linalg.norm(out["pcg64"] - base["pcg64"]) <= 1e-7 * np.linalg.norm(base["pcg64"])
        assert np.linalg.norm(out["guess64"] - base["guess64"]) <= 1e-7 * np.linalg.norm(base["guess64"])
        assert np.linalg.norm(out["guess64"] - base["pcg64"]) > 1e-3 * np.linalg.norm(base["pcg64"])  # the guess did enter
        assert np.linalg.norm(out["pcg32"] - base["pcg32"]) <= 2e-4 * np.linalg.norm(base["pcg32"])
        assert st["conv64"]["converged"] and abs(st["conv64"]["iterations"] - st1["conv64"]["iterations"]) <= 2
        assert np.linalg.norm(out["conv64"] - base["conv64"]) <= 1e-4 * np.linalg.norm(base["conv64"])


@pytest.mark.parametrize("target,select", [
    ("tests/test_gpu_assembly.py", "not device_resident"),
    ("tests/test_gpu_solve.py", "operator_matches or 1d_known or generic_rows or jacobi or randomised_golden"),
])
def test_gpu_test_bodies_pass_on_the_emulator(target, select):
    """The bodies of the `-m gpu` parity tests (triplet view bit for bit against the reference's fixtures and the port;
    the operator against the explicit normal equations; 1D/2D/3D golden solutions; generic rows; Jacobi sweeps), run in
    a child pytest whose library handle is the emulator build (FI_B200_TEST_EMU=1, tests/conftest.py).  Selections are
    the cases the emulator finishes in seconds and that need no torch CUDA tensor."""
    env = dict(os.environ, FI_B200_TEST_EMU="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, target), "-m", "gpu", "-q", "-x", "-k", select, "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout and " failed" not in r.stdout


def test_cpp_drop_in_callers_pass_on_the_emulator(tmp_path):
    """tests/cpp/api_driver.cpp (the reference's own callers replayed against the drop-in headers, iso-surface sequence
    included) with libfi_b200.so resolved to the emulator build: exercises the C++ host layer end to end on the CPU."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    os.symlink(build_emu.build(), tmp_path / "libfi_b200.so")
    env = dict(os.environ, LD_LIBRARY_PATH=str(tmp_path) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_cpp_api.py"),
                        os.path.join(ROOT, "tests", "test_gpu_zz_isosurface.py"), "-m", "gpu", "-q", "-x", "-k",
                        "readme or field_1d or interpolate_2d or hand_written or cpp_drop_in", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout and " failed" not in r.stdout


def test_slab_custom_and_balanced_cuts_match_one_rank():
    """fi_comm_set_slab_cuts / fi_slab_balanced_cuts: a deliberately lopsided partition and the cost-balanced one (point
    weight exaggerated so that it differs visibly from the uniform cuts) give the same iterates and fields as one rank,
    over NCCL, over peer memory, and through the sharded multigrid."""
    solves = dict(SLAB_SOLVES)
    case = {"sizes": [32, 16, 40], "points": 3000, "seed": 5, "weights": {}, "solves": solves}
    base, st1 = _run_ranks(1, case)
    for cuts, p2p in (([0, 6, 29, 40], False), ("balanced", True)):
        c = dict(case, cuts=cuts, point_weight=60.0, min_planes=4)
        out, st = _run_ranks(3, c, p2p=p2p)
        if cuts == "balanced":
            got = st["cuts"]
            assert got[0] == 0 and got[-1] == 40 and all(b - a >= 4 for a, b in zip(got, got[1:])) and got != [0, 13, 26, 40]
            assert got[2] - got[1] < 13  # the middle slab holds most of the cloud: it gets fewer planes
        else:
            assert st["cuts"] == cuts
        assert np.linalg.norm(out["pcg64"] - base["pcg64"]) <= 1e-7 * np.linalg.norm(base["pcg64"])
        assert np.linalg.norm(out["pcg32"] - base["pcg32"]) <= 2e-4 * np.linalg.norm(base["pcg32"])
        assert st["mg64"]["converged"] and abs(st["mg64"]["iterations"] - st1["mg64"]["iterations"]) <= 1
        assert np.linalg.norm(out["mg64"] - base["mg64"]) <= 1e-6 * np.linalg.norm(base["mg64"])


def test_slab_peer_fold_mode_matches_one_rank():
    """FI_B200_PEER_FOLD=1 (opt-in until measured on GPUs): p.Ap is published by block 0 of the update kernel and the
    iteration is finished by its last block — two kernels fewer per iteration.  Same iterates, same stop."""
    solves = {"pcg64": {"precision": "f64", "max_iterations": 12, "tolerance": 1e-30},
              "conv64": {"precision": "f64", "max_iterations": 2000, "tolerance": 1e-3}}
    case = {"sizes": [32, 16, 24], "points": 2500, "seed": 5, "weights": {}, "solves": solves}
    base, st1 = _run_ranks(1, case)
    for world in (3,):
        out, st = _run_ranks(world, case, p2p=True, extra_env={"FI_B200_PEER_FOLD": "1"})
        assert st["pcg64"]["iterations"] == 12 and np.linalg.norm(out["pcg64"] - base["pcg64"]) <= 1e-7 * np.linalg.norm(base["pcg64"])
        assert st["conv64"]["converged"] and abs(st["conv64"]["iterations"] - st1["conv64"]["iterations"]) <= 2
        assert np.linalg.norm(out["conv64"] - base["conv64"]) <= 1e-4 * np.linalg.norm(base["conv64"])


def _two_phase_cuts(hist, plane, world, w, min_planes, ratio=1.22):
    """numpy restatement of dist.cu: balanced_cuts — minimise max A + max B over contiguous partitions."""
    nz = len(hist)
    cum = np.concatenate([[0.0], np.cumsum(plane + w * hist.astype(np.float64))])

    def fill(target, cap):
        cuts = [0]
        for k in range(world):
            z0 = cuts[-1]
            room = nz - z0 - (world - 1 - k) * min_planes
            z1 = min(z0 + min(cap, room), nz)
            if z1 - z0 < min_planes:
                return None
            hi = int(np.searchsorted(cum[z0 + 1:z1 + 1], cum[z0] + target, side="right")) + z0
            z1 = min(z1, hi)
            if z1 - z0 < min_planes:
                return None
            cuts.append(z1)
        return cuts if cuts[-1] == nz else None

    best, best_cost = None, -1.0
    for cap in range((nz + world - 1) // world, nz - (world - 1) * min_planes + 1):
        lo, hi = cum[-1] / world, cum[-1]
        if fill(hi, cap) is None:
            continue
        it = 0
        while it < 60 and hi - lo > 0.25 * plane * 1e-3:
            mid = 0.5 * (lo + hi)
            if fill(mid, cap) is not None:
                hi = mid
            else:
                lo = mid
            it += 1
        cuts = fill(hi, cap)
        if cuts is None:
            continue
        a = max(cum[cuts[k + 1]] - cum[cuts[k]] for k in range(world))
        p = max(cuts[k + 1] - cuts[k] for k in range(world))
        cost = a + ratio * plane * p
        if best_cost < 0 or cost < best_cost * (1.0 - 1e-12):
            best, best_cost = cuts, cost
    return best, best_cost


def test_balanced_cuts_of_the_bench_cloud(emu):
    """fi_slab_balanced_cuts on bench.py's workload (512^3 lattice, 1M sphere+torus points): the partition the 8-, 4- and
    2-GPU bench lines run on, against the same two-phase cost model in numpy; the slabs that hold the torus get fewer planes,
    and the model's iteration cost beats both the uniform partition and round 1's single-sum balance."""
    from field_interpolation_b200 import dist as fid
    n = 512
    cloud = W.sphere_torus_3d(1_000_000, seed=0)
    pos = W.to_lattice(cloud["unit_pos"], [n, n, n])
    z = np.floor(pos[:, 2])
    hist = np.bincount(z[(z >= 0) & (z < n)].astype(np.int64), minlength=n)
    plane = float(n) * n
    for world in (2, 4, 8):
        want, cost = _two_phase_cuts(hist, plane, world, 30.0, 8)
        got = fid.balanced_cuts([n, n, n], world, pos, 0.0, 8)
        assert got == want, (world, got, want)
    planes = np.diff(got)
    assert planes[3] < 56 and planes[4] < 64 and planes.max() <= 72 and planes.sum() == n, planes

    def model(cuts):
        cum = np.concatenate([[0.0], np.cumsum(plane + 30.0 * hist)])
        return max(cum[b] - cum[a] for a, b in zip(cuts, cuts[1:])) + 1.22 * plane * max(b - a for a, b in zip(cuts, cuts[1:]))
    uniform = [64 * k for k in range(9)]
    round1 = np.concatenate([[0], np.cumsum([68, 67, 64, 57, 57, 64, 67, 68])]).tolist()
    assert model(got) < 0.97 * model(round1) < model(uniform)
