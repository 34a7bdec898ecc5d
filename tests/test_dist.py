"""Multi-GPU path (csrc/dist.cu, field_interpolation_b200/dist.py).

CPU: the slab partition and the communicator bootstrap plumbing (world_size-2 gloo group).
GPU (one device): a 1-rank communicator still runs the whole sharded code path — window geometry with halo planes
beyond the lattice, owned-row filtering, NCCL all-reduces inside the CUDA graph — and must reproduce the plain
solve.  The 2..8 rank parity run is scripts/slab_check.py (needs gpurun --gpus N)."""
import os
import socket

import numpy as np
import pytest

from field_interpolation_b200 import workloads as W


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("nz", [5, 8, 37, 512, 1000, 1024])
def test_slab_ranges_partition_the_lattice(nz):
    from field_interpolation_b200 import dist as fid
    for world in range(1, 9):
        if world > nz:
            continue
        r = [fid.slab_range(nz, world, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == nz
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        sizes = [b - a for a, b in r]
        assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1


def _axis_tables(nf, nc):
    """Independent restatement of the transfer geometry along one axis (upscale_field's align-corners rule,
    reference field_interpolation.cpp:462): base coarse node of every fine node, and for every coarse node the
    fine nodes that interpolate from it with a non-zero weight."""
    sc = (nc - 1) / (nf - 1) if nf > 1 else 0.0
    base = np.minimum(np.floor(np.arange(nf) * sc).astype(int), nc - 1)
    frac = (np.arange(nf) * sc - base).astype(np.float32)
    touch = [[] for _ in range(nc)]
    for i in range(nf):
        if np.float32(1) - frac[i] != 0:
            touch[base[i]].append(i)
        if frac[i] != 0 and base[i] + 1 < nc:
            touch[base[i] + 1].append(i)
    return base, touch


@pytest.mark.parametrize("sizes,world,radius,gather", [
    ([512, 512, 512], 8, 2, 0), ([512, 512, 512], 2, 2, 0), ([1024, 1024, 1024], 8, 2, 0), ([256, 256, 256], 8, 2, 0),
    ([256, 256, 256], 4, 1, 1000), ([128, 64, 37], 2, 2, 1000), ([96, 40, 64], 3, 4, 1000), ([64, 48, 40], 1, 2, 1000),
    ([512, 256, 1000], 7, 3, 0), ([320, 200, 129], 5, 2, 100)])
def test_slab_multigrid_plan_windows(sizes, world, radius, gather):
    """fi_slab_mg_plan: every level's planes are partitioned over the ranks, restriction into a rank's coarse planes
    and prolongation into its fine planes read only planes inside its stored window (owned + halo)."""
    from field_interpolation_b200 import dist as fid
    plan = fid.slab_mg_plan(sizes, world, radius, gather)
    nd, halo = plan["sharded_levels"], plan["halo"]
    assert nd >= 1 and halo == max(radius, 2)
    assert plan["sizes"][0] == sizes
    assert plan["own"][0] == [fid.slab_range(sizes[2], world, k) for k in range(world)]
    for l in range(nd + 1):
        own = plan["own"][l]
        assert own[0][0] == 0 and own[-1][1] == plan["sizes"][l][2]
        assert all(own[k][1] == own[k + 1][0] for k in range(world - 1))
        if l > 0:
            assert plan["sizes"][l] == [(v + 1) // 2 for v in plan["sizes"][l - 1]]
        if l < nd:  # sharded: thick enough for the halo exchange, sizes the TMA stencil kernel takes
            assert all(b - a >= halo for a, b in own)
            assert plan["sizes"][l][0] % 4 == 0 and plan["sizes"][l][0] >= 32 and plan["sizes"][l][1] >= 8
    for l in range(nd):
        base, touch = _axis_tables(plan["sizes"][l][2], plan["sizes"][l + 1][2])
        for k in range(world):
            z0, z1 = plan["own"][l][k]
            c0, c1 = plan["own"][l + 1][k]
            for C in range(c0, c1):  # restriction reads the fine planes that touch C
                assert touch[C] and min(touch[C]) >= z0 - halo and max(touch[C]) < z1 + halo
            if l + 1 < nd:  # prolongation from a sharded level reads its planes base, base + 1
                for z in range(z0, z1):
                    assert base[z] >= c0 - halo and min(base[z] + 1, plan["sizes"][l + 1][2] - 1) < c1 + halo


def test_slab_multigrid_plan_refuses_thin_slabs():
    from field_interpolation_b200 import dist as fid
    from field_interpolation_b200 import _lib as L
    with pytest.raises(L.FiError):
        fid.slab_mg_plan([64, 64, 8], 8, 2)       # one plane per rank
    with pytest.raises(L.FiError):
        fid.slab_mg_plan([30, 64, 64], 2, 2)      # x size the TMA stencil kernel does not take


def _bootstrap_worker(rank, world, port, out):
    import torch.distributed as dist
    from field_interpolation_b200 import dist as fid
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    uid = fid.broadcast_unique_id(dist, rank)
    z = fid.slab_range(64, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, (uid.hex(), z))
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


def test_unique_id_broadcast_and_ranges_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bootstrap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (id0, z0), (id1, z1) = got
    assert id0 == id1 and len(id0) == 256 and int(id0, 16) != 0  # every rank holds rank 0's 128-byte NCCL id
    assert z0 == (0, 32) and z1 == (32, 64)


@pytest.mark.gpu
@pytest.mark.parametrize("orders", [dict(), dict(model_1=0.4), dict(model_2=0.0, model_3=0.3), dict(model_0=0.1, model_4=0.2)])
def test_one_rank_slab_matches_plain_solve(orders):
    import torch
    import field_interpolation_b200 as fi
    from field_interpolation_b200 import dist as fid

    class OneRank:  # the little of torch.distributed the bootstrap uses
        @staticmethod
        def get_backend():
            return "gloo"

        @staticmethod
        def broadcast(t, src=0):
            return None

    sizes = [64, 24, 19]
    cloud = W.sphere_torus_3d(3000, seed=11)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    weights = fi.Weights(**orders)
    runner = fid.SlabRunner(sizes, weights, 0, 1, OneRank)
    f = fi.sdf_from_points(sizes, weights, pos, cloud["normals"])
    # outputs are float32: the fp64 runs may differ by an ulp of the stored float where the two paths sum in another order
    for prec, tol in ((fi.FI_F64, 2e-8), (fi.FI_F32, 1e-5)):
        for its in (1, 3, 30):
            opt = fi.solve_options(prec, its, 1e-30, check_every=4)
            out = np.zeros(runner.local_cells, np.float32)
            st = runner.step(pos, cloud["normals"], opt, out)
            ref, st1 = f.solve(opt)
            assert st["iterations"] == st1["iterations"] == its
            assert np.linalg.norm(out - ref) <= tol * np.linalg.norm(ref)
            assert abs(st["true_residual"] - st1["true_residual"]) <= 1e-4 * st1["true_residual"]
    # device buffers and a warm start
    d_pos, d_nrm = torch.from_numpy(pos).cuda(), torch.from_numpy(cloud["normals"]).cuda()
    guess = torch.from_numpy(ref).cuda()
    out = torch.zeros(runner.local_cells, device="cuda")
    st = runner.step(d_pos, d_nrm, fi.solve_options(fi.FI_F32, 5, 1e-30), out, guess=guess)
    ref2, _ = f.solve(fi.solve_options(fi.FI_F32, 5, 1e-30), guess=ref)
    assert np.linalg.norm(out.cpu().numpy() - ref2) <= 1e-5 * np.linalg.norm(ref2)
    runner.close()
