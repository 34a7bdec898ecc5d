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
