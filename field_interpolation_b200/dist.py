"""Multi-GPU host glue: one process per GPU (torchrun), the 3D lattice z-slab sharded across ranks.

The compute and the NCCL traffic (halo planes, scalar all-reduces) live in libfi_b200.so (csrc/dist.cu); this
module only bootstraps the communicator — rank 0 obtains the 128-byte NCCL id from the library and
``torch.distributed`` broadcasts it — and wraps ``fi_slab_sdf_solve``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .api import Weights, _Buf, _common_loc, _is_device, solve_options


def slab_range(nz: int, world: int, rank: int):
    """Planes [z0, z1) owned by `rank` (fi_slab_range: contiguous, balanced to one plane).  Host only."""
    z0, z1 = C.c_int32(0), C.c_int32(0)
    L.check(L.lib().fi_slab_range(int(nz), int(world), int(rank), C.byref(z0), C.byref(z1)))
    return int(z0.value), int(z1.value)


def balanced_cuts(sizes, world: int, positions, point_weight: float = 0.0, min_planes: int = 4):
    """fi_slab_balanced_cuts: world + 1 plane numbers that balance lattice planes plus data points over the ranks
    (positions: lattice coordinates, numpy or CUDA tensor).  The same on every rank for the same cloud."""
    p = _Buf(positions)
    n = (p.keep.numel() if p.loc == L.FI_DEVICE else p.keep.size) // 3 if p.keep is not None else 0
    sz = (C.c_int32 * 3)(*[int(v) for v in sizes])
    cuts = (C.c_int32 * (int(world) + 1))()
    L.check(L.lib().fi_slab_balanced_cuts(sz, int(world), n, p.ptr, p.loc if p.loc is not None else L.FI_HOST, float(point_weight), int(min_planes), cuts))
    return [int(v) for v in cuts]


def slab_mg_plan(sizes, world: int, stencil_radius: int = 2, gather_cells: int = 0):
    """How a multigrid-preconditioned slab solve shards its V-cycle (fi_slab_mg_plan, host only): a dict with
    `halo`, `sharded_levels`, `sizes[l]` and `own[l][rank] = (z0, z1)` for l = 0 .. sharded_levels (the last level
    listed is the first replicated one; its ranges say who restricts into which planes)."""
    sz = (C.c_int32 * 3)(*[int(v) for v in sizes])
    nd, halo = C.c_int32(0), C.c_int32(0)
    lsz = (C.c_int32 * 27)()
    rng = (C.c_int32 * (18 * int(world)))()
    L.check(L.lib().fi_slab_mg_plan(sz, int(world), int(stencil_radius), int(gather_cells), C.byref(nd), C.byref(halo), lsz, rng))
    levels = nd.value + 1
    return {"halo": int(halo.value), "sharded_levels": int(nd.value),
            "sizes": [[int(lsz[3 * l + d]) for d in range(3)] for l in range(levels)],
            "own": [[(int(rng[2 * (l * world + k)]), int(rng[2 * (l * world + k) + 1])) for k in range(world)] for l in range(levels)]}


def broadcast_unique_id(dist, rank: int, device=None) -> bytes:
    """Rank 0 asks the library for an NCCL id; everyone receives it through torch.distributed."""
    import torch
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        L.check(L.lib().fi_comm_unique_id(buf, 128))
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


class SlabComm:
    def __init__(self, rank: int, world: int, unique_id: bytes):
        self.rank, self.world = int(rank), int(world)
        h = C.c_void_p()
        idbuf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        L.check(L.lib().fi_comm_create(self.rank, self.world, idbuf, C.byref(h)))
        self._h = h

    def set_cuts(self, nz: int, cuts=None):
        """fi_comm_set_slab_cuts: a non-uniform partition for later solves on nz-plane lattices (None: uniform again)."""
        arr = None if cuts is None else (C.c_int32 * (self.world + 1))(*[int(v) for v in cuts])
        L.check(L.lib().fi_comm_set_slab_cuts(self._h, int(nz), arr))

    def close(self):
        if getattr(self, "_h", None):
            L.lib().fi_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SlabRunner:
    """sdf_from_points + PCG with the lattice sharded over the ranks of a torch.distributed group."""

    def __init__(self, sizes, weights: Weights, rank: int, world: int, dist, cuts=None):
        import torch
        assert len(sizes) == 3, "slab sharding is for 3D lattices"
        self.sizes, self.weights, self.rank, self.world = [int(s) for s in sizes], weights, rank, world
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else None
        self.comm = SlabComm(rank, world, broadcast_unique_id(dist, rank, dev))
        self.set_cuts(cuts)

    def set_cuts(self, cuts=None):
        """cuts: world + 1 plane numbers (balanced_cuts) or None for the uniform partition; fixes this rank's planes."""
        self.comm.set_cuts(self.sizes[2], cuts)
        self.cuts = None if cuts is None else [int(v) for v in cuts]
        self.z0, self.z1 = (self.cuts[self.rank], self.cuts[self.rank + 1]) if cuts is not None else slab_range(self.sizes[2], self.world, self.rank)
        self.local_cells = (self.z1 - self.z0) * self.sizes[0] * self.sizes[1]

    def step(self, positions, normals, options=None, out=None, guess=None, point_weights=None):
        """Every rank passes the whole cloud (lattice coordinates); returns this rank's solve stats, `out` holds
        its owned planes."""
        opt = options if options is not None else solve_options()
        p, nr, pw = _Buf(positions), _Buf(normals), _Buf(point_weights)
        n = (p.keep.numel() if p.loc == L.FI_DEVICE else p.keep.size) // 3
        loc = _common_loc([p, nr, pw])
        if out is None:
            out = np.empty(self.local_cells, np.float32)
        o, g = _Buf(out, self.local_cells), _Buf(guess, self.local_cells if guess is not None else None)
        if g.loc is not None and g.loc != o.loc:
            raise ValueError("guess and out must live in the same place")
        sz = (C.c_int32 * 3)(*self.sizes)
        w = self.weights.c()
        st = L.fi_solve_stats()
        L.check(L.lib().fi_slab_sdf_solve(self.comm._h, sz, C.byref(w), n, p.ptr, nr.ptr, pw.ptr, loc, C.byref(opt), g.ptr, o.ptr,
                                          o.loc, C.byref(st)))
        return st.as_dict()

    def close(self):
        self.comm.close()
