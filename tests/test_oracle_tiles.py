"""CPU: the oracle's restatement of tile_solver_square (reference sparse_linear.cpp:246-390) against two independent
statements of the same mathematics in scipy fp64 — (1) a single tile covering the lattice is the direct solve of
(AtA + 1e-6 I) x = Atb, whatever the guess; (2) in general it is the block-Jacobi system the CUDA path solves:
(B + 1e-6 I) y = Atb - 2 (AtA - B) g with B = AtA restricted to entries inside one tile."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from field_interpolation_b200 import workloads as W
from oracle import oracle as O


def tile_ids(sizes, tile):
    coords = np.meshgrid(*[np.arange(n) for n in sizes], indexing="ij")  # coords[d][i0, i1, ...]
    tiles = [(n + tile - 1) // tile for n in sizes]
    tid = np.zeros(sizes, np.int64)
    ts = 1
    for d in range(len(sizes)):
        tid += (coords[d] // tile) * ts
        ts *= tiles[d]
    return tid.ravel(order="F")  # x fastest


def block_jacobi_reference(M, atb, guess, sizes, tile):
    M = M.tocoo()
    tid = tile_ids(sizes, tile)
    same = tid[M.row] == tid[M.col]
    B = sp.csr_matrix((M.data[same], (M.row[same], M.col[same])), shape=M.shape)
    off = sp.csr_matrix((M.data[~same], (M.row[~same], M.col[~same])), shape=M.shape)
    rhs = atb - 2.0 * (off @ guess)
    n = M.shape[0]
    y = spla.spsolve((B + 1e-6 * sp.identity(n)).tocsc(), rhs)
    # tiles without any stored entry keep the guess
    has = np.zeros(tid.max() + 1, bool)
    has[tid[M.row[same]]] = True
    return np.where(has[tid], y, guess)


@pytest.mark.parametrize("sizes,tile", [([23], 8), ([20, 17], 8), ([16, 16], 16), ([9, 10, 11], 4)])
def test_tile_solver_is_block_jacobi(port, sizes, tile):
    D = len(sizes)
    rng = np.random.default_rng(7)
    if D == 1:
        f = port.field(sizes)
        f.add_field_constraints(O.make_weights())
        for x, v in ((3.2, 1.0), (11.7, -0.5), (19.1, 2.0)):
            f.add_value_constraint([x], v, 1.0)
        sys_ = f.system()
    else:
        cloud = W.circles_2d(150, seed=1) if D == 2 else W.sphere_torus_3d(400, seed=1)
        pos = W.to_lattice(cloud["unit_pos"], sizes)
        sys_ = port.sdf_from_points(sizes, O.make_weights(), pos, cloud["normals"]).system()
    n = int(np.prod(sizes))
    N64 = port.normal(sys_, n, "f64")
    M, atb = N64.csr()
    guess = rng.normal(size=n)
    got, fails = N64.tile_solve(guess, sizes, tile)
    assert fails == 0
    want = block_jacobi_reference(M, atb, guess, sizes, tile)
    assert np.linalg.norm(got - want) <= 1e-8 * np.linalg.norm(want)
    # float restatement (what the reference runs) agrees with the double one to float accuracy
    N32 = port.normal(sys_, n, "f32")
    got32, fails32 = N32.tile_solve(guess.astype(np.float32), sizes, tile)
    assert fails32 == 0
    assert np.linalg.norm(got32 - want) <= 2e-3 * np.linalg.norm(want)


def test_single_tile_is_the_direct_solve(port):
    sizes = [12, 12]
    cloud = W.circles_2d(80, seed=3)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    sys_ = port.sdf_from_points(sizes, O.make_weights(), pos, cloud["normals"]).system()
    N64 = port.normal(sys_, 144, "f64")
    M, atb = N64.csr()
    want = spla.spsolve((M + 1e-6 * sp.identity(144)).tocsc(), atb)
    got, fails = N64.tile_solve(np.full(144, 5.0), sizes, 12)
    assert fails == 0 and np.linalg.norm(got - want) <= 1e-9 * np.linalg.norm(want)


def test_tiles_without_entries_keep_the_guess(port):
    # values only, no smoothness: tiles the points do not touch have no AtA entry at all (sparse_linear.cpp:345-348)
    sizes = [16, 16]
    f = port.field(sizes)
    f.add_value_constraint([1.5, 2.5], 3.0, 1.0)
    f.add_value_constraint([2.25, 1.75], -1.0, 2.0)
    sys_ = f.system()
    N64 = port.normal(sys_, 256, "f64")
    guess = np.arange(256, dtype=np.float64)
    got, fails = N64.tile_solve(guess, sizes, 4)
    tid = tile_ids(sizes, 4)
    assert fails == 0
    assert np.array_equal(got[tid != 0], guess[tid != 0])
    # inside the touched tile: nodes no row touches solve 1e-6 y = 0
    M, atb = N64.csr()
    untouched = (tid == 0) & (M.diagonal() == 0)
    assert untouched.any() and np.all(got[untouched] == 0)
