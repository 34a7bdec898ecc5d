"""GPU: the triplet view exported by the CUDA path is bit-identical to the reference's (golden fixtures frozen
from the reference build, plus the CPU port on seeded random inputs).  Everything goes through the C ABI."""
import numpy as np
import pytest

from conftest import assert_system_bit_exact, bits, golden_names, load_golden, weights_kwargs
from field_interpolation_b200 import workloads as W
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fi():
    import field_interpolation_b200 as m
    return m


def test_kat1_readme_system(fi):
    g = load_golden("kat1_readme_1d")
    f = fi.LatticeField([6])
    ret = [fi.add_value_constraint(f, [0.0], 4.0, 1.0), fi.add_value_constraint(f, [5.0], 2.0, 1.0),
           fi.add_gradient_constraint(f, [0.0], [1.0], 1.0, fi.GradientKernel.kNearestNeighbor),
           fi.add_gradient_constraint(f, [5.0], [-1.0], 1.0, fi.GradientKernel.kNearestNeighbor)]
    assert ret == [True, True, True, False]
    fi.add_field_constraints(f, fi.Weights(model_2=1.0))
    assert f.counts() == (7, 17)
    assert_system_bit_exact(f.eq, g["rows"], g["cols"], g["vals"], g["rhs"])


@pytest.mark.parametrize("res", [12, 100])
def test_kat2_field_1d(fi, res):
    g = load_golden(f"kat2_field_1d_res{res}")
    c, w = W.field_1d(res), fi.Weights()
    f = fi.LatticeField(c["sizes"])
    for p, v, gr in zip(c["pos"], c["value"], c["gradient"]):
        assert fi.add_value_constraint(f, p, float(v), w.data_pos)
        assert fi.add_gradient_constraint(f, p, gr, w.data_gradient, w.gradient_kernel)
    fi.add_field_constraints(f, w)
    assert_system_bit_exact(f.eq, g["rows"], g["cols"], g["vals"], g["rhs"])


def test_kat3_structure_counts(fi):
    for n, rows, trips in load_golden("kat3_counts")["counts"]:
        rng = np.random.default_rng(int(n))
        pos = rng.uniform(0.01, n - 1.01, size=(1000, 3)).astype(np.float32)
        nrm = rng.normal(size=(1000, 3)).astype(np.float32)
        f = fi.sdf_from_points([int(n)] * 3, fi.Weights(), pos, nrm)
        assert f.counts() == (rows, trips)


def test_kat4_all_weights(fi):
    g = load_golden("kat4_all_weights_7x6x9")
    f = fi.LatticeField([7, 6, 9])
    fi.add_field_constraints(f, fi.Weights(**weights_kwargs(g["weights"])))
    assert f.counts() == (5756, 17354)
    assert_system_bit_exact(f.eq, g["rows"], g["cols"], g["vals"], g["rhs"])


@pytest.mark.parametrize("name", golden_names("rand_"))
def test_randomised_golden(fi, name):
    g = load_golden(name)
    f = fi.sdf_from_points(g["sizes"], fi.Weights(**weights_kwargs(g["weights"])), g["positions"], g["normals"],
                           g["point_weights"])
    assert_system_bit_exact(f.eq, g["rows"], g["cols"], g["vals"], g["rhs"])


def test_values_only_and_appended_rows(fi):
    g = load_golden("values_only_plus_rows_2d")
    f = fi.sdf_from_points(g["sizes"], fi.Weights(**weights_kwargs(g["weights"])), g["positions"], None, None)
    fi.add_equation(f, 0.001, 3.25, [(5, 1.0)])
    fi.add_equation(f, 0.5, -1.0, [(0, 1.0), (71, -1.0)])
    fi.add_equation(f, 0.0, 1.0, [(3, 1.0)])      # zero weight: dropped (sparse_linear.cpp:37)
    fi.add_equation(f, 1.0, 1.0, [(3, 0.0)])      # all-zero row: dropped (:47)
    assert_system_bit_exact(f.eq, g["rows"], g["cols"], g["vals"], g["rhs"])


@pytest.mark.parametrize("sizes", [[13], [2], [1], [9, 7], [1, 5], [6, 5, 7], [3, 3, 3], [2, 9, 2], [33, 17, 9]])
@pytest.mark.parametrize("vk", [0, 1])
@pytest.mark.parametrize("gk", [0, 1, 2])
def test_random_vs_port(fi, port, sizes, vk, gk):
    """Odd sizes, points on and over every boundary, exact lattice hits, zero per-point weights, all kernels."""
    D = len(sizes)
    seed = 1000 * D + 10 * vk + gk + sum(sizes)
    pos, nrm = W.random_cloud(D, 700, sizes, seed)
    pw = np.random.default_rng(seed).uniform(0, 2, 700).astype(np.float32)
    pw[::5] = 0
    kw = dict(model_0=0.3 * (seed % 2), model_1=0.2, model_2=0.5, model_3=0.1 * (seed % 3 == 0), model_4=0.05,
              gradient_smoothness=0.2 * (seed % 2 == 0), value_kernel=vk, gradient_kernel=gk)
    for weights_arr in (None, pw):
        f = fi.sdf_from_points(sizes, fi.Weights(**kw), pos, nrm, weights_arr)
        want = port.sdf_from_points(sizes, O.make_weights(**kw), pos, nrm, weights_arr).system()
        assert f.counts() == (want.num_rows, want.num_triplets)
        assert_system_bit_exact(f.eq, want.rows, want.cols, want.vals, want.rhs)


def test_single_point_builders_vs_port(fi, port):
    rng = np.random.default_rng(5)
    for sizes in ([10], [6, 7], [4, 5, 6]):
        D = len(sizes)
        fa, fb = fi.LatticeField(sizes), port.field(sizes)
        pos, nrm = W.random_cloud(D, 60, sizes, 77)
        for p, g in zip(pos, nrm):
            v, w = float(rng.normal()), float(rng.choice([0.0, 0.5, 1.0, 2.0]))
            k = int(rng.integers(0, 3))
            assert fi.add_value_constraint(fa, p, v, w) == fb.add_value_constraint(p, v, w)
            assert fi.add_value_constraint_nearest_neighbor(fa, p, g, v, w) == fb.add_value_constraint_nearest_neighbor(p, g, v, w)
            assert fi.add_gradient_constraint(fa, p, g, w, k) == fb.add_gradient_constraint(p, g, w, k)
        sb = fb.system()
        assert_system_bit_exact(fa.eq, sb.rows, sb.cols, sb.vals, sb.rhs)


def test_empty_and_degenerate_inputs(fi):
    f = fi.LatticeField([5, 4])
    assert f.counts() == (0, 0) and f.eq.num_rows == 0
    assert fi.add_points(f, 1.0, 1, 1.0, 1, np.zeros((0, 2), np.float32)) == 0
    fi.add_field_constraints(f, fi.Weights(model_2=0.0))  # every weight zero: no rows
    assert f.counts() == (0, 0)
    x, st = f.solve()                                       # zero rhs => x = 0 (Eigen semantics)
    assert not x.any() and st["iterations"] == 0
    with pytest.raises(fi.FiError):
        fi.LatticeField([0, 3])
    with pytest.raises(fi.FiError):
        fi.LatticeField([2, 2, 2, 2])
    with pytest.raises(fi.FiError):  # nearest-neighbour value kernel without normals: the reference CHECK-aborts
        fi.sdf_from_points([4, 4], fi.Weights(value_kernel=0), np.ones((3, 2), np.float32), None)
    with pytest.raises(fi.FiError):
        fi.add_gradient_constraint(fi.LatticeField([4]), [1.0], [1.0], 1.0, 7)


def test_device_resident_inputs(fi, port):
    import torch
    sizes = [12, 10, 9]
    pos, nrm = W.random_cloud(3, 500, sizes, 3)
    f = fi.sdf_from_points(sizes, fi.Weights(), torch.from_numpy(pos).cuda(), torch.from_numpy(nrm).cuda())
    want = port.sdf_from_points(sizes, O.make_weights(), pos, nrm).system()
    assert_system_bit_exact(f.eq, want.rows, want.cols, want.vals, want.rhs)


@pytest.mark.parametrize("name", golden_names("upscale_"))
def test_upscale_golden(fi, name):
    g = load_golden(name)
    out = fi.upscale_field(g["small"], g["small_sizes"], g["large_sizes"])
    assert np.array_equal(bits(out), bits(g["large"]))


def test_upscale_vs_port_large(fi, port):
    rng = np.random.default_rng(1)
    for small, large in (([17, 23], [40, 61]), ([9, 11, 13], [33, 30, 41]), ([16, 16, 16], [32, 32, 32])):
        src = rng.normal(size=int(np.prod(small))).astype(np.float32)
        assert np.array_equal(bits(fi.upscale_field(src, small, large)), bits(port.upscale_field(src, small, large)))
