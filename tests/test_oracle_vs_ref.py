"""CPU, build container only: the oracle port against the reference's own code (oracle/_ref), bit for bit,
on randomised inputs that hit every skip rule.  Skipped where oracle/_ref has not been built."""
import numpy as np
import pytest

from conftest import assert_system_bit_exact, bits
from field_interpolation_b200 import workloads as W
from oracle import oracle as O


@pytest.mark.parametrize("sizes", [[13], [2], [9, 7], [1, 5], [6, 5, 7], [3, 3, 3], [2, 9, 2]])
@pytest.mark.parametrize("vk", [0, 1])
@pytest.mark.parametrize("gk", [0, 1, 2])
def test_sdf_from_points_matches_reference(port, ref, sizes, vk, gk):
    D = len(sizes)
    seed = 1000 * D + 10 * vk + gk + sum(sizes)
    pos, nrm = W.random_cloud(D, 300, sizes, seed)
    pw = np.random.default_rng(seed).uniform(0, 2, 300).astype(np.float32)
    pw[::5] = 0
    w = O.make_weights(model_0=0.3 * (seed % 2), model_1=0.2, model_2=0.5, model_3=0.1 * (seed % 3 == 0),
                       model_4=0.05, gradient_smoothness=0.2 * (seed % 2 == 0), value_kernel=vk, gradient_kernel=gk)
    for weights_arr in (None, pw):
        a = port.sdf_from_points(sizes, w, pos, nrm, weights_arr).system()
        b = ref.sdf_from_points(sizes, w, pos, nrm, weights_arr).system()
        assert_system_bit_exact(a, b.rows, b.cols, b.vals, b.rhs)


def test_single_point_builders_match_reference(port, ref):
    rng = np.random.default_rng(5)
    for sizes in ([10], [6, 7], [4, 5, 6]):
        D = len(sizes)
        fa, fb = port.field(sizes), ref.field(sizes)
        pos, nrm = W.random_cloud(D, 200, sizes, 77)
        for p, g in zip(pos, nrm):
            v, w = float(rng.normal()), float(rng.choice([0.0, 0.5, 1.0, 2.0]))
            k = int(rng.integers(0, 3))
            assert fa.add_value_constraint(p, v, w) == fb.add_value_constraint(p, v, w)
            assert fa.add_value_constraint_nearest_neighbor(p, g, v, w) == fb.add_value_constraint_nearest_neighbor(p, g, v, w)
            assert fa.add_gradient_constraint(p, g, w, k) == fb.add_gradient_constraint(p, g, w, k)
        sa, sb = fa.system(), fb.system()
        assert_system_bit_exact(sa, sb.rows, sb.cols, sb.vals, sb.rhs)


def test_upscale_and_error_map_match_reference(port, ref):
    rng = np.random.default_rng(9)
    for small, large in (([3], [10]), ([5, 3], [11, 8]), ([4, 3, 5], [9, 8, 13]), ([8, 8, 8], [16, 16, 16])):
        src = rng.normal(size=int(np.prod(small))).astype(np.float32)
        assert np.array_equal(bits(port.upscale_field(src, small, large)), bits(ref.upscale_field(src, small, large)))
    s = ref.sdf_from_points([8, 9], O.make_weights(), *W.random_cloud(2, 100, [8, 9], 3)).system()
    sol = rng.normal(size=72).astype(np.float32)
    assert np.array_equal(bits(port.generate_error_map(s, sol)), bits(ref.generate_error_map(s, sol)))


def test_iso_surface_helpers_match_reference(port, ref):
    """Randomised: fields with exact zeros, negative zeros, saddles and 1-wide shapes."""
    rng = np.random.default_rng(21)
    for h, w in ((2, 2), (5, 7), (33, 20), (1, 5), (6, 1), (64, 64), (3, 200)):
        a = rng.standard_normal((h, w)).astype(np.float32)
        a[rng.random((h, w)) < 0.1] = 0.0
        a[rng.random((h, w)) < 0.05] = np.float32(-0.0)
        for iso in (0.0, 0.3, -1.5):
            r, p = ref.iso_surface(a, iso), port.iso_surface(a, iso)
            assert r.shape == p.shape and np.array_equal(bits(r), bits(p))
            assert ref.calc_area(r) == port.calc_area(p)
        assert np.array_equal(bits(ref.marching_squares(a)), bits(port.marching_squares(a)))
        for up in (2, 3, 7):
            assert np.array_equal(bits(ref.bicubic_upsample(a, up)), bits(port.bicubic_upsample(a, up)))
