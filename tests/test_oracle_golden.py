"""CPU: the oracle port (oracle/fi_oracle.cpp) against the golden vectors frozen from the reference build
(tests/golden/make_golden.py) and against SURVEY.md §8c's known answers."""
import numpy as np
import pytest

from conftest import assert_system_bit_exact, bits, golden_names, load_golden, weights_kwargs
from field_interpolation_b200 import workloads as W
from oracle import oracle as O


def test_kat1_readme_system(port):
    g = load_golden("kat1_readme_1d")
    f = port.field([6])
    ret = [f.add_value_constraint([0.0], 4.0, 1.0), f.add_value_constraint([5.0], 2.0, 1.0),
           f.add_gradient_constraint([0.0], [1.0], 1.0, O.GRAD_NEAREST),
           f.add_gradient_constraint([5.0], [-1.0], 1.0, O.GRAD_NEAREST)]
    assert ret == [True, True, True, False] == g["returns"].astype(bool).tolist()
    f.add_field_constraints(O.make_weights(model_2=1.0))
    s = f.system()
    assert_system_bit_exact(s, g["rows"], g["cols"], g["vals"], g["rhs"])
    assert (s.num_rows, s.num_triplets) == (7, 17)
    want = [3.877193, 4.2631579, 4.1578947, 3.6842105, 2.9649123, 2.122807]
    np.testing.assert_allclose(O.exact_solve(s, 6), want, rtol=1e-6)
    np.testing.assert_allclose(g["solution"], want, rtol=1e-6)


@pytest.mark.parametrize("res", [12, 100])
def test_kat2_field_1d(port, res):
    g = load_golden(f"kat2_field_1d_res{res}")
    c, w = W.field_1d(res), O.make_weights()
    f = port.field(c["sizes"])
    for p, v, gr in zip(c["pos"], c["value"], c["gradient"]):
        assert f.add_value_constraint(p, float(v), w.data_pos)
        assert f.add_gradient_constraint(p, gr, w.data_gradient, w.gradient_kernel)
    f.add_field_constraints(w)
    s = f.system()
    assert_system_bit_exact(s, g["rows"], g["cols"], g["vals"], g["rhs"])
    if res == 12:
        want = [-0.1846154, -0.1006993, -0.0167832, 0.0671329, 0.1230769, 0.151049, 0.151049, 0.1230769,
                0.0671329, -0.0167832, -0.1006993, -0.1846154]
        np.testing.assert_allclose(g["solution"], want, atol=2e-7)
    N = port.normal(s, res, "f64")
    x, it, err = N.pcg(tol=1e-13, max_iter=20 * res)
    assert np.linalg.norm(x - g["solution"]) <= 1e-6 * np.linalg.norm(g["solution"])


def test_kat3_structure_counts(port):
    for n, rows, trips in load_golden("kat3_counts")["counts"]:
        rng = np.random.default_rng(int(n))
        pos = rng.uniform(0.01, n - 1.01, size=(1000, 3)).astype(np.float32)
        nrm = rng.normal(size=(1000, 3)).astype(np.float32)
        s = port.sdf_from_points([int(n)] * 3, O.make_weights(), pos, nrm).system()
        assert (s.num_rows, s.num_triplets) == (rows, trips)
        assert rows == 3 * n * n * (n - 2) + 4 * 1000 and trips == 3 * 3 * n * n * (n - 2) + 32 * 1000


def test_kat4_all_weights(port):
    g = load_golden("kat4_all_weights_7x6x9")
    f = port.field([7, 6, 9])
    f.add_field_constraints(O.make_weights(**weights_kwargs(g["weights"])))
    s = f.system()
    assert (s.num_rows, s.num_triplets) == (5756, 17354)
    assert_system_bit_exact(s, g["rows"], g["cols"], g["vals"], g["rhs"])


@pytest.mark.parametrize("name", golden_names("rand_"))
def test_randomised_cases(port, name):
    g = load_golden(name)
    w = O.make_weights(**weights_kwargs(g["weights"]))
    s = port.sdf_from_points(g["sizes"], w, g["positions"], g["normals"], g["point_weights"]).system()
    assert_system_bit_exact(s, g["rows"], g["cols"], g["vals"], g["rhs"])
    n = int(np.prod(g["sizes"]))
    x = O.exact_solve(s, n)
    np.testing.assert_allclose(x, g["solution"], rtol=1e-9, atol=1e-12)
    xp, it, err = port.normal(s, n, "f64").pcg(tol=1e-13, max_iter=50 * n)
    assert np.linalg.norm(xp - g["solution"]) <= 1e-6 * np.linalg.norm(g["solution"])


def test_values_only_and_appended_rows(port):
    g = load_golden("values_only_plus_rows_2d")
    w = O.make_weights(**weights_kwargs(g["weights"]))
    f = port.sdf_from_points(g["sizes"], w, g["positions"], None, None)
    f.add_equation(0.001, 3.25, [5], [1.0])
    f.add_equation(0.5, -1.0, [0, 71], [1.0, -1.0])
    assert_system_bit_exact(f.system(), g["rows"], g["cols"], g["vals"], g["rhs"])


@pytest.mark.parametrize("name", golden_names("upscale_"))
def test_upscale_field(port, name):
    g = load_golden(name)
    out = port.upscale_field(g["small"], g["small_sizes"], g["large_sizes"])
    assert np.array_equal(bits(out), bits(g["large"]))


def test_error_map(port):
    g = load_golden("error_map_2d")
    s = O.System(g["rows"], g["cols"], g["vals"], g["rhs"])
    assert np.array_equal(bits(port.generate_error_map(s, g["sol_in"])), bits(g["heat"]))


def test_solve_half_restatement(port):
    """Normal equations of the port vs scipy, and BiCGSTAB / PCG / Jacobi behaviour on a 3D SDF system."""
    sizes = [12, 11, 10]
    pos, nrm = W.random_cloud(3, 400, sizes, seed=11)
    s = port.sdf_from_points(sizes, O.make_weights(), pos, nrm).system()
    n = int(np.prod(sizes))
    M, atb = port.normal(s, n, "f64").csr()
    M2, atb2 = O.normal_equations_f64(s, n)
    assert abs(M - M2).max() < 1e-12 and abs(atb - atb2).max() < 1e-12
    exact = O.exact_solve(s, n)
    N32 = port.normal(s, n, "f32")
    x, it, err = N32.bicgstab(tol=1e-6)
    assert err <= 1e-6 and np.linalg.norm(x - exact) <= 1e-3 * np.linalg.norm(exact)
    x, it, err = N32.pcg(tol=1e-6)
    assert err <= 1e-6 and np.linalg.norm(x - exact) <= 1e-3 * np.linalg.norm(exact)
    x0 = N32.jacobi(np.zeros(n, np.float32), 0, 0.5)
    assert not x0.any()
    r0 = np.linalg.norm(atb)
    xj = N32.jacobi(np.zeros(n, np.float32), 50, 2.0 / 3.0)
    assert np.linalg.norm(atb - M @ xj) < r0
    # bb == 0 => x = 0 (Eigen semantics)
    f = port.field([8])
    f.add_field_constraints(O.make_weights())
    xz, it, err = port.normal(f.system(), 8, "f32").bicgstab(guess=np.ones(8, np.float32))
    assert it == 0 and not xz.any()


# ---- iso-surface helpers --------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("iso_"))
def test_iso_surface_helpers_match_reference_fixtures(port, name):
    """marching_squares / calc_area (third_party/emilib/emilib/marching_squares.cpp), bicubic_upsample / iso_surface
    (src/sdf_field.cpp:555-614): the port against outputs frozen from the reference's own code, bit for bit."""
    g = load_golden(name)
    field, up = g["field"], int(g["upsample"])
    lines = port.marching_squares(field)
    assert lines.shape == g["lines"].shape and np.array_equal(bits(lines), bits(g["lines"]))
    assert np.float32(port.calc_area(lines)) == g["area"]
    big = port.bicubic_upsample(field, up)
    assert big.shape == g["upsampled"].shape and np.array_equal(bits(big), bits(g["upsampled"]))
    zl = port.iso_surface(big, 0.0)
    assert zl.shape == g["zero_lines_up"].shape and np.array_equal(bits(zl), bits(g["zero_lines_up"]))
    assert np.float32(port.calc_area(zl)) == g["area_up"]
    for src, key in ((big, "iso_lines_up"), (field, "iso_lines")):
        il = port.iso_surface(src, float(g["iso"]))
        assert il.shape == g[key].shape and np.array_equal(bits(il), bits(g[key]))


def test_marching_squares_contour_of_a_disc(port):
    """Known answer independent of the reference: the contour of a radius-40 disc SDF is closed, clockwise in a
    y-down system (negative shoelace sign convention of calc_area gives +area), and encloses ~pi r^2."""
    yy, xx = np.mgrid[0:129, 0:129].astype(np.float32)
    sdf = (np.hypot(xx - 64, yy - 64) - 40).astype(np.float32)
    lines = port.marching_squares(sdf)
    assert abs(port.calc_area(lines) - np.pi * 1600) / (np.pi * 1600) < 2e-3
    starts = {tuple(p) for p in lines[:, :2].view(np.uint32).reshape(-1, 2).tolist()}
    ends = {tuple(p) for p in lines[:, 2:].copy().view(np.uint32).reshape(-1, 2).tolist()}
    assert starts == ends  # every segment end is another segment's start: closed loops
