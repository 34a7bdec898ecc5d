"""GPU: the matrix-free normal equations and the PCG solves against the oracle.

Tolerances (BASELINE.json north_star): relative L2 error of the solved field vs the reference's exact solve
<= 1e-5 in fp64 and <= 1e-3 in fp32.  The "exact solve" is the fp64 sparse direct solve of the
reference-assembled rows (golden fixtures / oracle), see oracle/fi_oracle.cpp header for why not Eigen."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, weights_kwargs
from field_interpolation_b200 import workloads as W
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL_F64, TOL_F32 = 1e-5, 1e-3


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def fi():
    import field_interpolation_b200 as m
    return m


ALL_ON = dict(model_0=0.1, model_1=0.2, model_2=0.5, model_3=0.3, model_4=0.25, gradient_smoothness=0.15)


@pytest.mark.parametrize("sizes", [[1], [2], [5], [9], [10], [37], [7, 6], [1, 9], [12, 3], [7, 6, 9], [3, 2, 4], [13, 11, 10], [10, 1, 10]])
@pytest.mark.parametrize("fast", [False, True])
def test_operator_matches_reference_normal_equations(fi, port, sizes, fast):
    """y = AtA x, Atb and diag(AtA) vs the oracle's explicit normal equations; every weight on, every kernel."""
    D, n = len(sizes), int(np.prod(sizes))
    rng = np.random.default_rng(n)
    for vk, gk in ((1, 1), (0, 0), (1, 2)):
        pos, nrm = W.random_cloud(D, 150, sizes, seed=n + gk)
        kw = dict(ALL_ON, value_kernel=vk, gradient_kernel=gk)
        f = fi.sdf_from_points(sizes, fi.Weights(**kw), pos, nrm)
        fi.add_equation(f, 0.7, 1.5, [(0, 1.0), (n - 1, -2.0), (0, 0.5)])  # duplicate column inside one row
        ref = port.sdf_from_points(sizes, O.make_weights(**kw), pos, nrm)
        ref.add_equation(0.7, 1.5, [0, n - 1, 0], [1.0, -2.0, 0.5])
        M, atb = O.normal_equations_f64(ref.system(), n)
        x = rng.normal(size=n)
        scale = abs(M).max() * np.abs(x).max()
        np.testing.assert_allclose(f.rhs(fi.FI_F64), atb, rtol=0, atol=1e-12 * max(1, np.abs(atb).max()))
        np.testing.assert_allclose(f.diagonal(fi.FI_F64), M.diagonal(), rtol=1e-12, atol=1e-13)
        # the fast flag only matters inside solve(); apply() always uses the operator's current setting
        np.testing.assert_allclose(f.apply(x, fi.FI_F64), M @ x, rtol=0, atol=1e-12 * scale * 30)
        np.testing.assert_allclose(f.apply(x.astype(np.float32), fi.FI_F32), M @ x, rtol=0, atol=2e-6 * scale * 30)
        np.testing.assert_allclose(f.rhs(fi.FI_F32), atb, rtol=0, atol=1e-6 * max(1, np.abs(atb).max()))


@pytest.mark.parametrize("sizes", [[32, 8, 8], [64, 24, 17], [36, 9, 40], [128, 10, 9], [160, 19, 33]])
@pytest.mark.parametrize("orders", [dict(model_1=0.7), dict(model_2=0.5), dict(model_0=0.2, model_1=0.3, model_2=0.5),
                                    dict(model_3=0.4), dict(model_0=0.1, model_1=0.2, model_2=0.5, model_3=0.3, model_4=0.25),
                                    dict(model_2=0.5, gradient_smoothness=0.3), dict(model_1=0.2, gradient_smoothness=0.4),
                                    dict(model_0=0.1, model_1=0.2, model_2=0.5, model_3=0.3, model_4=0.25, gradient_smoothness=0.15)])
def test_fast_stencil_matches_generic_and_oracle(fi, port, sizes, orders):
    """The TMA-staged 3D kernel (mode 1) and the tiled one without TMA (mode 2) against the generic kernel (mode 0)
    and the explicit AtA, fp32 and fp64, radius 1 / 2 / 4, lattices with ragged tiles and short z chunks; with the
    gradient-smoothness cross terms (field_interpolation.cpp:303-315) the TMA kernel's GS variant runs (mode 2 falls back
    to the generic kernel there)."""
    n = int(np.prod(sizes))
    rng = np.random.default_rng(n)
    kw = dict(model_2=0.0)
    kw.update(orders)
    pos, nrm = W.random_cloud(3, 200, sizes, seed=n)
    f = fi.sdf_from_points(sizes, fi.Weights(**kw), pos, nrm)
    M, _ = O.normal_equations_f64(port.sdf_from_points(sizes, O.make_weights(**kw), pos, nrm).system(), n)
    x = rng.normal(size=n)
    want = M @ x
    scale = abs(M).max() * np.abs(x).max() * 30
    for enable in (1, 2, 0):
        f.use_fast_stencil(enable)
        np.testing.assert_allclose(f.apply(x, fi.FI_F64), want, rtol=0, atol=1e-12 * scale)
        np.testing.assert_allclose(f.apply(x.astype(np.float32), fi.FI_F32), want, rtol=0, atol=2e-6 * scale)


@pytest.mark.parametrize("sizes", [[32, 8], [64, 17], [36, 40], [128, 33], [260, 50], [48, 9]])
@pytest.mark.parametrize("orders", [dict(model_1=0.7), dict(model_2=0.5), dict(model_1=0.1, model_2=1.0), dict(model_3=0.4),
                                    dict(model_2=0.3, gradient_smoothness=0.2), dict(gradient_smoothness=0.5),
                                    dict(model_0=0.1, model_1=0.2, model_2=0.5, model_3=0.3, model_4=0.25, gradient_smoothness=0.15)])
def test_fast_stencil_2d_matches_generic_and_oracle(fi, port, sizes, orders):
    """The TMA-staged 2D kernel (mode 1; stencil_2d.cu) against the generic kernel (mode 0) and the explicit AtA of the
    reference's rows: fp32 and fp64, radius 1 / 2 / 4, gradient-smoothness cross terms alone and with every order on
    (KAT-4 style), ragged tiles in x and y, lattices narrower than a tile."""
    n = int(np.prod(sizes))
    rng = np.random.default_rng(n)
    kw = dict(model_2=0.0)
    kw.update(orders)
    pos, nrm = W.random_cloud(2, 200, sizes, seed=n)
    f = fi.sdf_from_points(sizes, fi.Weights(**kw), pos, nrm)
    M, _ = O.normal_equations_f64(port.sdf_from_points(sizes, O.make_weights(**kw), pos, nrm).system(), n)
    x = rng.normal(size=n)
    want = M @ x
    scale = abs(M).max() * np.abs(x).max() * 30
    for enable in (1, 0):
        f.use_fast_stencil(enable)
        np.testing.assert_allclose(f.apply(x, fi.FI_F64), want, rtol=0, atol=1e-12 * scale)
        np.testing.assert_allclose(f.apply(x.astype(np.float32), fi.FI_F32), want, rtol=0, atol=2e-6 * scale)


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("orders", [dict(), dict(model_1=0.1, model_2=1.0, gradient_smoothness=0.3)])
def test_fused_pcg_2d_matches_unfused(fi, prec, orders):
    """2D: the fused direction+stencil iteration (stencil_2d.cu, p ping-pong) vs the three-kernel iteration with the generic
    kernel: same iterates, same iteration counts."""
    sizes = [96, 70]
    cloud = W.circles_2d(600, seed=4)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    f = fi.sdf_from_points(sizes, fi.Weights(**orders), pos, cloud["normals"])
    P = fi.FI_F32 if prec == "f32" else fi.FI_F64
    for its in (1, 2, 7, 40):
        xa, sa = f.solve(fi.solve_options(P, its, 1e-30, check_every=4, use_fast_stencil=1))
        xb, sb = f.solve(fi.solve_options(P, its, 1e-30, check_every=4, use_fast_stencil=False))
        assert sa["iterations"] == sb["iterations"] == its
        assert rel(xa, xb.astype(np.float64)) <= (2e-4 if prec == "f32" else 1e-10)
    t = f.time_iterations(4, fi.solve_options(P, 0, 1e-6))
    assert t["fused"]  # the fast path is the one that ran
    xa, sa = f.solve(fi.solve_options(P, 0, 1e-5, use_fast_stencil=1))
    xb, sb = f.solve(fi.solve_options(P, 0, 1e-5, use_fast_stencil=False))
    assert sa["relative_residual"] <= 1e-5 and sb["relative_residual"] <= 1e-5, (sa, sb)
    assert abs(sa["iterations"] - sb["iterations"]) <= max(3, 0.03 * sb["iterations"])


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_fused_pcg_matches_unfused(fi, prec, mode):
    """Fused direction+stencil iteration (p ping-pong; mode 1 TMA-staged, mode 2 tiled) vs the three-kernel
    iteration: same iterates."""
    sizes = [64, 40, 24]
    cloud = W.sphere_torus_3d(3000, seed=9)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, cloud["normals"])
    P = fi.FI_F32 if prec == "f32" else fi.FI_F64
    for its in (1, 2, 7, 40):
        xa, sa = f.solve(fi.solve_options(P, its, 1e-30, check_every=4, use_fast_stencil=mode))
        xb, sb = f.solve(fi.solve_options(P, its, 1e-30, check_every=4, use_fast_stencil=False))
        assert sa["iterations"] == sb["iterations"] == its
        assert rel(xa, xb.astype(np.float64)) <= (2e-4 if prec == "f32" else 1e-10)
    xa, sa = f.solve(fi.solve_options(P, 0, 1e-5, use_fast_stencil=mode))
    xb, sb = f.solve(fi.solve_options(P, 0, 1e-5, use_fast_stencil=False))
    # (the recurrence decides when to stop; `converged` speaks about the residual recomputed from x, which fp32 may miss by a little)
    assert sa["relative_residual"] <= 1e-5 and sb["relative_residual"] <= 1e-5, (sa, sb)
    assert max(sa["true_residual"], sb["true_residual"]) <= (1e-4 if prec == "f32" else 1.25e-5), (sa, sb)
    assert prec == "f32" or (sa["converged"] and sb["converged"])
    assert abs(sa["iterations"] - sb["iterations"]) <= max(3, 0.03 * sb["iterations"])
    assert rel(xa, xb.astype(np.float64)) <= 1e-3


@pytest.mark.parametrize("name", ["kat1_readme_1d", "kat2_field_1d_res12", "kat2_field_1d_res100"])
def test_1d_known_answers(fi, name):
    g = load_golden(name)
    if name.startswith("kat1"):
        f = fi.LatticeField([6])
        fi.add_value_constraint(f, [0.0], 4.0, 1.0)
        fi.add_value_constraint(f, [5.0], 2.0, 1.0)
        fi.add_gradient_constraint(f, [0.0], [1.0], 1.0, 0)
        fi.add_gradient_constraint(f, [5.0], [-1.0], 1.0, 0)
        fi.add_field_constraints(f, fi.Weights(model_2=1.0))
    else:
        res = int(name.split("res")[1])
        c, w = W.field_1d(res), fi.Weights()
        f = fi.LatticeField(c["sizes"])
        for p, v, gr in zip(c["pos"], c["value"], c["gradient"]):
            fi.add_value_constraint(f, p, float(v), w.data_pos)
            fi.add_gradient_constraint(f, p, gr, w.data_gradient, w.gradient_kernel)
        fi.add_field_constraints(f, w)
    x = fi.solve_sparse_linear_exact(f)
    assert rel(x, g["solution"]) <= TOL_F64
    x32 = fi.solve_sparse_linear_fast(f)
    assert rel(x32, g["solution"]) <= TOL_F32


@pytest.mark.parametrize("name", golden_names("rand_"))
def test_randomised_golden_solutions(fi, name):
    g = load_golden(name)
    f = fi.sdf_from_points(g["sizes"], fi.Weights(**weights_kwargs(g["weights"])), g["positions"], g["normals"],
                           g["point_weights"])
    x, st = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-13))
    assert rel(x, g["solution"]) <= TOL_F64, st
    x, st = f.solve(fi.solve_options(fi.FI_MIXED, 0, 1e-9))  # refinement needs cond*eps_fp32 < 1: small lattices only
    assert rel(x, g["solution"]) <= TOL_F32, st
    x, st = f.solve(fi.solve_options(fi.FI_F32, 0, 1e-6))
    assert rel(x, g["solution"]) <= TOL_F32, st


@pytest.mark.parametrize("sizes,npts", [([40, 37], 400), ([24, 20, 22], 1500)])
def test_sdf_solve_vs_exact_and_same_algorithm(fi, port, sizes, npts):
    cloud = W.circles_2d(npts, seed=2) if len(sizes) == 2 else W.sphere_torus_3d(npts, seed=2)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, cloud["normals"])
    sys_ = port.sdf_from_points(sizes, O.make_weights(), pos, cloud["normals"]).system()
    n = int(np.prod(sizes))
    exact = O.exact_solve(sys_, n)
    x64, st = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-12))
    assert st["converged"] and rel(x64, exact) <= TOL_F64
    # like for like: CPU PCG (same algorithm, same stopping rule) at the same tolerance
    N64 = port.normal(sys_, n, "f64")
    for tol in (1e-4, 1e-8):
        xg, sg = f.solve(fi.solve_options(fi.FI_F64, 0, tol))
        xc, itc, errc = N64.pcg(tol=tol)
        assert abs(sg["iterations"] - itc) <= max(2, 0.02 * itc), (sg["iterations"], itc)
        assert rel(xg, xc) <= 1e-5
    x32, st32 = f.solve(fi.solve_options(fi.FI_F32, 0, 1e-6))
    assert rel(x32, exact) <= TOL_F32, st32
    xm, stm = f.solve(fi.solve_options(fi.FI_MIXED, 0, 1e-10))
    assert stm["converged"] and rel(xm, exact) <= TOL_F64, stm


def test_with_guess_and_iteration_cap(fi, port):
    sizes = [30, 28]
    cloud = W.circles_2d(300, seed=4)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, cloud["normals"])
    n = 30 * 28
    exact = O.exact_solve(port.sdf_from_points(sizes, O.make_weights(), pos, cloud["normals"]).system(), n)
    x, st = f.solve(fi.solve_options(fi.FI_F32, 0, 1e-6), guess=exact.astype(np.float32))
    assert st["iterations"] <= 3 and st["initial_residual"] < 1e-4
    x5, st5 = f.solve(fi.solve_options(fi.FI_F32, 5, 1e-12, check_every=2))
    assert st5["iterations"] == 5 and not st5["converged"]           # last iterate returned, not an error
    g = np.zeros(n, np.float32)
    assert fi.solve_tiled_with_guess(f, g[:-1]).size == 0              # incomplete guess => {} (:402-405)
    xt = fi.solve_tiled_with_guess(f, g, sizes, fi.SolveOptions(error_tolerance=1e-6))
    assert rel(xt, exact) <= TOL_F32
    xw = fi.solve_sparse_linear_with_guess(f, g, 0, 1e-6)
    assert rel(xw, exact) <= TOL_F32


def test_jacobi_iterations_vs_port(fi, port):
    sizes = [14, 12]
    pos, nrm = W.random_cloud(2, 120, sizes, 8)
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, nrm)
    N32 = port.normal(port.sdf_from_points(sizes, O.make_weights(), pos, nrm).system(), 14 * 12, "f32")
    g = np.random.default_rng(0).normal(size=14 * 12).astype(np.float32)
    assert np.array_equal(fi.jacobi_iterations(f, g, 0, 0.5), g)
    for its, w in ((1, 0.5), (25, 2.0 / 3.0)):
        got, want = fi.jacobi_iterations(f, g, its, w), N32.jacobi(g, its, w)
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-4 * max(1.0, np.abs(want).max()))


def test_generic_rows_only_system(fi, port):
    """A bare LinearEquation (no lattice structure, the line_2d / bipolar pattern): rows appended by hand."""
    rng = np.random.default_rng(12)
    n, m = 40, 160
    f, ref = fi.LatticeField([n]), port.field([n])
    for _ in range(m):
        k = int(rng.integers(1, 5))
        cols = rng.integers(0, n, k).tolist()
        vals = rng.normal(size=k).astype(np.float32).tolist()
        w, b = float(rng.uniform(0.2, 2)), float(rng.normal())
        fi.add_equation(f, w, b, list(zip(cols, vals)))
        ref.add_equation(w, b, cols, vals)
    s = ref.system()
    e = f.eq
    assert np.array_equal(e.rows, s.rows) and np.array_equal(e.cols, s.cols) and np.array_equal(e.vals, s.vals)
    exact = O.exact_solve(s, n)
    x, st = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-13))
    assert rel(x, exact) <= TOL_F64, st


def test_cascade_matches_manual_recipe(fi, port):
    """fi_sdf_solve_cascade == the reference recipe (src/sdf_field.cpp:272-288) done by hand with the oracle."""
    sizes, factor = [48, 48], 2
    cloud = W.circles_2d(500, seed=6)
    w = fi.Weights()
    x, st = fi.sdf_solve_cascade(sizes, w, cloud["unit_pos"], cloud["normals"], options=fi.solve_options(fi.FI_F64, 0, 1e-12),
                                 factor=factor, coarsest_size=20, coarse_tolerance=1e-12)
    assert st["levels"] == 2 and st["level_cells"] == [48 * 48, 24 * 24]
    # by hand: coarse exact solve -> upscale -> x2 -> that is the fine level's initial guess
    small = [24, 24]
    ps = W.to_lattice(cloud["unit_pos"], small)
    coarse = O.exact_solve(port.sdf_from_points(small, O.make_weights(), ps, cloud["normals"]).system(), 24 * 24)
    guess = port.upscale_field(coarse.astype(np.float32), small, sizes) * np.float32(factor)
    pl = W.to_lattice(cloud["unit_pos"], sizes)
    sys_l = port.sdf_from_points(sizes, O.make_weights(), pl, cloud["normals"]).system()
    M, atb = O.normal_equations_f64(sys_l, 48 * 48)
    want_r0 = np.linalg.norm(atb - M @ guess.astype(np.float64)) / np.linalg.norm(atb)
    assert abs(st["level_initial_residual"][0] - want_r0) <= 1e-3 * want_r0
    assert rel(x, O.exact_solve(sys_l, 48 * 48)) <= TOL_F64
    # and the guess pays: far fewer iterations than from zero
    f = fi.sdf_from_points(sizes, w, pl, cloud["normals"])
    _, st0 = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-6))
    _, stc = fi.sdf_solve_cascade(sizes, w, cloud["unit_pos"], cloud["normals"], options=fi.solve_options(fi.FI_F64, 0, 1e-6),
                                  factor=factor, coarsest_size=20)
    assert stc["level_iterations"][0] < st0["iterations"]


# ---- multigrid-preconditioned CG (FI_PRECOND_MULTIGRID): same system, same stopping rule, different preconditioner ----
_EXACT_CACHE = {}


@pytest.mark.parametrize("sizes,npts", [([97], 20), ([40, 37], 400), ([65, 64], 900), ([24, 20, 22], 1500), ([48, 33, 40], 4000)])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_multigrid_pcg_vs_exact(fi, port, sizes, npts, prec):
    D = len(sizes)
    if D == 1:
        rng = np.random.default_rng(5)
        unit = rng.uniform(0.05, 0.95, (npts, 1)).astype(np.float32)
        nrm = np.where(rng.uniform(size=(npts, 1)) < 0.5, -1.0, 1.0).astype(np.float32)
        cloud = {"unit_pos": unit, "normals": nrm}
    else:
        cloud = W.circles_2d(npts, seed=2) if D == 2 else W.sphere_torus_3d(npts, seed=2)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, cloud["normals"])
    n = int(np.prod(sizes))
    if tuple(sizes) not in _EXACT_CACHE:  # the sparse direct solve of the largest case takes over a minute of host time
        sys_ = port.sdf_from_points(sizes, O.make_weights(), pos, cloud["normals"]).system()
        _EXACT_CACHE[tuple(sizes)] = O.exact_solve(sys_, n)
    exact = _EXACT_CACHE[tuple(sizes)]
    if prec == "f64":
        x, st = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-11, preconditioner=fi.FI_PRECOND_MULTIGRID))
        assert st["converged"] and rel(x, exact) <= TOL_F64, st
        assert st["true_residual"] <= 1e-9, st
    else:
        x, st = f.solve(fi.solve_options(fi.FI_F32, 0, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID))
        assert rel(x, exact) <= TOL_F32, st
    # the point of the preconditioner: far fewer iterations than Jacobi-PCG at the same tolerance
    _, sj = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-6))
    _, sm = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-6, preconditioner=fi.FI_PRECOND_MULTIGRID))
    assert sm["converged"] and sm["iterations"] <= max(8, sj["iterations"] // 2), (sm["iterations"], sj["iterations"])


@pytest.mark.parametrize("orders", [dict(model_2=0.0, model_1=0.7), dict(model_0=0.2, model_1=0.3, model_2=0.5),
                                    dict(model_2=0.0, model_3=0.4), dict(model_2=0.3, gradient_smoothness=0.2)])
def test_multigrid_pcg_other_models(fi, port, orders):
    sizes = [33, 30, 28]
    cloud = W.sphere_torus_3d(2500, seed=8)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    kw = dict(orders)
    f = fi.sdf_from_points(sizes, fi.Weights(**kw), pos, cloud["normals"])
    sys_ = port.sdf_from_points(sizes, O.make_weights(**kw), pos, cloud["normals"]).system()
    exact = O.exact_solve(sys_, int(np.prod(sizes)))
    x, st = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-11, preconditioner=fi.FI_PRECOND_MULTIGRID))
    assert st["converged"] and rel(x, exact) <= TOL_F64, st


def test_multigrid_with_guess_and_cap(fi, port):
    sizes = [40, 40, 40]
    cloud = W.sphere_torus_3d(3000, seed=9)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, cloud["normals"])
    x, st = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-10, preconditioner=fi.FI_PRECOND_MULTIGRID))
    # restarting from the (fp32-rounded) solution needs next to nothing; a cap of 2 iterations returns the last
    # iterate, not an error
    _, st2 = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-8, preconditioner=fi.FI_PRECOND_MULTIGRID), guess=x)
    assert st2["converged"] and st2["iterations"] <= 3 and st2["initial_residual"] <= 1e-4, st2
    _, st3 = f.solve(fi.solve_options(fi.FI_F64, 2, 1e-12, preconditioner=fi.FI_PRECOND_MULTIGRID))
    assert not st3["converged"] and st3["iterations"] == 2, st3


@pytest.mark.parametrize("sizes,npts,wkw", [([3000], 40, {}), ([200, 150], 2000, {}), ([40, 36, 33], 3000, {}),
                                            ([36, 30, 28], 2500, dict(model_1=0.2, gradient_smoothness=0.2))])
def test_multigrid_tail_kernel_matches_per_level_launches(fi, port, sizes, npts, wkw, monkeypatch):
    """The last levels of the cycle walked by one kernel (mg_tail_kernel) against the same levels launched one kernel per
    step (FI_B200_MG_TAIL_CELLS=0): the same preconditioner up to summation order — same iteration count, same field —
    with far fewer launches."""
    D = len(sizes)
    if D == 1:
        rng = np.random.default_rng(5)
        cloud = {"unit_pos": rng.uniform(0.05, 0.95, (npts, 1)).astype(np.float32),
                 "normals": np.where(rng.uniform(size=(npts, 1)) < 0.5, -1.0, 1.0).astype(np.float32)}
    else:
        cloud = W.circles_2d(npts, seed=4) if D == 2 else W.sphere_torus_3d(npts, seed=4)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    f = fi.sdf_from_points(sizes, fi.Weights(**wkw), pos, cloud["normals"])
    opt = fi.solve_options(fi.FI_F64, 0, 1e-9, preconditioner=fi.FI_PRECOND_MULTIGRID)
    out = {}
    for cells in ("0", "4096"):
        monkeypatch.setenv("FI_B200_MG_TAIL_CELLS", cells)
        f.solve(opt)  # builds the hierarchy and its graph
        fi.kernel_launches_reset()
        x, st = f.solve(opt)
        out[cells] = (x, st, fi.kernel_launches())
    (x0, st0, n0), (x1, st1, n1) = out["0"], out["4096"]
    assert st0["converged"] and st1["converged"], (st0, st1)
    assert abs(st0["iterations"] - st1["iterations"]) <= 1, (st0, st1)
    assert rel(x1, x0) <= 1e-7, rel(x1, x0)
    assert n1 < 0.9 * n0, (n0, n1)


@pytest.mark.parametrize("sizes,npts,wkw", [([700], 30, {}), ([90, 70], 900, dict(model_1=0.2, gradient_smoothness=0.3)),
                                            ([40, 36, 33], 3000, {}), ([30, 26, 28], 2000, dict(model_2=0.2, model_3=0.3, gradient_smoothness=0.2)),
                                            ([20, 18, 16], 900, dict(model_0=0.1, model_4=0.2))])
def test_multigrid_dense_coarsest_matrix_written_directly(fi, port, sizes, npts, wkw, monkeypatch):
    """The coarsest operator's dense matrix written by dense_stencil_kernel / dense_blocks_kernel against the same matrix
    column by column from operator applications (FI_B200_MG_DENSE_APPLY=1): the same preconditioner — identical iteration
    counts, fields equal to rounding."""
    D = len(sizes)
    if D == 1:
        rng = np.random.default_rng(5)
        cloud = {"unit_pos": rng.uniform(0.05, 0.95, (npts, 1)).astype(np.float32),
                 "normals": np.where(rng.uniform(size=(npts, 1)) < 0.5, -1.0, 1.0).astype(np.float32)}
    else:
        cloud = W.circles_2d(npts, seed=7) if D == 2 else W.sphere_torus_3d(npts, seed=7)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    opt = fi.solve_options(fi.FI_F64, 0, 1e-9, preconditioner=fi.FI_PRECOND_MULTIGRID)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("FI_B200_MG_DENSE_APPLY", flag)
        f = fi.sdf_from_points(sizes, fi.Weights(**wkw), pos, cloud["normals"])  # a fresh field: the hierarchy is built per field
        out[flag] = f.solve(opt)
        f.close()
    (x0, st0), (x1, st1) = out["1"], out["0"]
    assert st0["converged"] and st1["converged"], (st0, st1)
    assert st0["iterations"] == st1["iterations"], (st0, st1)
    assert rel(x1, x0) <= 1e-8, rel(x1, x0)


# ---- tile phase of solve_tiled_with_guess (tile_solver_square, reference sparse_linear.cpp:246-390) ----------------
def _tile_case(port, sizes, npts, seed, **wkw):
    D = len(sizes)
    if D == 1:
        rng = np.random.default_rng(seed)
        cloud = {"unit_pos": rng.uniform(0.05, 0.95, (npts, 1)).astype(np.float32),
                 "normals": np.where(rng.uniform(size=(npts, 1)) < 0.5, -1.0, 1.0).astype(np.float32)}
    else:
        cloud = W.circles_2d(npts, seed=seed) if D == 2 else W.sphere_torus_3d(npts, seed=seed)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    sys_ = port.sdf_from_points(sizes, O.make_weights(**wkw), pos, cloud["normals"]).system()
    return pos, cloud["normals"], sys_


@pytest.mark.parametrize("sizes,tile,npts,wkw", [
    ([50], 16, 12, {}), ([40, 37], 16, 400, {}), ([33, 20], 8, 200, dict(model_1=0.3, gradient_smoothness=0.2)),
    ([16, 16], 16, 100, {}), ([18, 15, 17], 4, 900, {}), ([24, 20, 22], 8, 1500, dict(gradient_kernel=2)),
    ([21, 22], 5, 300, dict(model_2=0.0, model_3=0.4, value_kernel=0))])
def test_tile_phase_matches_reference_tile_solver(fi, port, sizes, tile, npts, wkw):
    """Tile phase alone (cg off) against the oracle's restatement of tile_solver_square, float and double."""
    pos, nrm, sys_ = _tile_case(port, sizes, npts, 11, **wkw)
    n = int(np.prod(sizes))
    f = fi.sdf_from_points(sizes, fi.Weights(**wkw), pos, nrm)
    rng = np.random.default_rng(3)
    guess = rng.normal(size=n).astype(np.float32)
    opts = fi.SolveOptions(tile=True, tile_size=tile, cg=False)
    want64, fails = port.normal(sys_, n, "f64").tile_solve(guess.astype(np.float64), sizes, tile)
    assert fails == 0
    got64, _, tst = fi.solve_tiled_with_guess(f, guess, sizes, opts, precision=fi.FI_F64, return_stats=True)
    assert tst["converged"], tst
    assert rel(got64, want64) <= TOL_F64, tst
    got32 = fi.solve_tiled_with_guess(f, guess, sizes, opts)
    assert rel(got32, want64) <= TOL_F32
    want32, fails32 = port.normal(sys_, n, "f32").tile_solve(guess, sizes, tile)  # what the reference computes, in float
    assert fails32 == 0 and rel(got32, want32) <= 2 * TOL_F32


def test_tile_phase_then_cg_and_degenerate_options(fi, port):
    sizes, tile = [40, 37], 16
    pos, nrm, sys_ = _tile_case(port, sizes, 400, 5)
    n = int(np.prod(sizes))
    f = fi.sdf_from_points(sizes, fi.Weights(), pos, nrm)
    exact = O.exact_solve(sys_, n)
    guess = np.zeros(n, np.float32)
    # neither phase: the guess comes back (sparse_linear.cpp:423-440 with both flags off)
    assert np.array_equal(fi.solve_tiled_with_guess(f, guess + 2, sizes, fi.SolveOptions(tile=False, cg=False)), guess + 2)
    # incomplete guess -> empty result (:402-405)
    assert fi.solve_tiled_with_guess(f, guess[:-1], sizes, fi.SolveOptions()).size == 0
    # tile + cg: the CG phase starts from the tile solution and reaches the exact solve
    x, st, tst = fi.solve_tiled_with_guess(f, guess, sizes, fi.SolveOptions(tile=True, tile_size=tile, cg=True, error_tolerance=1e-7),
                                           return_stats=True)
    assert rel(x, exact) <= TOL_F32, (st, tst)
    tiled = fi.solve_tiled_with_guess(f, guess, sizes, fi.SolveOptions(tile=True, tile_size=tile, cg=False))
    M, atb = O.normal_equations_f64(sys_, n)
    r_tiled = np.linalg.norm(atb - M @ tiled.astype(np.float64)) / np.linalg.norm(atb)
    assert abs(st["initial_residual"] - r_tiled) <= 1e-3 * r_tiled + 1e-6
    # the operator is back in its normal state after a tile phase: a plain solve still matches
    x2, st2 = f.solve(fi.solve_options(fi.FI_F64, 0, 1e-11))
    assert st2["converged"] and rel(x2, exact) <= TOL_F64


def test_tile_phase_points_only_and_caller_rows(fi, port):
    """No smoothness: tiles no row touches keep the guess (:345-348), untouched nodes of touched tiles become 0; plus
    caller-appended rows that span tiles (generic-rows path)."""
    sizes, tile = [16, 16], 4
    of = port.field(sizes)
    of.add_value_constraint([1.5, 2.5], 3.0, 1.0)
    of.add_value_constraint([2.25, 1.75], -1.0, 2.0)
    of.add_value_constraint([3.5, 3.5], 0.5, 1.0)      # straddles four tiles
    of.add_equation(1.5, 2.0, [5, 200, 77], [1.0, -1.0, 0.5])  # a caller row across three tiles
    sys_ = of.system()
    f = fi.LatticeField(sizes)
    for p, v, w in (([1.5, 2.5], 3.0, 1.0), ([2.25, 1.75], -1.0, 2.0), ([3.5, 3.5], 0.5, 1.0)):
        assert fi.add_value_constraint(f, p, v, w)
    fi.add_equation(f, 1.5, 2.0, [(5, 1.0), (200, -1.0), (77, 0.5)])
    guess = np.arange(256, dtype=np.float32) / 16
    want, fails = port.normal(sys_, 256, "f64").tile_solve(guess.astype(np.float64), sizes, tile)
    got = fi.solve_tiled_with_guess(f, guess, sizes, fi.SolveOptions(tile=True, tile_size=tile, cg=False), precision=fi.FI_F64)
    assert fails == 0 and rel(got, want) <= TOL_F64
    untouched_tiles = np.isclose(want, guess.astype(np.float64)) & (guess != 0)
    assert untouched_tiles.sum() > 100 and np.array_equal(got[untouched_tiles], guess[untouched_tiles])
