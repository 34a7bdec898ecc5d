"""The C++ drop-in API (include/field_interpolation/*.hpp + field_interpolation_b200/host/*.cpp over the C ABI).

CPU part: the headers compile as C++14, a program written against the reference's API links, and the host
library exports the reference's entry points.  GPU part: tests/cpp/api_driver.cpp replays the reference's own
callers (src/field_1d.cpp, src/interpolate_2d.cpp, src/sdf_field.cpp, hand-written rows) and its outputs are
compared with the oracle: triplets and right-hand sides bit for bit, solved fields within the stated
tolerance (relative L2 <= 1e-5 for the fp64 "exact" solves)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_system_bit_exact, load_golden
from field_interpolation_b200 import workloads as W
from oracle import oracle as O

DRIVER_SRC = os.path.join(ROOT, "tests", "cpp", "api_driver.cpp")
DRIVER = os.path.join(ROOT, "tests", "cpp", "api_driver")


def _build():
    from field_interpolation_b200 import build as B
    B.build()
    return B.build_cpp_driver(DRIVER_SRC, DRIVER)


def _read(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            head = f.read(28)
            if len(head) < 28:
                break
            name = head[:16].split(b"\0")[0].decode()
            count, dtype = struct.unpack("<qi", head[16:])
            out[name] = np.frombuffer(f.read(4 * count), dtype=np.float32 if dtype == 0 else np.int32).copy()
    return out


def _run(scenario, tmp_path, inputs=None):
    exe = _build()
    out = str(tmp_path / f"{scenario}.bin")
    cmd = [exe, scenario, out]
    if inputs is not None:
        inp = str(tmp_path / f"{scenario}.in")
        with open(inp, "wb") as f:
            for a in inputs:
                a = np.ascontiguousarray(a, np.float32).ravel()
                f.write(struct.pack("<q", a.size))
                f.write(a.tobytes())
        cmd.append(inp)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return _read(out)


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))


# ---- CPU: compile / link / symbols -------------------------------------------------------------------------
def test_headers_compile_and_driver_links():
    exe = _build()
    assert os.path.exists(exe)
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-x", "c++", "-"], input="#include <field_interpolation/field_interpolation.hpp>\n"
                                                  "#include <field_interpolation/sparse_linear.hpp>\nint main(){return 0;}\n",
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_host_library_exports_reference_entry_points():
    from field_interpolation_b200 import build as B
    B.build()
    syms = subprocess.run(["nm", "-D", "-C", "--defined-only", B.HOST_OUT], capture_output=True, text=True).stdout
    for name in ["add_field_constraints", "add_value_constraint(", "add_value_constraint_nearest_neighbor", "add_gradient_constraint",
                 "add_points", "sdf_from_points", "generate_error_map", "upscale_field", "add_equation", "solve_sparse_linear_fast",
                 "solve_sparse_linear_exact", "solve_sparse_linear_with_guess", "jacobi_iterations", "solve_tiled_with_guess",
                 "operator<<(std::ostream&, field_interpolation::LinearEquation const&)", "b200::defer_triplets", "b200::materialize",
                 "b200::sdf_solve_cascade", "b200::solve("]:
        assert f"field_interpolation::{name}" in syms or name in syms, name


# ---- GPU: the reference's callers, replayed ----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("res", [12, 100])
def test_field_1d_caller(tmp_path, res):
    got = _run(f"field_1d_{res}", tmp_path)
    g = load_golden(f"kat2_field_1d_res{res}")
    assert np.array_equal(got["rows"], g["rows"]) and np.array_equal(got["cols"], g["cols"])
    assert np.array_equal(got["vals"].view(np.uint32), g["vals"].view(np.uint32))
    assert np.array_equal(got["rhs"].view(np.uint32), g["rhs"].view(np.uint32))
    assert list(got["accepted"]) == [1, 1, 1, 1]
    assert rel(got["solution"], g["solution"]) <= 1e-5
    assert got["printed"][0] > 0


@pytest.mark.gpu
def test_readme_system(tmp_path):
    got = _run("readme", tmp_path)
    g = load_golden("kat1_readme_1d")
    assert np.array_equal(got["rows"], g["rows"]) and np.array_equal(got["cols"], g["cols"])
    assert np.array_equal(got["vals"].view(np.uint32), g["vals"].view(np.uint32))
    assert np.array_equal(got["rhs"].view(np.uint32), g["rhs"].view(np.uint32))
    assert list(got["accepted"]) == list(g["returns"].astype(int))
    assert rel(got["solution"], g["solution"]) <= 1e-5


@pytest.mark.gpu
def test_interpolate_2d_caller(tmp_path, port):
    got = _run("interpolate_2d", tmp_path)
    res = 24
    values = [5, 4, 2, 3, 4, 2, 1, 5, 6, 3, 5, 2, 1, 2, 4, 1]
    w = O.make_weights()
    f = port.field([res, res])
    f.add_field_constraints(w)
    for y in range(4):
        for x in range(4):
            pos = np.array([np.float32(x) / np.float32(3.0) * np.float32(res - 1.0), np.float32(y) / np.float32(3.0) * np.float32(res - 1.0)], np.float32)
            f.add_value_constraint(pos, float(values[y * 4 + x]), w.data_pos)
            f.add_gradient_constraint(pos, np.zeros(2, np.float32), w.data_gradient, w.gradient_kernel)
    want = f.system()
    assert np.array_equal(got["rows"], want.rows) and np.array_equal(got["cols"], want.cols)
    assert np.array_equal(got["vals"].view(np.uint32), want.vals.view(np.uint32))
    assert np.array_equal(got["rhs"].view(np.uint32), want.rhs.view(np.uint32))
    assert rel(got["solution"], O.exact_solve(want, res * res)) <= 1e-5


@pytest.mark.gpu
def test_sdf_2d_caller_with_border_rows_and_coarse_to_fine(tmp_path, port):
    width, height, bw, factor, npts = 44, 37, 0.5, 2, 600
    cloud = W.circles_2d(npts, seed=5)
    got = _run("sdf_2d", tmp_path, [[width, height, bw, factor], cloud["unit_pos"], cloud["normals"]])

    def oracle_field(w, h):
        pos = (cloud["unit_pos"] * np.array([np.float32(w) - np.float32(1), np.float32(h) - np.float32(1)], np.float32)[None, :]).astype(np.float32)
        f = port.sdf_from_points([w, h], O.make_weights(), pos, cloud["normals"])
        for y in range(h):
            for x in range(w):
                if x in (0, w - 1) or y in (0, h - 1):
                    dx, dy = pos[:, 0] - np.float32(x), pos[:, 1] - np.float32(y)
                    d2 = (dx * dx + dy * dy).astype(np.float32).min()
                    f.add_equation(bw, float(np.sqrt(np.float32(d2))), [y * w + x], [1.0])
        return f.system()

    big = oracle_field(width, height)
    assert np.array_equal(got["rows"], big.rows) and np.array_equal(got["cols"], big.cols)
    assert np.array_equal(got["vals"].view(np.uint32), big.vals.view(np.uint32))
    assert np.array_equal(got["rhs"].view(np.uint32), big.rhs.view(np.uint32))
    exact = O.exact_solve(big, width * height)
    assert rel(got["exact"], exact) <= 1e-5
    ws, hs = (width + factor - 1) // factor, (height + factor - 1) // factor
    small = O.exact_solve(oracle_field(ws, hs), ws * hs)
    assert rel(got["small"], small) <= 1e-5
    assert np.array_equal(got["upscaled"].view(np.uint32), port.upscale_field(got["small"], [ws, hs], [width, height]).view(np.uint32))
    M, atb = O.normal_equations_f64(big, width * height)
    res = np.linalg.norm(M @ got["approx"].astype(np.float64) - atb) / np.linalg.norm(atb)
    assert res <= 2e-4 and rel(got["approx"], exact) <= 5e-2
    assert got["bad_guess"].size == 0
    # tile phase (sparse_linear.cpp:246-390) through the C++ API, against the oracle's restatement
    g = (got["upscaled"] * np.float32(factor)).astype(np.float32)
    want_tiled, fails = port.normal(big, width * height, "f64").tile_solve(g.astype(np.float64), [width, height], 16)
    assert fails == 0 and rel(got["tiled"], want_tiled) <= 1e-3
    res_tc = np.linalg.norm(M @ got["tiled_cg"].astype(np.float64) - atb) / np.linalg.norm(atb)
    assert res_tc <= 2e-4 and rel(got["tiled_cg"], exact) <= 5e-2
    np.testing.assert_allclose(got["heatmap"], port.generate_error_map(big, got["exact"]), rtol=2e-3, atol=1e-7)
    guess = (got["upscaled"] * np.float32(factor)).astype(np.float32)
    np.testing.assert_allclose(got["jacobi"], port.normal(big, width * height, "f32").jacobi(guess, 5, 0.5), rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_hand_written_rows_without_a_lattice(tmp_path, port):
    got = _run("hand_rows", tmp_path)
    n = 9
    f = port.field([n])
    for i in range(n - 1):
        f.add_equation(0.5, 1.0, [i, i + 1], [-1.0, 1.0])
    f.add_equation(2.0, 3.0, [0], [1.0])
    f.add_equation(0.0, 3.0, [0], [1.0])
    f.add_equation(1.0, 3.0, [4], [0.0])
    f.add_equation(1.0, 1.0, [3, 3, 5], [1.0, 0.5, 0.0])
    want = f.system()
    assert np.array_equal(got["rows"], want.rows) and np.array_equal(got["cols"], want.cols)
    assert np.array_equal(got["vals"].view(np.uint32), want.vals.view(np.uint32))
    assert np.array_equal(got["rhs"].view(np.uint32), want.rhs.view(np.uint32))
    exact = O.exact_solve(want, n)
    assert rel(got["exact"], exact) <= 1e-5 and rel(got["fast"], exact) <= 1e-3 and rel(got["guess"], exact) <= 1e-3


@pytest.mark.gpu
def test_copies_are_independent_and_rewritten_equations_are_noticed(tmp_path, port):
    """LatticeField is a value type in the reference (field_interpolation.hpp:97-114): `b = a` copies the system.  Here
    the device description is shared until somebody writes (clone on write), hand-written rows pending on one owner never
    reach the other, and an equation the caller cleared or truncated is rebuilt from its triplet list."""
    got = _run("value_semantics", tmp_path)
    n = 14
    w = O.make_weights()
    p0, p1, p2 = [3.25, 4.5], [9.75, 8.125], [6.5, 2.25]

    def base():
        f = port.field([n, n])
        f.add_field_constraints(w)
        f.add_value_constraint(p0, 1.0, 1.0)
        f.add_value_constraint(p1, -2.0, 1.0)
        f.add_value_constraint([1.5, 11.25], 0.5, 1.0)
        f.add_value_constraint([11.0, 1.75], 3.0, 1.0)
        f.add_value_constraint([6.0, 7.0], -1.0, 1.0)
        return f
    a = base().system()
    assert list(got["a_counts"]) == [a.num_rows, a.num_triplets]
    xa = O.exact_solve(a, n * n)
    assert rel(got["a_exact"], xa) <= 1e-5 and rel(got["a_again"], xa) <= 1e-5 and rel(got["c_exact"], xa) <= 1e-5
    fb = base()
    fb.add_value_constraint(p2, 5.0, 2.0)
    fb.add_equation(1.5, 0.25, [0, n * n - 1], [1.0, -1.0])
    xb = O.exact_solve(fb.system(), n * n)
    assert rel(got["b_exact"], xb) <= 1e-5 and rel(xa, xb) > 1e-2  # the two systems really differ
    fa2 = base()
    fa2.add_equation(3.0, 1.0, [5], [1.0])
    assert rel(got["a_plus_row"], O.exact_solve(fa2.system(), n * n)) <= 1e-5
    fd = port.field([n, n])
    fd.add_field_constraints(w)
    fd.add_value_constraint(p2, 7.0, 1.0)
    fd.add_value_constraint(p0, 2.0, 1.0)
    fd.add_value_constraint([1.5, 11.25], -3.0, 1.0)
    fd.add_value_constraint([11.0, 1.75], 0.25, 1.0)
    fd.add_value_constraint([6.0, 7.0], 1.0, 1.0)
    d = fd.system()
    assert list(got["d_counts"]) == [d.num_rows, d.num_triplets]
    assert rel(got["d_exact"], O.exact_solve(d, n * n)) <= 1e-5
    fe = base()
    fe.add_value_constraint(p2, 5.0, 2.0)
    assert rel(got["e_exact"], O.exact_solve(fe.system(), n * n)) <= 1e-5  # b without its hand-written row


@pytest.mark.gpu
def test_deferred_triplets_3d(tmp_path, port):
    got = _run("deferred_3d", tmp_path)
    n = 20
    want = port.sdf_from_points([n, n, n], O.make_weights(), got["points"].reshape(-1, 3), got["normals"].reshape(-1, 3)).system()
    assert list(got["counts"]) == [want.num_rows, want.num_triplets, 0]
    assert np.array_equal(got["rows"], want.rows) and np.array_equal(got["cols"], want.cols)
    assert np.array_equal(got["vals"].view(np.uint32), want.vals.view(np.uint32))
    assert np.array_equal(got["rhs"].view(np.uint32), want.rhs.view(np.uint32))
    assert got["stats"][1] == 1
    assert rel(got["solution"], O.exact_solve(want, n ** 3)) <= 1e-5
