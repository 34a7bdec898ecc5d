"""TEST INFRASTRUCTURE — ctypes loaders for the two CPU checkers (never imported by the product).

* ``port()``      -> oracle/libfi_oracle.so, the in-repo CPU restatement (oracle/fi_oracle.cpp)
* ``reference()`` -> oracle/_ref/libfi_ref.so, the reference's own assembly TU compiled unmodified
                     (oracle/Makefile); ``None`` when it has not been built.

Both export the same assembly API under a different prefix (``ora_`` / ``ref_``), so tests run the same
code against either.  The solve half (normal equations, BiCGSTAB, PCG, Jacobi) exists only in the port,
because the reference's sparse_linear.cpp cannot be built without Eigen (see fi_oracle.cpp header).
``exact_solve`` is the fp64 sparse direct solve of the normal equations with scipy — the stand-in for
``solve_sparse_linear_exact`` (reference sparse_linear.cpp:154-184: SimplicialLLT<double> on AᵀA).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

VALUE_NEAREST, VALUE_LINEAR = 0, 1                      # ValueKernel, field_interpolation.hpp:47-51
GRAD_NEAREST, GRAD_CELL_EDGES, GRAD_LINEAR = 0, 1, 2    # GradientKernel, field_interpolation.hpp:54-59


class CWeights(C.Structure):
    """Field order of Weights, field_interpolation.hpp:75-95."""
    _fields_ = [(n, C.c_float) for n in (
        "data_pos", "data_gradient", "model_0", "model_1", "model_2", "model_3", "model_4",
        "gradient_smoothness")] + [("value_kernel", C.c_int), ("gradient_kernel", C.c_int)]


def make_weights(**kw) -> CWeights:
    w = CWeights(1.0, 1.0, 0.0, 0.0, 0.5, 0.0, 0.0, 0.0, VALUE_LINEAR, GRAD_CELL_EDGES)
    for k, v in kw.items():
        if not hasattr(w, k):
            raise KeyError(k)
        setattr(w, k, v)
    return w


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ty=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


@dataclass
class System:
    rows: np.ndarray
    cols: np.ndarray
    vals: np.ndarray
    rhs: np.ndarray

    @property
    def num_rows(self):
        return int(self.rhs.shape[0])

    @property
    def num_triplets(self):
        return int(self.vals.shape[0])


class Field:
    """One LatticeField living inside a checker library."""

    def __init__(self, lib: "CpuLib", handle, sizes):
        self.lib, self.h, self.sizes = lib, handle, [int(s) for s in sizes]

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.fn("field_destroy")(self.h)
            self.h = None

    def add_field_constraints(self, w: CWeights):
        self.lib.fn("add_field_constraints")(self.h, C.byref(w))

    def add_value_constraint(self, pos, value, weight) -> bool:
        p = _f32(pos)
        return bool(self.lib.fn("add_value_constraint")(self.h, _ptr(p), C.c_float(value), C.c_float(weight)))

    def add_value_constraint_nearest_neighbor(self, pos, gradient, value, weight) -> bool:
        p, g = _f32(pos), _f32(gradient)
        return bool(self.lib.fn("add_value_constraint_nearest_neighbor")(
            self.h, _ptr(p), _ptr(g), C.c_float(value), C.c_float(weight)))

    def add_gradient_constraint(self, pos, gradient, weight, kernel) -> bool:
        p, g = _f32(pos), _f32(gradient)
        return bool(self.lib.fn("add_gradient_constraint")(self.h, _ptr(p), _ptr(g), C.c_float(weight), int(kernel)))

    def add_points(self, value_weight, value_kernel, gradient_weight, gradient_kernel, positions, normals=None,
                   point_weights=None):
        p, n, pw = _f32(positions), _f32(normals), _f32(point_weights)
        npts = 0 if p.size == 0 else p.size // len(self.sizes)
        self.lib.fn("add_points")(self.h, C.c_float(value_weight), int(value_kernel), C.c_float(gradient_weight),
                                  int(gradient_kernel), int(npts), _ptr(p), _ptr(n), _ptr(pw))

    def add_equation(self, weight, rhs, columns, values):
        cols = np.ascontiguousarray(columns, dtype=np.int32)
        vals = _f32(values)
        self.lib.fn("add_equation")(self.h, C.c_float(weight), C.c_float(rhs), int(cols.size), _ptr(cols, C.c_int),
                                    _ptr(vals))

    def system(self) -> System:
        nr = int(self.lib.fn("num_rows")(self.h))
        nt = int(self.lib.fn("num_triplets")(self.h))
        rows, cols = np.empty(nt, np.int32), np.empty(nt, np.int32)
        vals, rhs = np.empty(nt, np.float32), np.empty(nr, np.float32)
        self.lib.fn("copy_system")(self.h, _ptr(rows, C.c_int), _ptr(cols, C.c_int), _ptr(vals), _ptr(rhs))
        return System(rows, cols, vals, rhs)


class CpuLib:
    def __init__(self, path: str, prefix: str):
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        self._sig()

    def fn(self, name):
        return getattr(self.dll, self.prefix + name)

    def _sig(self):
        f, i, i64, vp = C.c_float, C.c_int, C.c_int64, C.c_void_p
        pf, pi = C.POINTER(C.c_float), C.POINTER(C.c_int)
        pw = C.POINTER(CWeights)
        S = {
            "field_create": (vp, [i, pi]), "field_destroy": (None, [vp]),
            "add_field_constraints": (None, [vp, pw]),
            "add_value_constraint": (i, [vp, pf, f, f]),
            "add_value_constraint_nearest_neighbor": (i, [vp, pf, pf, f, f]),
            "add_gradient_constraint": (i, [vp, pf, pf, f, i]),
            "add_points": (None, [vp, f, i, f, i, i, pf, pf, pf]),
            "add_equation": (None, [vp, f, f, i, pi, pf]),
            "sdf_from_points": (vp, [i, pi, pw, i, pf, pf, pf]),
            "num_rows": (i64, [vp]), "num_triplets": (i64, [vp]),
            "copy_system": (None, [vp, pi, pi, pf, pf]),
            "upscale_field": (None, [pf, i, pi, pi, pf]),
            "generate_error_map": (None, [i64, pi, pi, pf, i64, pf, i64, pf, pf]),
            "marching_squares": (i64, [i64, i64, pf, pf]), "calc_area": (f, [i64, pf]),
            "bicubic_upsample": (None, [i, i, pf, i, pf]), "iso_surface": (i64, [i, i, pf, f, pf]),
        }
        for name, (res, args) in S.items():
            fn = self.fn(name)
            fn.restype, fn.argtypes = res, args
        if self.prefix == "ora_":
            pd, pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int64)
            T = {
                "normal_create": (vp, [i64, pi, pi, pf, i64, pf, i64, i]),
                "normal_destroy": (None, [vp, i]), "normal_nnz": (i64, [vp, i]),
                "normal_copy": (None, [vp, i, pi64, pi, pd, pd]),
                "bicgstab_f32": (None, [vp, pf, i64, f, pi64, pf]),
                "bicgstab_f64": (None, [vp, pd, i64, C.c_double, pi64, pd]),
                "pcg_f32": (None, [vp, pf, i64, f, pi64, pf]),
                "pcg_f64": (None, [vp, pd, i64, C.c_double, pi64, pd]),
                "jacobi_f32": (None, [vp, pf, i, f]),
                "tile_solve_f32": (C.c_int, [vp, pf, i, C.POINTER(C.c_int), i, pf]),
                "tile_solve_f64": (C.c_int, [vp, pd, i, C.POINTER(C.c_int), i, pd]),
                "apply_f64": (None, [vp, pd, pd]), "apply_f32": (None, [vp, pf, pf]),
            }
            for name, (res, args) in T.items():
                fn = self.fn(name)
                fn.restype, fn.argtypes = res, args

    # ---- assembly -------------------------------------------------------------------------
    def field(self, sizes) -> Field:
        sz = np.ascontiguousarray(sizes, dtype=np.int32)
        return Field(self, self.fn("field_create")(len(sz), _ptr(sz, C.c_int)), sz)

    def sdf_from_points(self, sizes, w: CWeights, positions, normals=None, point_weights=None) -> Field:
        sz = np.ascontiguousarray(sizes, dtype=np.int32)
        p, n, pw = _f32(positions), _f32(normals), _f32(point_weights)
        npts = 0 if p is None or p.size == 0 else p.size // len(sz)
        h = self.fn("sdf_from_points")(len(sz), _ptr(sz, C.c_int), C.byref(w), int(npts), _ptr(p), _ptr(n), _ptr(pw))
        return Field(self, h, sz)

    def upscale_field(self, small, small_sizes, large_sizes) -> np.ndarray:
        ss = np.ascontiguousarray(small_sizes, dtype=np.int32)
        ls = np.ascontiguousarray(large_sizes, dtype=np.int32)
        src = _f32(small).ravel()
        out = np.empty(int(np.prod(ls, dtype=np.int64)), np.float32)
        self.fn("upscale_field")(_ptr(src), len(ss), _ptr(ss, C.c_int), _ptr(ls, C.c_int), _ptr(out))
        return out

    def generate_error_map(self, sys: System, solution) -> np.ndarray:
        sol = _f32(solution)
        out = np.empty(sol.size, np.float32)
        self.fn("generate_error_map")(sys.num_triplets, _ptr(sys.rows, C.c_int), _ptr(sys.cols, C.c_int),
                                      _ptr(sys.vals), sol.size, _ptr(sol), sys.num_rows, _ptr(sys.rhs), _ptr(out))
        return out

    # ---- iso-surface helpers (emilib/marching_squares.cpp, src/sdf_field.cpp:555-614) ----------------
    def marching_squares(self, iso) -> np.ndarray:
        """`iso`: (height, width) float32, row-major.  Returns (num_segments, 4) float32: x0 y0 x1 y1."""
        a = np.ascontiguousarray(iso, dtype=np.float32)
        h, w = a.shape
        n = int(self.fn("marching_squares")(w, h, _ptr(a), None))
        out = np.empty(max(n, 1), np.float32)
        self.fn("marching_squares")(w, h, _ptr(a), _ptr(out))
        return out[:n].reshape(-1, 4)

    def iso_surface(self, values, iso: float) -> np.ndarray:
        a = np.ascontiguousarray(values, dtype=np.float32)
        h, w = a.shape
        n = int(self.fn("iso_surface")(w, h, _ptr(a), float(iso), None))
        out = np.empty(max(n, 1), np.float32)
        self.fn("iso_surface")(w, h, _ptr(a), float(iso), _ptr(out))
        return out[:n].reshape(-1, 4)

    def calc_area(self, lines) -> float:
        a = np.ascontiguousarray(lines, dtype=np.float32).reshape(-1, 4)
        return float(self.fn("calc_area")(a.shape[0], _ptr(a) if a.size else None))

    def bicubic_upsample(self, values, upsample: int) -> np.ndarray:
        a = np.ascontiguousarray(values, dtype=np.float32)
        h, w = a.shape
        out = np.empty((upsample * h - upsample + 1, upsample * w - upsample + 1), np.float32)
        self.fn("bicubic_upsample")(w, h, _ptr(a), int(upsample), _ptr(out))
        return out

    # ---- solve half (port only) -----------------------------------------------------------------
    def normal(self, sys: System, num_columns: int, precision: str = "f32") -> "Normal":
        assert self.prefix == "ora_", "the reference's solve half cannot be built here (needs Eigen)"
        return Normal(self, sys, int(num_columns), precision)


class Normal:
    """AᵀA / Aᵀb of a triplet system in the port (sparse_linear.cpp:59-113)."""

    def __init__(self, lib: CpuLib, sys: System, ncols: int, precision: str):
        self.lib, self.n, self.prec = lib, ncols, {"f32": 0, "f64": 1}[precision]
        self.h = lib.fn("normal_create")(sys.num_triplets, _ptr(sys.rows, C.c_int), _ptr(sys.cols, C.c_int),
                                         _ptr(sys.vals), sys.num_rows, _ptr(sys.rhs), ncols, self.prec)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.fn("normal_destroy")(self.h, self.prec)
            self.h = None

    def csr(self):
        import scipy.sparse as sp
        nnz = int(self.lib.fn("normal_nnz")(self.h, self.prec))
        indptr, indices = np.empty(self.n + 1, np.int64), np.empty(nnz, np.int32)
        data, atb = np.empty(nnz, np.float64), np.empty(self.n, np.float64)
        self.lib.fn("normal_copy")(self.h, self.prec, _ptr(indptr, C.c_int64), _ptr(indices, C.c_int),
                                   _ptr(data, C.c_double), _ptr(atb, C.c_double))
        return sp.csr_matrix((data, indices, indptr), shape=(self.n, self.n)), atb

    def _run(self, name, guess, max_iter, tol):
        dt, ct = (np.float32, C.c_float) if self.prec == 0 else (np.float64, C.c_double)
        x = np.zeros(self.n, dt) if guess is None else np.array(guess, dtype=dt, copy=True)
        iters, err = C.c_int64(0), ct(0)
        suffix = "_f32" if self.prec == 0 else "_f64"
        self.lib.fn(name + suffix)(self.h, _ptr(x, ct), int(max_iter), ct(tol), C.byref(iters), C.byref(err))
        return x, int(iters.value), float(err.value)

    def bicgstab(self, guess=None, max_iter=0, tol=0.0):
        return self._run("bicgstab", guess, max_iter, tol)

    def pcg(self, guess=None, max_iter=0, tol=1e-6):
        return self._run("pcg", guess, max_iter, tol)

    def jacobi(self, guess, iterations, weight):
        assert self.prec == 0
        x = np.array(guess, dtype=np.float32, copy=True)
        self.lib.fn("jacobi_f32")(self.h, _ptr(x), int(iterations), C.c_float(weight))
        return x

    def tile_solve(self, guess, sizes, tile_size):
        """tile_solver_square, sparse_linear.cpp:246-390.  Returns (solution, failed tiles)."""
        dt, ct, nm = (np.float32, C.c_float, "tile_solve_f32") if self.prec == 0 else (np.float64, C.c_double, "tile_solve_f64")
        g = np.ascontiguousarray(guess, dtype=dt).ravel()
        sz = np.ascontiguousarray(sizes, dtype=np.int32)
        out = np.empty_like(g)
        fails = self.lib.fn(nm)(self.h, _ptr(g, ct), len(sz), _ptr(sz, C.c_int), int(tile_size), _ptr(out, ct))
        return out, int(fails)

    def apply(self, x):
        dt, ct, nm = (np.float32, C.c_float, "apply_f32") if self.prec == 0 else (np.float64, C.c_double, "apply_f64")
        xi = np.ascontiguousarray(x, dtype=dt)
        y = np.empty_like(xi)
        self.lib.fn(nm)(self.h, _ptr(xi, ct), _ptr(y, ct))
        return y


def exact_solve(sys: System, num_columns: int) -> np.ndarray:
    """fp64 direct solve of AᵀA x = Aᵀb (explicit zeros dropped, as sparse_linear.cpp:84-86), cast to fp32."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    keep = sys.vals != 0
    A = sp.csr_matrix((sys.vals[keep].astype(np.float64), (sys.rows[keep], sys.cols[keep])),
                      shape=(sys.num_rows, num_columns))
    AtA = (A.T @ A).tocsc()
    Atb = A.T @ sys.rhs.astype(np.float64)
    return spla.spsolve(AtA, Atb)


def normal_equations_f64(sys: System, num_columns: int):
    import scipy.sparse as sp
    keep = sys.vals != 0
    A = sp.csr_matrix((sys.vals[keep].astype(np.float64), (sys.rows[keep], sys.cols[keep])),
                      shape=(sys.num_rows, num_columns))
    return (A.T @ A).tocsr(), A.T @ sys.rhs.astype(np.float64)


_cache = {}


def build(quiet=True):
    """Compile the checkers (building the checker is not using it)."""
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


def port() -> CpuLib:
    if "port" not in _cache:
        path = os.path.join(HERE, "libfi_oracle.so")
        if not os.path.exists(path):
            build()
        _cache["port"] = CpuLib(path, "ora_")
    return _cache["port"]


def reference():
    if "ref" not in _cache:
        path = os.path.join(HERE, "_ref", "libfi_ref.so")
        _cache["ref"] = CpuLib(path, "ref_") if os.path.exists(path) else None
    return _cache["ref"]
