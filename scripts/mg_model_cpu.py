"""Exploration (not part of the product; CPU only): a scipy model of the multigrid-preconditioned CG of csrc/mg.cu,
used to choose hierarchy parameters without spending GPU time.  Mirrors build_multigrid: re-discretised coarse
operators (same points in coarse coordinates, weights rescaled), align-corners multilinear transfers, Chebyshev
smoothing, dense coarsest solve.  Prints PCG iterations to 1e-6 for a few variants.

    python scripts/mg_model_cpu.py [n=48] [npts=9000] [dim=3]
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, ".")
from field_interpolation_b200 import workloads as W
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 3
port = O.port()
cloud = W.sphere_torus_3d(npts, seed=0) if D == 3 else W.circles_2d(npts, seed=0)


def level_sizes(n0, coarsest_cells, min_next=4):
    out = [[n0] * D]
    while True:
        cur = out[-1]
        nxt = [(s + 1) // 2 for s in cur]
        if np.prod(cur) <= coarsest_cells or min(nxt) < min_next:
            break
        out.append(nxt)
    return out


def operator(sizes, l, fine_sizes, boost=1.0):
    k = np.arange(5)
    scale = 2.0 ** ((D - 2 * k) * l / 2.0)  # squared weights scale by 2^((D-2k) l)
    w = O.make_weights(model_2=0.5 * scale[2] * boost ** (l / 2.0), data_gradient=0.5 ** l)
    sc = (np.asarray(sizes, np.float64) - 1) / (np.asarray(fine_sizes, np.float64) - 1)
    pos = (W.to_lattice(cloud["unit_pos"], fine_sizes).astype(np.float64) * sc[None, :]).astype(np.float32)
    f = port.sdf_from_points(sizes, w, pos, cloud["normals"])
    M, atb = O.normal_equations_f64(f.system(), int(np.prod(sizes)))
    return M.tocsr(), atb


def interp_1d(nf, nc, cubic=False):
    rows, cols, vals = [], [], []
    sc = (nc - 1) / (nf - 1) if nf > 1 else 0.0
    for i in range(nf):
        t = i * sc
        b = min(int(t), nc - 1)
        fr = t - b
        if not cubic:
            taps = [(b, 1 - fr), (min(b + 1, nc - 1), fr)]
        else:  # Catmull-Rom / cubic Lagrange through b-1..b+2, falling back to linear at the ends
            if b - 1 < 0 or b + 2 > nc - 1:
                taps = [(b, 1 - fr), (min(b + 1, nc - 1), fr)]
            else:
                x = fr
                taps = [(b - 1, -x * (x - 1) * (x - 2) / 6), (b, (x + 1) * (x - 1) * (x - 2) / 2), (b + 1, -(x + 1) * x * (x - 2) / 2),
                        (b + 2, (x + 1) * x * (x - 1) / 6)]
        for c, v in taps:
            if v != 0:
                rows.append(i), cols.append(c), vals.append(v)
    return sp.csr_matrix((vals, (rows, cols)), shape=(nf, nc))


def prolongation(fs, cs, cubic=False):
    P = None
    for d in range(D):  # axis 0 fastest: kron(slowest, ..., fastest)
        Pd = interp_1d(fs[d], cs[d], cubic)
        P = Pd if P is None else sp.kron(Pd, P, format="csr")
    return P.tocsr()


class MG:
    def __init__(self, sizes_list, nu=3, ratio=12.0, cubic=False, galerkin=False, restrict_cubic=None, cycle="V", boost=1.0):
        self.nu, self.ratio, self.cycle = nu, ratio, cycle
        self.A, self.P, self.R, self.minv, self.lmax = [], [], [], [], []
        for l, sz in enumerate(sizes_list):
            if l == 0 or not galerkin:
                A, _ = operator(sz, l, sizes_list[0], boost)
            else:
                A = (self.R[-1] @ self.A[-1] @ self.P[-1]).tocsr()
            self.A.append(A)
            self.minv.append(1.0 / A.diagonal())
            if l + 1 < len(sizes_list):
                P = prolongation(sz, sizes_list[l + 1], cubic)
                self.P.append(P)
                rc = cubic if restrict_cubic is None else restrict_cubic
                self.R.append(P.T.tocsr() if rc == cubic else prolongation(sz, sizes_list[l + 1], rc).T.tocsr())
                v = np.random.default_rng(0).normal(size=A.shape[0])
                for _ in range(12):
                    v = self.minv[l] * (A @ v)
                    lam = np.linalg.norm(v)
                    v /= lam
                self.lmax.append(lam * 1.1)
        self.Ainv = np.linalg.inv(self.A[-1].toarray())

    def smooth(self, l, r, e):
        A, minv = self.A[l], self.minv[l]
        lmax = self.lmax[l]
        lmin = lmax / self.ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        res = r - A @ e if e is not None else r.copy()
        d = minv * res / theta
        e = d.copy() if e is None else e + d
        rho = 1.0 / sigma
        for _ in range(1, self.nu):
            rho_new = 1.0 / (2 * sigma - rho)
            res = res - A @ d
            d = rho_new * rho * d + 2 * rho_new / delta * minv * res
            e = e + d
            rho = rho_new
        return e

    def vcycle(self, l, r):
        if l == len(self.A) - 1:
            return self.Ainv @ r
        e = self.smooth(l, r, None)
        for _ in range(2 if (self.cycle == "W" and l > 0) else 1):
            res = r - self.A[l] @ e
            ec = self.vcycle(l + 1, self.R[l] @ res)
            e = e + self.P[l] @ ec
        return self.smooth(l, r, e)


def pcg(A, b, prec, tol=1e-6, maxit=300):
    x = np.zeros_like(b)
    r = b.copy()
    z = prec(r)
    p = z.copy()
    rz = r @ z
    bb = b @ b
    for it in range(1, maxit + 1):
        q = A @ p
        alpha = rz / (p @ q)
        x += alpha * p
        r -= alpha * q
        if r @ r <= tol * tol * bb:
            return it
        z = prec(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return maxit


if __name__ == "__main__":
    A0, b0 = operator([n] * D, 0, [n] * D)
    print(f"n={n} D={D} npts={npts} unknowns={A0.shape[0]} nnz={A0.nnz}")
    variants = []
    for cc in (150, 600):
        for b in (1.0, 1.5, 2.0, 2.5, 3.0):
            variants.append((f"linear V(3,3) coarsest<={cc} boost {b}", dict(coarsest=cc, boost=b)))
    variants += [("linear V(2,2) boost 2", dict(coarsest=600, nu=2, boost=2.0)), ("linear W(3,3) boost 2", dict(coarsest=600, cycle="W", boost=2.0)),
                 ("linear V(3,3) boost 2 ratio 6", dict(coarsest=600, boost=2.0, ratio=6.0)),
                 ("linear V(3,3) boost 2 ratio 20", dict(coarsest=600, boost=2.0, ratio=20.0))]
    for name, kw in variants:
        t0 = time.time()
        sl = level_sizes(n, kw.pop("coarsest"))
        mg = MG(sl, **kw)
        its = pcg(A0, b0, lambda r: mg.vcycle(0, r))
        print(f"{name:34s} levels {[s[0] for s in sl]}  iterations {its:4d}   ({time.time() - t0:.1f} s)", flush=True)
