"""Synthetic inputs for the BASELINE.json configurations.  numpy only, deterministic.

Point clouds are produced in *unit* coordinates ([0,1]^D) like the reference demo does
(reference src/sdf_field.cpp:198-210 scales them per lattice with ``pos * (resolution - 1)``); normals are
unit length and passed unscaled (src/sdf_field.cpp:215-216).
"""
from __future__ import annotations

import numpy as np

# 4x4 value table of the 2D interpolation demo (reference src/interpolate_2d.cpp:17-22)
TABLE_2D = np.array([[5, 4, 2, 3], [4, 2, 1, 5], [6, 3, 5, 2], [1, 2, 4, 1]], dtype=np.float64)


def to_lattice(unit_positions: np.ndarray, sizes) -> np.ndarray:
    """fp32 ``pos * (size - 1.0f)`` per axis, the reference's on_lattice (src/sdf_field.cpp:198-210)."""
    scale = (np.asarray(sizes, dtype=np.float32) - np.float32(1.0)).astype(np.float32)
    return (unit_positions.astype(np.float32) * scale[None, :]).astype(np.float32)


def field_1d(resolution: int = 100):
    """C1 — reference src/field_1d.cpp:20-29,98-107: two points with value and gradient, data rows first."""
    pts = np.array([[0.2, 0.0, +1.0], [0.8, 0.0, -1.0]], dtype=np.float32)
    pos = (pts[:, 0] * np.float32(resolution - 1)).astype(np.float32)
    grad = (pts[:, 2] / np.float32(resolution - 1)).astype(np.float32)
    return {"sizes": [resolution], "pos": pos.reshape(-1, 1), "value": pts[:, 1].copy(), "gradient": grad.reshape(-1, 1)}


def _catmull_rom(table, u, v):
    def w(t):
        return np.stack([-0.5 * t**3 + t**2 - 0.5 * t, 1.5 * t**3 - 2.5 * t**2 + 1.0,
                         -1.5 * t**3 + 2.0 * t**2 + 0.5 * t, 0.5 * t**3 - 0.5 * t**2], axis=-1)
    x, y = u * 3.0, v * 3.0
    ix, iy = np.clip(np.floor(x).astype(int), 0, 2), np.clip(np.floor(y).astype(int), 0, 2)
    wx, wy = w(x - ix), w(y - iy)
    out = np.zeros_like(u)
    for a in range(4):
        for b in range(4):
            out += wy[:, a] * wx[:, b] * table[np.clip(iy + a - 1, 0, 3), np.clip(ix + b - 1, 0, 3)]
    return out


def interpolate_2d(n: int = 512, num_points: int = 10_000, seed: int = 1):
    """C2 — noisy samples of a smooth surface through the demo's 4x4 table; values only (no gradients)."""
    rng = np.random.default_rng(seed)
    unit = rng.random((num_points, 2))
    val = _catmull_rom(TABLE_2D, unit[:, 0], unit[:, 1]) + rng.normal(0.0, 0.1, num_points)
    return {"sizes": [n, n], "unit_pos": unit.astype(np.float32), "value": val.astype(np.float32),
            "weights": dict(data_pos=1.0, data_gradient=0.0, model_1=0.1, model_2=1.0)}


def circles_2d(num_points: int = 200_000, seed: int = 0, pos_stddev=0.005, normal_stddev=0.05):
    """C3 — the SDF demo's two default shapes (circle r=0.35, inverted circle r=0.1; src/sdf_field.cpp:25-71)
    with its noise model (src/sdf_field.cpp:40-46,306-321): position noise, normal-angle noise."""
    rng = np.random.default_rng(seed)
    n_outer = num_points * 3 // 4
    n_inner = num_points - n_outer
    ang = np.concatenate([rng.random(n_outer), rng.random(n_inner)]) * 2.0 * np.pi
    rad = np.concatenate([np.full(n_outer, 0.35), np.full(n_inner, 0.1)])
    sign = np.concatenate([np.ones(n_outer), -np.ones(n_inner)])
    pos = 0.5 + rad[:, None] * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    pos += rng.normal(0.0, pos_stddev, pos.shape)
    nang = np.where(sign > 0, ang, ang + np.pi) + rng.normal(0.0, normal_stddev, ang.shape)
    nrm = np.stack([np.cos(nang), np.sin(nang)], axis=1)
    return {"unit_pos": pos.astype(np.float32), "normals": nrm.astype(np.float32)}


def sphere_torus_3d(num_points: int = 1_000_000, seed: int = 0, pos_stddev=0.005):
    """C4/C5 — half the samples on a sphere (c=0.5, r=0.3), half on a torus (R=0.25, r=0.10, axis z),
    analytic outward unit normals, isotropic position noise."""
    rng = np.random.default_rng(seed)
    ns = num_points // 2
    nt = num_points - ns
    d = rng.normal(size=(ns, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    sphere_p, sphere_n = 0.5 + 0.3 * d, d
    u, v = rng.random(nt) * 2 * np.pi, rng.random(nt) * 2 * np.pi
    R, r = 0.25, 0.10
    ring = np.stack([np.cos(u), np.sin(u), np.zeros(nt)], axis=1)
    tn = ring * np.cos(v)[:, None] + np.array([0.0, 0.0, 1.0])[None, :] * np.sin(v)[:, None]
    torus_p = 0.5 + R * ring + r * tn
    pos = np.concatenate([sphere_p, torus_p]) + rng.normal(0.0, pos_stddev, (num_points, 3))
    nrm = np.concatenate([sphere_n, tn])
    perm = rng.permutation(num_points)  # scanners do not deliver points sorted by shape
    return {"unit_pos": pos[perm].astype(np.float32), "normals": nrm[perm].astype(np.float32)}


def random_cloud(ndim: int, num_points: int, sizes, seed: int, margin: float = 1.5):
    """Test helper: points spread over and slightly beyond the lattice (exercises every skip rule), a share
    of them snapped to exact lattice coordinates (t == 0 hits) and cell centres."""
    rng = np.random.default_rng(seed)
    sz = np.asarray(sizes, dtype=np.float64)
    pos = rng.uniform(-margin, sz - 1 + margin, size=(num_points, ndim))
    snap = rng.random(num_points) < 0.15
    pos[snap] = np.round(pos[snap])
    half = rng.random(num_points) < 0.1
    pos[half] = np.floor(pos[half]) + 0.5
    nrm = rng.normal(size=(num_points, ndim))
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-9)
    return pos.astype(np.float32), nrm.astype(np.float32)
