"""Analytic synthetic scenes rendered to TUM-style u16 depth (5000 units per metre).

No dataset is available offline and the reference's own inputs (assets/T0.png, T1.png,
Application.cpp:28-29) are git-ignored, so every input of the tests and the bench is rendered
from closed-form geometry along a known trajectory.  Pure numpy, fp64,
deterministic; this is input generation, not part of the measured path.

Conventions: camera frame x right, y down, z forward; pose = camera->world 4x4 (row-major);
a pixel (u, v) looks along K^-1 (u, v, 1), so the ray parameter IS the camera-frame depth z.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Scene:
    planes_z: list = field(default_factory=list)    # world planes z = const
    planes: list = field(default_factory=list)      # general planes (nx, ny, nz, d): n.p = d
    spheres: list = field(default_factory=list)     # (cx, cy, cz, r)
    boxes: list = field(default_factory=list)       # inward-facing (xmin, xmax, ymin, ymax, zmin, zmax): camera inside


def scene_S1() -> Scene:
    """Plane z = 2.5 m plus a sphere at (0, 0, 2.0) of radius 0.5 (config C1 / C2)."""
    return Scene(planes_z=[2.5], spheres=[(0.0, 0.0, 2.0, 0.5)])


def scene_S1T() -> Scene:
    """Tracking variant of S1: S1 is rotationally symmetric about the optical axis (rotation about z and
    in-plane translation are weakly observable), so a floor (y = 1.0), a left wall (x = -1.5) and an
    off-axis sphere are added; three orthogonal planes constrain all six degrees of freedom (config C2/C5)."""
    return Scene(planes_z=[2.5], planes=[(0.0, 1.0, 0.0, 1.0), (1.0, 0.0, 0.0, -1.5)],
                 spheres=[(0.0, 0.0, 2.0, 0.5), (0.9, 0.4, 1.8, 0.3)])


def scene_S2() -> Scene:
    """'Room': inward-facing box [-3,3]x[-1.5,1.5]x[0,6] (camera at z ~ 0.3) plus three spheres (config C3)."""
    return Scene(boxes=[(-3.0, 3.0, -1.5, 1.5, -0.5, 6.0)],
                 spheres=[(-1.2, 0.6, 3.0, 0.6), (1.0, 0.2, 4.0, 0.8), (0.1, -0.6, 2.2, 0.35)])


def scene_S3() -> Scene:
    """'Hall': box [-10,10]x[-3,3]x[-1,12] plus eight spheres 4-12 m away (config C4)."""
    sph = [(-6.0, 1.5, 8.0, 1.2), (-3.0, -0.5, 6.0, 0.9), (0.0, 1.0, 9.0, 1.5), (3.0, 0.0, 7.0, 1.0),
           (6.0, 1.8, 10.0, 1.1), (-1.5, 2.0, 5.0, 0.7), (1.8, -1.2, 5.5, 0.6), (4.5, 1.0, 11.0, 0.9)]
    return Scene(boxes=[(-10.0, 10.0, -3.0, 3.0, -1.0, 12.0)], spheres=sph)


def rot_y(deg: float) -> np.ndarray:
    a = np.deg2rad(deg)
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float64)


def trans(x: float, y: float, z: float) -> np.ndarray:
    m = np.eye(4)
    m[:3, 3] = (x, y, z)
    return m


def trajectory_C2(k: int) -> np.ndarray:
    """T_k = Trans(0.002 k, 0, 0) . RotY(0.05 deg . k)   (SURVEY.md section 8d, config C2)."""
    return trans(0.002 * k, 0.0, 0.0) @ rot_y(0.05 * k)


def trajectory_C3(k: int) -> np.ndarray:
    """Slow orbit: Trans(0.001 k, 0, 0) . RotY(0.03 deg . k)   (config C3 / C4)."""
    return trans(0.001 * k, 0.0, 0.0) @ rot_y(0.03 * k)


def pingpong(k: int, n: int) -> int:
    """0..n-1, n-2..1, 0.. : keeps the motion continuous when a bench runs past the sequence length."""
    if n <= 1:
        return 0
    period = 2 * (n - 1)
    k %= period
    return k if k < n else period - k


def render_depth(scene: Scene, pose: np.ndarray, width: int, height: int, fx: float, fy: float, cx: float, cy: float,
                 depth_scale: float = 5000.0, max_depth: float = 13.0) -> np.ndarray:
    """u16 depth image: d16 = round(depth_scale * z) of the first hit along each pixel ray, 0 = miss."""
    pose = np.asarray(pose, dtype=np.float64).reshape(4, 4)
    R, t = pose[:3, :3], pose[:3, 3]
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    dc = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1)      # camera ray, z = 1
    d = dc @ R.T                                                                  # world direction
    o = t
    best = np.full((height, width), np.inf)

    with np.errstate(divide="ignore", invalid="ignore"):
        for zp in scene.planes_z:
            s = (zp - o[2]) / d[..., 2]
            s = np.where(s > 1e-6, s, np.inf)
            best = np.minimum(best, s)
        for (nx, ny, nz, dd) in scene.planes:
            nvec = np.array([nx, ny, nz], dtype=np.float64)
            s = (dd - float(nvec @ o)) / (d @ nvec)
            s = np.where(np.isfinite(s) & (s > 1e-6), s, np.inf)
            best = np.minimum(best, s)
        for (sx, sy, sz, r) in scene.spheres:
            oc = o - np.array([sx, sy, sz])
            a = np.sum(d * d, axis=-1)
            b = 2.0 * (d @ oc)
            c = float(oc @ oc) - r * r
            disc = b * b - 4 * a * c
            sq = np.sqrt(np.where(disc >= 0, disc, np.nan))
            s0 = (-b - sq) / (2 * a)
            s1 = (-b + sq) / (2 * a)
            s = np.where(s0 > 1e-6, s0, np.where(s1 > 1e-6, s1, np.inf))
            s = np.where(np.isnan(s), np.inf, s)
            best = np.minimum(best, s)
        for (x0, x1, y0, y1, z0, z1) in scene.boxes:
            lo = np.array([x0, y0, z0])
            hi = np.array([x1, y1, z1])
            wall = np.where(d > 0, hi, lo)
            s = (wall - o) / d
            s = np.where(np.isfinite(s) & (s > 1e-6), s, np.inf)
            best = np.minimum(best, np.min(s, axis=-1))

    z = np.where(np.isfinite(best) & (best < max_depth), best, 0.0)
    d16 = np.rint(z * depth_scale)
    d16 = np.where(d16 > 65535, 0, d16)
    return d16.astype(np.uint16)


def render_sequence(scene: Scene, poses, width, height, fx, fy, cx, cy, depth_scale=5000.0) -> np.ndarray:
    return np.stack([render_depth(scene, p, width, height, fx, fy, cx, cy, depth_scale) for p in poses])
