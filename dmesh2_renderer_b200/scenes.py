"""Pinned synthetic inputs for the BASELINE.json configurations.

All random draws come from one ``numpy.random.default_rng(seed)`` per scene, in a fixed order
(centres, offsets, colours, opacity, intense, Wc, Wd), cast to fp32 on the host.  The camera is an
OpenGL-style perspective projection with fovy 45 deg, aspect 1, near 1.0 (the reference's ray set-up
is only exact for near = 1, reference dmesh2_renderer/__init__.py:225-231), far 10, looking at the
origin from distance 3.
"""
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np


def perspective(fovy_deg=45.0, aspect=1.0, near=1.0, far=10.0):
    t = math.tan(math.radians(fovy_deg) / 2.0)
    m = np.zeros((4, 4), dtype=np.float64)
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = 1.0 / t
    m[2, 2] = -(far + near) / (far - near)
    m[2, 3] = -2.0 * far * near / (far - near)
    m[3, 2] = -1.0
    return m.astype(np.float32)


def look_at(eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)):
    eye = np.asarray(eye, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - eye
    f /= np.linalg.norm(f)
    upv = np.asarray(up, dtype=np.float64)
    if abs(np.dot(f, upv)) > 0.999:          # looking straight along the up axis
        upv = np.array([0.0, 0.0, 1.0])
    s = np.cross(f, upv)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float64)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[:3, 3] = -m[:3, :3] @ eye
    return m.astype(np.float32)


def fibonacci_eyes(n=64, radius=3.0):
    i = np.arange(n, dtype=np.float64)
    y = 1.0 - 2.0 * (i + 0.5) / n
    r = np.sqrt(1.0 - y * y)
    th = i * math.pi * (3.0 - math.sqrt(5.0))
    return radius * np.stack((r * np.cos(th), y, r * np.sin(th)), axis=1)


@dataclass
class SoupScene:
    name: str
    width: int
    height: int
    mv: np.ndarray            # [V, 4, 4]
    proj: np.ndarray          # [V, 4, 4]
    verts: np.ndarray         # [P, 3]
    faces: np.ndarray         # [F, 3] int32
    verts_color: np.ndarray   # [P, 3]
    faces_opacity: np.ndarray  # [F]
    faces_intense: np.ndarray  # [V, F]
    background: np.ndarray    # [3]
    w_color: np.ndarray       # [V, H, W, 3] loss weights
    w_depth: np.ndarray       # [V, H, W]
    aa_temperature: float = 1.0
    extra: dict = field(default_factory=dict)


def triangle_soup(name, n_faces, size, sigma, seed, eyes=((0.0, 0.0, 3.0),), xy_squeeze=1.0,
                  weights_seed: Optional[int] = None):
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-1.0, 1.0, size=(n_faces, 3))
    if xy_squeeze != 1.0:
        centres[:, :2] *= xy_squeeze
    offsets = sigma * rng.standard_normal(size=(n_faces, 3, 3))
    verts = (centres[:, None, :] + offsets).reshape(-1, 3).astype(np.float32)
    faces = np.arange(3 * n_faces, dtype=np.int32).reshape(n_faces, 3)
    colour = rng.uniform(0.0, 1.0, size=(3 * n_faces, 3)).astype(np.float32)
    opacity = rng.uniform(0.05, 0.95, size=(n_faces,)).astype(np.float32)
    n_view = len(eyes)
    intense = rng.uniform(0.5, 1.0, size=(n_view, n_faces)).astype(np.float32)
    wrng = rng if weights_seed is None else np.random.default_rng(weights_seed)
    w_color = wrng.uniform(-1.0, 1.0, size=(n_view, size, size, 3)).astype(np.float32)
    w_depth = wrng.uniform(-1.0, 1.0, size=(n_view, size, size)).astype(np.float32)
    mv = np.stack([look_at(e) for e in eyes]).astype(np.float32)
    proj = np.stack([perspective() for _ in eyes]).astype(np.float32)
    return SoupScene(name, size, size, mv, proj, verts, faces, colour, opacity, intense,
                     np.ones(3, dtype=np.float32), w_color, w_depth)


def cfg1():
    """10k semi-transparent triangles, 256x256, one view."""
    return triangle_soup("cfg1_10k_256", 10_000, 256, 0.09, seed=1)


def cfg2(n_faces=1_000_000, size=1024):
    """The headline configuration: 1M triangles at 1024x1024, single view."""
    return triangle_soup("cfg2_1M_1024", n_faces, size, 0.0115, seed=2)


def cfg4(n_views=64, size=512, n_faces=1_000_000):
    """cfg2's scene seen from 64 Fibonacci-sphere cameras at 512x512 (loss weights from seed 4)."""
    eyes = [tuple(e) for e in fibonacci_eyes(64)[:n_views]]
    return triangle_soup("cfg4_64v_512", n_faces, size, 0.0115, seed=2, eyes=eyes, weights_seed=4)


def cfg5(n_faces=10_000_000, size=2048):
    """Stress: 10M triangles at 2048x2048, centre x,y squeezed to [-0.5, 0.5]."""
    return triangle_soup("cfg5_10M_2048", n_faces, size, 0.008, seed=5, xy_squeeze=0.5)


def small(n_faces=600, size=64, sigma=0.12, seed=7, n_views=1):
    """A seconds-scale scene for oracle parity tests."""
    eyes = [tuple(e) for e in fibonacci_eyes(64)[:n_views]] if n_views > 1 else ((0.0, 0.0, 3.0),)
    return triangle_soup(f"small_{n_faces}_{size}", n_faces, size, sigma, seed=seed, eyes=eyes)


@dataclass
class TetScene:
    name: str
    width: int
    height: int
    mv: np.ndarray
    proj: np.ndarray
    verts: np.ndarray          # [P, 3]
    faces: np.ndarray          # [F, 3] int32 (sorted vertex triples, unique)
    tets: np.ndarray           # [T, 4] int32
    face_tets: np.ndarray      # [F, 2] int32, -1 = none
    tet_faces: np.ndarray      # [T, 4] int32
    face_existence: np.ndarray  # [F] int32
    num_layers: int = 8


def delaunay_tets(name, n_points, size, seed, p_exist=0.15, num_layers=8, eyes=((0.0, 0.0, 3.0),), points=None,
                  qhull_options=None):
    """points: use these (n, 3) points instead of n_points uniform ones (e.g. a lattice: cospherical points, flat
    slivers, rays through vertices and along faces); qhull_options: e.g. "QJ" for such degenerate input."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1.0, 1.0, size=(n_points, 3)) if points is None else np.asarray(points, dtype=np.float64)
    tri = Delaunay(pts) if qhull_options is None else Delaunay(pts, qhull_options=qhull_options)
    tets = tri.simplices.astype(np.int32)                                      # [T, 4]
    combos = np.array([[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]])
    all_faces = np.sort(tets[:, combos].reshape(-1, 3), axis=1)                # [4T, 3]
    faces, inverse = np.unique(all_faces, axis=0, return_inverse=True)
    inverse = inverse.reshape(-1)
    tet_faces = inverse.reshape(-1, 4).astype(np.int32)
    n_faces = faces.shape[0]
    face_tets = -np.ones((n_faces, 2), dtype=np.int32)
    owner = np.repeat(np.arange(tets.shape[0], dtype=np.int32), 4)
    order = np.argsort(inverse, kind="stable")
    fs, ts = inverse[order], owner[order]
    first = np.ones_like(fs, dtype=bool)
    first[1:] = fs[1:] != fs[:-1]
    face_tets[fs[first], 0] = ts[first]
    face_tets[fs[~first], 1] = ts[~first]
    existence = (rng.uniform(size=(n_faces,)) < p_exist).astype(np.int32)
    mv = np.stack([look_at(e) for e in eyes]).astype(np.float32)
    proj = np.stack([perspective() for _ in eyes]).astype(np.float32)
    return TetScene(name, size, size, mv, proj, pts.astype(np.float32), faces.astype(np.int32), tets, face_tets,
                    tet_faces, existence, num_layers)


def cfg3(n_points=37_300, size=1024):
    """LayeredRenderer: Delaunay tet mesh of 37 300 points, 8 layers, 1024x1024."""
    return delaunay_tets("cfg3_layers_1024", n_points, size, seed=3)


def small_tets(n_points=300, size=64, seed=11, num_layers=4):
    return delaunay_tets(f"small_tets_{n_points}_{size}", n_points, size, seed=seed, num_layers=num_layers)
