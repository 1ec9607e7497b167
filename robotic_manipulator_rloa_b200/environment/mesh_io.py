"""
Mesh collision geometry -> convex vertex cloud.

PyBullet's URDF / SDF importer (reached through ``p.loadURDF`` / ``p.loadSDF``,
/root/reference/robotic_manipulator_rloa/environment/environment.py:228-233) turns a ``<mesh>`` collision
element into a ``btConvexHullShape`` over the mesh vertices with a 0.001 collision margin, and a
``<cylinder>`` into the hull of 2 x 32 rim points.  GJK over a vertex cloud equals GJK over
its hull, so only the vertices are kept; interior points are dropped with scipy's Qhull wrapper when it is
importable (an optimisation of the support loop, not a change of the shape).
"""
from __future__ import annotations

import logging
import os
import struct
from typing import List, Optional, Tuple

import numpy as np

HULL_MARGIN = 0.001            # gUrdfDefaultCollisionMargin of pybullet's importer
CYLINDER_STEPS = 32            # rim points per end cap of an imported <cylinder>
MAX_SHAPE_VERTS = 512          # per shape: hulls with more vertices are thinned to their support points (below)

logger = logging.getLogger(__name__)


class MeshError(Exception):
    """Raised for mesh files that cannot be read."""


def _load_stl(data: bytes) -> np.ndarray:
    # binary STL: 80-byte header, uint32 triangle count, 50 bytes per triangle.  An ASCII file starts with
    # "solid" too, so the size test decides.
    if len(data) >= 84:
        ntri = struct.unpack_from('<I', data, 80)[0]
        if 84 + 50 * ntri == len(data):
            rec = np.frombuffer(data, dtype=np.dtype([('n', '<f4', 3), ('v', '<f4', (3, 3)), ('a', '<u2')]),
                                count=ntri, offset=84)
            return rec['v'].reshape(-1, 3).astype(np.float64)
    verts: List[List[float]] = []
    for line in data.decode('ascii', errors='replace').splitlines():
        tok = line.split()
        if len(tok) == 4 and tok[0] == 'vertex':
            verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
    if not verts:
        raise MeshError('no vertices found in STL data')
    return np.asarray(verts, np.float64)


def _load_obj(data: bytes) -> np.ndarray:
    verts: List[List[float]] = []
    for line in data.decode('utf-8', errors='replace').splitlines():
        tok = line.split()
        if len(tok) >= 4 and tok[0] == 'v':
            verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
    if not verts:
        raise MeshError('no vertices found in OBJ data')
    return np.asarray(verts, np.float64)


def load_mesh_vertices(path: str) -> np.ndarray:
    """Vertices [n][3] of a binary / ASCII STL or a Wavefront OBJ file."""
    try:
        with open(path, 'rb') as f:
            data = f.read()
    except OSError as err:
        raise MeshError(str(err))
    ext = os.path.splitext(path)[1].lower()
    if ext == '.stl':
        return _load_stl(data)
    if ext == '.obj':
        return _load_obj(data)
    raise MeshError(f'unsupported mesh format "{ext}" (STL and OBJ are read)')


def convex_vertex_cloud(verts: np.ndarray) -> np.ndarray:
    """De-duplicated vertices; reduced to the hull's vertices when Qhull is available and the cloud is 3-D."""
    v = np.unique(np.asarray(verts, np.float64).reshape(-1, 3), axis=0)
    if v.shape[0] > 4:
        try:
            from scipy.spatial import ConvexHull, QhullError
            try:
                v = v[np.sort(ConvexHull(v).vertices)]
            except (QhullError, ValueError):
                pass                                   # flat or degenerate cloud: keep every point
        except ImportError:
            pass
    if v.shape[0] > MAX_SHAPE_VERTS:
        n_full = v.shape[0]
        v, dev = thin_vertex_cloud(v, MAX_SHAPE_VERTS)
        logger.warning(f'collision hull with {n_full} vertices thinned to {v.shape[0]} support points '
                       f'(the thinned hull lies inside the original, at most {dev * 1e3:.3f} mm from it)')
    return v


def _fibonacci_directions(n: int) -> np.ndarray:
    k = np.arange(n) + 0.5
    phi = np.arccos(1.0 - 2.0 * k / n)
    th = np.pi * (1.0 + 5.0 ** 0.5) * k
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)


def thin_vertex_cloud(verts: np.ndarray, budget: int) -> Tuple[np.ndarray, float]:
    """At most ``budget`` vertices of a convex cloud: the support points of ``budget`` evenly spread directions plus the
    six axis extremes (scaled to the cloud's bounding box so that thin shapes are sampled evenly too).  The result is a
    subset of the hull's vertices, so its hull lies inside the original; the second return value is the largest gap
    between the two support functions over a dense direction set (the Hausdorff distance of the hulls, sampled)."""
    v = np.asarray(verts, np.float64)
    c = 0.5 * (v.min(axis=0) + v.max(axis=0))
    ext = 0.5 * (v.max(axis=0) - v.min(axis=0))
    ext = np.maximum(ext, max(1e-3 * float(ext.max()), 1e-12))     # flat clouds: keep the direction set balanced
    dirs = _fibonacci_directions(max(budget - 6, 8)) / ext          # even coverage of the normalised shape
    dirs = np.concatenate([dirs, np.eye(3), -np.eye(3)])
    keep = np.unique(np.argmax((v - c) @ dirs.T, axis=0))
    thin = v[np.sort(keep)][:budget]
    probe = _fibonacci_directions(8192)
    gap = (v @ probe.T).max(axis=0) - (thin @ probe.T).max(axis=0)
    return thin, float(gap.max())


def cylinder_vertex_cloud(radius: float, length: float) -> np.ndarray:
    """The 2 x 32 rim points pybullet's importer feeds to btConvexHullShape for a <cylinder> (axis = local z)."""
    k = np.arange(CYLINDER_STEPS)
    ang = 2.0 * np.pi * k / CYLINDER_STEPS
    rim = np.stack([radius * np.sin(ang), radius * np.cos(ang)], axis=1)
    top = np.concatenate([rim, np.full((CYLINDER_STEPS, 1), 0.5 * length)], axis=1)
    bot = np.concatenate([rim, np.full((CYLINDER_STEPS, 1), -0.5 * length)], axis=1)
    return np.concatenate([top, bot], axis=0)


def resolve_mesh_file(uri: str, model_dir: str, search: Optional[List[str]] = None) -> str:
    """``filename`` / ``<uri>`` of a mesh element -> path: as given, relative to the model file, or below a search dir."""
    name = uri
    for prefix in ('package://', 'model://', 'file://'):
        if name.startswith(prefix):
            name = name[len(prefix):]
    cands = [name, os.path.join(model_dir, name)]
    parts = name.replace('\\', '/').split('/')
    for k in range(1, len(parts)):                     # package://<pkg>/meshes/x.stl -> <model_dir>/meshes/x.stl
        cands.append(os.path.join(model_dir, *parts[k:]))
    for d in search or []:
        cands.append(os.path.join(d, name))
    for c in cands:
        if os.path.isfile(c):
            return c
    raise MeshError(f'mesh file not found: {uri}')
