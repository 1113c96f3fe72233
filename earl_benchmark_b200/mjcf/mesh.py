"""Binary-STL loading and MuJoCo 2.1 ("legacy") mesh processing.

MuJoCo re-centres every mesh at compile time and the geom frame moves with it; the reference reads such a
frame as an OBSERVATION (`data.get_geom_xpos('handle')`, reference `earl_benchmark/envs/sawyer_door.py:113` and
metaworld's `_get_pos_objects`).  MuJoCo 2.1.0 does not use the exact signed-volume centre of mass: it takes the
area-weighted centroid of the triangle centroids, then the centre of |volume|-weighted tetrahedra built from
that point (SURVEY.md section 0.4 / Appendix E.1 -- reproduces the golden handle positions of
`sawyer_door.py:13-16` to 1.6e-8, whereas the exact centre is 5e-3 off).
"""
import struct

import numpy as np


def load_stl(path, scale=(1.0, 1.0, 1.0)):
    """Triangles [T,3,3] float64 of a binary STL (vertex data is float32 in the file)."""
    with open(path, "rb") as f:
        data = f.read()
    n = struct.unpack_from("<I", data, 80)[0]
    if 84 + 50 * n != len(data):
        raise ValueError(f"{path}: not a binary STL ({len(data)} bytes, {n} triangles declared)")
    rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
    return rec["v"].astype(np.float64) * np.asarray(scale, dtype=np.float64)


def legacy_center(tris):
    """Mesh frame origin used by MuJoCo 2.1.0 (see module docstring)."""
    v0, v1, v2 = tris[:, 0], tris[:, 1], tris[:, 2]
    area = 0.5 * np.linalg.norm(np.cross(v1 - v0, v2 - v0), axis=1)
    cen = (v0 + v1 + v2) / 3.0
    c_area = (area[:, None] * cen).sum(0) / area.sum()
    a, b, c = v0 - c_area, v1 - c_area, v2 - c_area
    vol = np.abs(np.einsum("ij,ij->i", a, np.cross(b, c))) / 6.0
    tet_cen = (a + b + c) / 4.0  # centroid of tetrahedron (c_area, v0, v1, v2), relative to c_area
    return c_area + (vol[:, None] * tet_cen).sum(0) / vol.sum()


def legacy_inertia(tris, center, density):
    """Mass and inertia tensor about `center` with the same |volume|-weighted tetrahedra (mesh frame axes)."""
    a, b, c = tris[:, 0] - center, tris[:, 1] - center, tris[:, 2] - center
    vol = np.abs(np.einsum("ij,ij->i", a, np.cross(b, c))) / 6.0
    # second moments of a tetrahedron with one vertex at the origin: integral x x^T dV = vol/20 * (sum_i p_i p_i^T + s s^T), s = a+b+c
    s = a + b + c
    P = np.einsum("i,ij,ik->jk", vol / 20.0, a, a) + np.einsum("i,ij,ik->jk", vol / 20.0, b, b) + \
        np.einsum("i,ij,ik->jk", vol / 20.0, c, c) + np.einsum("i,ij,ik->jk", vol / 20.0, s, s)
    mass = density * vol.sum()
    inertia = density * (np.trace(P) * np.eye(3) - P)
    return mass, inertia


def convex_hull_vertices(tris):
    """Unique vertices of the convex hull (what MuJoCo collides against for mesh geoms)."""
    from scipy.spatial import ConvexHull
    pts = np.unique(tris.reshape(-1, 3), axis=0)
    hull = ConvexHull(pts)
    return pts[hull.vertices]
