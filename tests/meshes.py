"""Procedural triangle meshes for the scene-preparation tests (closed, outward-facing; no file of the reference is read)."""
import numpy as np


def cone(segments=32, radius=1.0, height=2.0):
    """The shape of a Blender cone: a rim of `segments` vertices at y = -height/2, the apex at y = +height/2 and the base closed
    by a triangle fan around the first rim vertex: segments + 1 vertices, 2 * segments - 2 triangles (33 / 62 at the default,
    like the cone of the reference's shipped DFSPH scene)."""
    a = 2.0 * np.pi * np.arange(segments) / segments
    rim = np.stack([radius * np.sin(a), np.full(segments, -0.5 * height), -radius * np.cos(a)], 1)
    verts = np.concatenate([rim, [[0.0, 0.5 * height, 0.0]]]).astype(np.float32)
    apex = segments
    tris = [(i, apex, (i + 1) % segments) for i in range(segments)]
    tris += [(0, i, i + 1) for i in range(1, segments - 1)]
    tris = np.array(tris, np.uint32)
    return verts, _outward(verts, tris)


def torus(major=1.0, minor=0.35, nu=24, nv=12):
    u = 2.0 * np.pi * np.arange(nu) / nu
    v = 2.0 * np.pi * np.arange(nv) / nv
    U, V = np.meshgrid(u, v, indexing="ij")
    verts = np.stack([(major + minor * np.cos(V)) * np.cos(U), minor * np.sin(V), (major + minor * np.cos(V)) * np.sin(U)], -1).reshape(-1, 3).astype(np.float32)
    tris = []
    for i in range(nu):
        for j in range(nv):
            a, b = i * nv + j, i * nv + (j + 1) % nv
            c, d = ((i + 1) % nu) * nv + j, ((i + 1) % nu) * nv + (j + 1) % nv
            tris += [(a, b, d), (a, d, c)]
    tris = np.array(tris, np.uint32)
    return verts, _outward(verts, tris)


def box(lo, hi):
    """The 8 vertices / 12 triangles of the reference's TriangleMesh(AABB) (TriangleMesh.cpp:18-37), same order."""
    a, b = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    verts = np.array([[a[0], a[1], a[2]], [b[0], a[1], a[2]], [b[0], a[1], b[2]], [a[0], a[1], b[2]],
                      [a[0], b[1], a[2]], [b[0], b[1], a[2]], [b[0], b[1], b[2]], [a[0], b[1], b[2]]], np.float32)
    tris = np.array([[0, 1, 2], [0, 2, 3], [4, 7, 6], [4, 6, 5], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7]], np.uint32)
    return verts, tris


def _outward(verts, tris):
    """Flips the whole mesh if its signed volume is negative (consistent winding is the generator's job)."""
    v = verts.astype(np.float64)
    vol = np.einsum("ij,ij->i", v[tris[:, 0]], np.cross(v[tris[:, 1]], v[tris[:, 2]])).sum() / 6.0
    return tris if vol > 0 else np.ascontiguousarray(tris[:, ::-1])


def transform(scale=(1.0, 1.0, 1.0), rotate_x_deg=0.0, translate=(0.0, 0.0, 0.0)):
    """Row-major 4x4: translate * rotate_x * scale."""
    c, s = np.cos(np.radians(rotate_x_deg)), np.sin(np.radians(rotate_x_deg))
    R = np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], np.float64)
    S = np.diag([scale[0], scale[1], scale[2], 1.0])
    T = np.eye(4); T[:3, 3] = translate
    return (T @ R @ S).astype(np.float32)
