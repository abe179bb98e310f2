"""Deterministic benchmark meshes: conforming subdivisions of a closed triangle mesh.

The reference's benchmark folders are named after bunny face counts (56 172 x {3, 6, 12, 24, 48, 96, 192}); the
meshes themselves are not shipped.  They are rebuilt here from the 56 172-face bunny with three conforming
operators (every one keeps the mesh closed and consistently oriented):
    M  edge-midpoint split      1 -> 4   new vertex (a + b) * 0.5f per undirected edge
    C  centroid split           1 -> 3   new vertex (a + b + c) / 3.0f per face
    B  barycentric split        1 -> 6   edge midpoints + centroid
1 348 128 faces = M(B(bunny)),  10 785 024 faces = M(M(M(C(bunny)))).  All arithmetic is float32, no RNG.
"""
from __future__ import annotations

import numpy as np

RECIPES = {
    168516: "C", 337032: "B", 674064: "CM", 1348128: "BM", 2696256: "CMM", 5392128: "BMM", 10785024: "CMMM",
}


def _edge_midpoints(verts: np.ndarray, tris: np.ndarray):
    """Returns (new_verts, mid[T,3]) with mid[t,i] = index of the midpoint of edge (tris[t,i], tris[t,(i+1)%3])."""
    t = tris.astype(np.int64)
    e = np.stack([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=1).reshape(-1, 2)
    lo, hi = e.min(axis=1), e.max(axis=1)
    key = lo * verts.shape[0] + hi
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    a = verts[lo[first]]
    b = verts[hi[first]]
    mids = ((a + b) * np.float32(0.5)).astype(np.float32)
    mid_idx = (inv.reshape(-1, 3) + verts.shape[0]).astype(np.uint32)
    return np.concatenate([verts, mids]).astype(np.float32), mid_idx


def midpoint(verts, tris):
    v, m = _edge_midpoints(verts, tris)
    a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
    ab, bc, ca = m[:, 0], m[:, 1], m[:, 2]
    out = np.stack([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1), np.stack([ca, bc, c], 1),
                    np.stack([ab, bc, ca], 1)], axis=1).reshape(-1, 3)
    return v, out.astype(np.uint32)


def centroid(verts, tris):
    a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
    cen = ((verts[a] + verts[b] + verts[c]) / np.float32(3.0)).astype(np.float32)
    g = (np.arange(tris.shape[0]) + verts.shape[0]).astype(np.uint32)
    out = np.stack([np.stack([a, b, g], 1), np.stack([b, c, g], 1), np.stack([c, a, g], 1)], axis=1).reshape(-1, 3)
    return np.concatenate([verts, cen]).astype(np.float32), out.astype(np.uint32)


def barycentric(verts, tris):
    v, m = _edge_midpoints(verts, tris)
    a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
    ab, bc, ca = m[:, 0], m[:, 1], m[:, 2]
    cen = ((verts[a] + verts[b] + verts[c]) / np.float32(3.0)).astype(np.float32)
    g = (np.arange(tris.shape[0]) + v.shape[0]).astype(np.uint32)
    out = np.stack([np.stack([a, ab, g], 1), np.stack([ab, b, g], 1), np.stack([b, bc, g], 1),
                    np.stack([bc, c, g], 1), np.stack([c, ca, g], 1), np.stack([ca, a, g], 1)], axis=1).reshape(-1, 3)
    return np.concatenate([v, cen]).astype(np.float32), out.astype(np.uint32)


_OPS = {"M": midpoint, "C": centroid, "B": barycentric}


def subdivide(verts, tris, recipe: str):
    verts = np.ascontiguousarray(verts, np.float32)
    tris = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
    for op in recipe:
        verts, tris = _OPS[op](verts, tris)
    return verts, tris


def bunny_with_faces(verts, tris, faces: int):
    """The benchmark mesh with the named face count, derived from the 56 172-face bunny."""
    if faces == tris.shape[0]:
        return verts, tris
    if tris.shape[0] != 56172 or faces not in RECIPES:
        raise ValueError(f"no recipe for {faces} faces from a {tris.shape[0]}-face mesh")
    return subdivide(verts, tris, RECIPES[faces])


def is_closed(tris) -> bool:
    """Every undirected edge shared by exactly two faces, with opposite directions (2-manifold, oriented)."""
    t = np.asarray(tris, np.int64).reshape(-1, 3)
    e = np.stack([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=1).reshape(-1, 2)
    nv = int(t.max()) + 1
    fwd = e[:, 0] * nv + e[:, 1]
    bwd = e[:, 1] * nv + e[:, 0]
    return np.unique(fwd).size == fwd.size and np.array_equal(np.sort(fwd), np.sort(bwd))
