"""Shared test helpers."""
import numpy as np


def transversal_targets(pos, tri, origin):
    """Points on a closed mesh's vertices and edge midpoints that a ray from `origin` crosses
    transversally: every triangle incident to the vertex (or both triangles of the edge) faces the same
    way relative to the ray.  (At contour vertices/edges the ray only touches the mesh in a point, so
    a correctly rounded ray may legitimately miss; those are excluded.)"""
    p = np.asarray(pos, np.float64)
    tri = np.asarray(tri, np.int64).reshape(-1, 3)
    fn = np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]])
    o = np.asarray(origin, np.float64)
    nv = len(p)
    # vertices
    dv = p - o
    s = np.sign(np.einsum("tk,tck->tc", fn, dv[tri]))
    pos_cnt, neg_cnt = np.zeros(nv, int), np.zeros(nv, int)
    for k in range(3):
        np.add.at(pos_cnt, tri[:, k], (s[:, k] > 0).astype(int))
        np.add.at(neg_cnt, tri[:, k], (s[:, k] < 0).astype(int))
    vert_ok = ((pos_cnt == 0) | (neg_cnt == 0)) & ((pos_cnt + neg_cnt) > 0)
    targets = [p[vert_ok]]
    # edges: midpoint, both incident faces same sign
    edges = {}
    for t, (a, b, c) in enumerate(tri):
        for e in ((a, b), (b, c), (c, a)):
            edges.setdefault((min(e), max(e)), []).append(t)
    mids = []
    for (a, b), ts in edges.items():
        if len(ts) != 2:
            continue
        m = 0.5 * (p[a] + p[b])
        sg = [np.sign(fn[t] @ (m - o)) for t in ts]
        if sg[0] == sg[1] and sg[0] != 0:
            mids.append(m)
    targets.append(np.asarray(mids).reshape(-1, 3))
    return np.concatenate(targets)


def rays_toward(origin, targets):
    d = np.asarray(targets, np.float64) - np.asarray(origin, np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((len(d), 8), np.float32)
    rays[:, :3], rays[:, 4:7], rays[:, 3], rays[:, 7] = origin, d, 1e-5, 1e10
    return rays
