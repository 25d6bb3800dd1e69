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


# ---- closed forms of the rendering equation (independent of any restatement of the reference) ----------------------
def rect_form_factor_center(a, b, h):
    """Form factor from a differential element to a parallel (2a x 2b) rectangle centred at height h above it
    (four corner configurations; e.g. Howell, A Catalog of Radiation Heat Transfer Configuration Factors, B-3)."""
    import math

    def corner(a, b, h):
        A, B = math.sqrt(a * a + h * h), math.sqrt(b * b + h * h)
        return (a / A * math.atan(b / A) + b / B * math.atan(a / B)) / (2 * math.pi)
    return 4 * corner(a, b, h)


DIRECT_ALBEDO, DIRECT_RADIANCE, DIRECT_HEIGHT = (0.6, 0.5, 0.4), (3.0, 2.0, 1.0), 1.5
DIRECT_DIRECTION = (0.3, 1.0, 0.2)  # towards the distant light


def direct_light_scene(kind, spp):
    """A large lambertian plane lit by one light 1.5 above the origin, seen through a 0.5 degree camera aimed at the
    origin, path depth 2: every pixel estimates the direct illumination of (almost) the same point.
    kind 'rect': a 2 x 1 one-sided rectangle facing down -> Lo = albedo * L * F (form factor above);
    kind 'point': Lo = albedo / pi * I / h^2 (the reference divides by dist + EPS and dist^2 + EPS, A.3-12: -0.2 %);
    kind 'distant': Lo = albedo / pi * L * cos(theta)."""
    from asuna_b200 import host, scenes, structs as S
    sc = host.Scene()
    sc.set_camera("perspective", 8, 8, fov=0.5)
    sc.state["spp"], sc.state["maxPathDepth"] = spp, 2
    sc.add_material("m", scenes.mat(S.MAT_LAMBERTIAN, diffuse=DIRECT_ALBEDO))
    sc.add_mesh("plane", *scenes.grid_plane(2, 50.0, 0.0))
    sc.add_instance("plane", "m")
    h = DIRECT_HEIGHT
    if kind == "rect":
        sc.add_light(scenes.rect_light((-1, h, -0.5), (1, h, -0.5), (-1, h, 0.5), DIRECT_RADIANCE))
    elif kind == "distant":
        sc.add_light(scenes.distant_light(DIRECT_DIRECTION, DIRECT_RADIANCE))
    else:
        sc.add_light(scenes.point_light((0, h, 0), DIRECT_RADIANCE))
    sc.shots.append(host.Shot((3.0, 2.0, 0.0), (0, 0, 0), (0, 1, 0)))
    return sc


def direct_light_expected(kind):
    import math
    alb, L = np.asarray(DIRECT_ALBEDO), np.asarray(DIRECT_RADIANCE)
    if kind == "rect":
        return alb * L * rect_form_factor_center(1.0, 0.5, DIRECT_HEIGHT)
    if kind == "distant":  # irradiance L cos(theta) of a directional source
        d = np.asarray(DIRECT_DIRECTION, np.float64)
        return alb / math.pi * L * (d[1] / np.linalg.norm(d))
    return alb / math.pi * L / DIRECT_HEIGHT ** 2


def mirror_plane_scene():
    """A mirror plane (reflectance 0.9, 0.8, 0.7) under a constant background (0.5, 0.25, 0.125): every pixel is the
    product, exactly (delta reflection, no MIS on specular paths, raytrace.default.rmiss:24-55)."""
    from asuna_b200 import host, scenes, structs as S
    sc = host.Scene()
    sc.set_camera("perspective", 16, 16, fov=30.0)
    sc.state["spp"], sc.state["maxPathDepth"] = 4, 3
    sc.state["bgColor"] = (0.5, 0.25, 0.125)
    sc.add_material("m", scenes.mat(S.MAT_MIRROR, diffuse=(0.9, 0.8, 0.7)))
    sc.add_mesh("plane", *scenes.grid_plane(2, 50.0, 0.0))
    sc.add_instance("plane", "m")
    sc.shots.append(host.Shot((3.0, 2.0, 0.0), (0, 0, 0), (0, 1, 0)))
    return sc


# ----------------------------------------------------------------------------- shade probes (oracle vs oracle/_ref)
import ctypes as _C  # noqa: E402

# ShadeProbe of oracle/refbuild/ref_bridge.h
PROBE = np.dtype([("ray_o", "f4", 3), ("ray_d", "f4", 3), ("radiance", "f4", 3), ("throughput", "f4", 3),
                  ("depth", "u4"), ("seed", "u4"), ("stop", "u4"), ("brec_d", "f4", 3), ("brec_pdf", "f4"),
                  ("brec_flags", "u4"), ("drec_radiance", "f4", 3), ("drec_dist", "f4"), ("drec_o", "f4", 3),
                  ("drec_d", "f4", 3), ("drec_skip", "u4"), ("channel", "f4", (8, 3))])
MISS = 0xFFFFFFFF


def random_probes(rng, n, ntri, inst):
    """n random shader invocations on instance `inst` (MISS = the miss shader): random triangle, barycentrics,
    incoming direction (both sides of the surface), RNG state, depth, throughput and previous-bounce record."""
    q = np.zeros(n, PROBE)
    d = rng.randn(n, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    q["ray_d"], q["ray_o"] = d, rng.rand(n, 3)
    q["throughput"], q["radiance"] = rng.rand(n, 3), rng.rand(n, 3) * 0.1
    q["depth"] = rng.randint(1, 4, n)
    q["seed"] = rng.randint(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    q["brec_pdf"] = rng.rand(n) * 3
    q["brec_flags"] = rng.choice([0, 1, 4, 16, 32], n)
    b1 = rng.rand(n).astype(np.float32)
    b2 = ((1 - b1) * rng.rand(n)).astype(np.float32)
    return (np.full(n, inst, np.uint32), rng.randint(0, max(ntri, 1), n).astype(np.uint32), b1, b2, q)


def run_probes(ctx, inst, prim, b1, b2, probes):
    q = probes.copy()
    p = lambda a: a.ctypes.data_as(_C.c_void_p)
    ctx._call("shade_probes", _C.c_uint32(len(q)), p(inst), p(prim), p(b1), p(b2), p(q))
    return q


def probe_mismatch(A, B, tol=2e-4):
    """Compares two probe result arrays.  Integer fields (RNG state after the shader = number of draws consumed,
    depth, stop, sampled-lobe flags, shadow-ray skip) must be identical; float fields are compared as vectors,
    |a - b| <= tol * max(1, |b|).  Returns {field: fraction of probes outside}."""
    out = {}
    for f in PROBE.names:
        a, b = A[f], B[f]
        if a.dtype.kind == "u":
            bad = a != b
        else:
            a, b = a.reshape(len(a), -1).astype(np.float64), b.reshape(len(b), -1).astype(np.float64)
            fa, fb = np.isfinite(a).all(axis=1), np.isfinite(b).all(axis=1)
            with np.errstate(invalid="ignore"):
                d = np.linalg.norm(np.where(np.isfinite(a - b), a - b, 0), axis=1)
                scale = np.maximum(1.0, np.linalg.norm(np.where(np.isfinite(b), b, 0), axis=1))
            bad = (fa != fb) | (fa & fb & (d > tol * scale))
        if bad.any():
            out[f] = float(bad.mean())
    return out


def random_material(rng, mtype, texture_ids):
    """A random but valid GpuMaterial of the given type (parameter ranges of SURVEY.md 8d)."""
    from asuna_b200 import structs as S
    m = S.default_material()
    m["type"] = mtype
    m["diffuse"] = rng.uniform(0.05, 0.9, 3)
    m["rhoSpec"] = rng.uniform(0.05, 0.9, 3)
    m["anisoAlpha"] = rng.uniform(0.02, 0.9, 2)
    m["ior"] = rng.uniform(1.05, 2.4)
    m["roughness"] = rng.uniform(0.02, 0.95)
    m["metalness"] = rng.uniform(0, 1)
    for k in ("subsurface", "specularTint", "anisotropic", "sheen", "sheenTint", "clearcoat", "clearcoatGloss"):
        m[k] = rng.uniform(0, 1)
    m["specular"] = rng.uniform(0, 1)
    m["radiance"] = rng.uniform(0.1, 4.0, 3)
    m["radianceFactor"] = rng.uniform(0.5, 5.0, 3)
    pick = lambda: int(rng.choice(texture_ids)) if (len(texture_ids) and rng.rand() < 0.5) else -1
    for k in ("diffuseTextureId", "roughnessTextureId", "metalnessTextureId", "radianceTextureId", "normalTextureId",
              "tangentTextureId", "opacityTextureId"):
        m[k] = pick()
    if mtype == S.MAT_PHONG:
        m["specular"] = rng.uniform(2, 200)  # shininess
    if mtype in (S.MAT_PLASTIC, S.MAT_ROUGH_PLASTIC):
        m["radiance"][0] = rng.uniform(0.3, 0.9)  # fdrInt
    # constant opacity (= pass-through probability) is aliased onto specular (pbr), metalness (kang18), rhoSpec.x (disney)
    opacity = rng.uniform(0, 0.5) if rng.rand() < 0.5 else 0.0
    if mtype == S.MAT_PBR:
        m["specular"] = opacity
    elif mtype == S.MAT_KANG18:
        m["metalness"] = opacity
    elif mtype == S.MAT_DISNEY:
        m["rhoSpec"][0] = opacity
    return m


# ----------------------------------------------------------------------------- compressed wide BVH, walked on the host
WIDE_NODE = np.dtype([("p", "f4", 3), ("e", "u1", 3), ("imask", "u1"), ("child_base", "u4"), ("prim_base", "u4"),
                      ("meta", "u1", 8), ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])  # asuna_b200/csrc/device_types.cuh
TRI_SLOT = np.dtype([("v0", "f4", 3), ("prim", "u4"), ("v1", "f4", 3), ("inst", "u4"), ("v2", "f4", 3), ("pad", "u4")])
assert WIDE_NODE.itemsize == 80 and TRI_SLOT.itemsize == 48


def download_accel(ctx):
    sizes = (_C.c_uint64 * 3)()
    ctx.L.lib.asuna_debug_accel_sizes(ctx.h, sizes)
    nodes, tris = np.zeros(sizes[0], WIDE_NODE), np.zeros(sizes[1], TRI_SLOT)
    rc = ctx.L.lib.asuna_debug_download_accel(ctx.h, nodes.ctypes.data_as(_C.c_void_p), tris.ctypes.data_as(_C.c_void_p))
    assert rc == 0
    return nodes, tris, int(sizes[2])


def walk_wide_bvh(nodes, tris, root, c_node=1.0, c_prim=1.0):
    """Breadth-first walk of one compressed 8-wide BVH.  Returns
      refs        : how often each triangle slot is referenced by a leaf child,
      outside     : number of triangle vertices outside the decoded box of ANY ancestor slot (must be 0: the boxes are
                    conservative at every level, which is what makes the quantised slab test safe),
      nodes_used, max_depth, sah (sum over slots of area x cost, / root area, from the decoded boxes)."""
    refs = np.zeros(len(tris), np.int64)
    frontier = np.array([root], np.int64)
    clip_lo = np.full((1, 3), -np.inf)
    clip_hi = np.full((1, 3), np.inf)
    outside = nodes_used = depth = 0
    sah = 0.0
    root_area = None
    below = (1 << np.arange(8)) - 1
    while frontier.size:
        depth += 1
        nodes_used += frontier.size
        nd = nodes[frontier]
        scale = (nd["e"].astype(np.uint32) << 23).view(np.float32).astype(np.float64)  # (n, 3)
        p = nd["p"].astype(np.float64)
        lo = p[:, :, None] + nd["qlo"].astype(np.float64) * scale[:, :, None]  # (n, 3, 8)
        hi = p[:, :, None] + nd["qhi"].astype(np.float64) * scale[:, :, None]
        meta = nd["meta"].astype(np.int64)  # (n, 8)
        used = meta != 0
        inner = used & ((meta & 31) >= 24)
        leaf = used & ~inner
        lo = np.maximum(lo, clip_lo[:, :, None])
        hi = np.minimum(hi, clip_hi[:, :, None])
        ext = np.maximum(hi - lo, 0.0)
        area = ext[:, 0] * ext[:, 1] + ext[:, 1] * ext[:, 2] + ext[:, 2] * ext[:, 0]  # (n, 8)
        if root_area is None:
            ulo, uhi = np.where(used[:, None, :], lo, np.inf).min(axis=2), np.where(used[:, None, :], hi, -np.inf).max(axis=2)
            e = uhi[0] - ulo[0]
            root_area = e[0] * e[1] + e[1] * e[2] + e[2] * e[0]
            sah += root_area * c_node
        count = np.where(leaf, ((meta >> 5) & 1) + ((meta >> 6) & 1) + ((meta >> 7) & 1), 0)
        sah += float((area * inner).sum() * c_node + (area * count).sum() * c_prim)
        # leaves: triangles inside the running intersection of all ancestor boxes
        ni, si = np.nonzero(leaf)
        for k in range(3):
            sel = count[ni, si] > k
            a, s = ni[sel], si[sel]
            slot = nd["prim_base"][a].astype(np.int64) + (meta[a, s] & 31) + k
            np.add.at(refs, slot, 1)
            for v in ("v0", "v1", "v2"):
                x = tris[v][slot].astype(np.float64)
                outside += int(((x < lo[a, :, s]) | (x > hi[a, :, s])).any(axis=1).sum())
        # inner children: next frontier with their clip boxes
        ni, si = np.nonzero(inner)
        rel = np.array([bin(int(m) & int(b)).count("1") for m, b in zip(nd["imask"][ni], below[si])], np.int64) if ni.size else np.zeros(0, np.int64)
        assert ((nd["imask"][ni].astype(np.int64) >> si) & 1).all(), "inner slot not in imask"
        frontier = nd["child_base"][ni].astype(np.int64) + rel
        clip_lo, clip_hi = lo[ni, :, si], hi[ni, :, si]
    return {"refs": refs, "outside": outside, "nodes_used": nodes_used, "max_depth": depth, "sah": sah / root_area}
