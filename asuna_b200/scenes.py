"""Deterministic synthetic scenes of the shapes BASELINE.json names.

The reference ships no scenes (CMakeLists.txt:347-354 has the copy step commented out), so
every config is generated here with fixed seeds.  Each builder returns a host.Scene; use
tools/gen_scenes.py to also write the JSON + OBJ + texture files the C++ host loads.
"""
import math

import numpy as np

from . import structs as S
from .host import Scene, Shot, make_vertices, f32, rotation_y, translation, scaling, mat_identity


# ----------------------------------------------------------------------------- mesh helpers
def quad(p0, p1, p2, p3, uv_scale=1.0):
    """Two triangles (0,1,2),(0,2,3) with the geometric normal on every vertex."""
    p = np.asarray([p0, p1, p2, p3], f32)
    n = np.cross(p[1] - p[0], p[2] - p[0])
    n = n / np.linalg.norm(n)
    uv = np.asarray([[0, 0], [1, 0], [1, 1], [0, 1]], f32) * f32(uv_scale)
    return make_vertices(p, uv, np.tile(n, (4, 1))), np.asarray([0, 1, 2, 0, 2, 3], np.uint32)


def merge(parts):
    vs, idx, base = [], [], 0
    for v, i in parts:
        vs.append(v)
        idx.append(np.asarray(i, np.uint32) + np.uint32(base))
        base += v.size
    return np.concatenate(vs), np.concatenate(idx)


def box(lo, hi):
    (x0, y0, z0), (x1, y1, z1) = lo, hi
    faces = [
        ((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)),  # +z
        ((x1, y0, z0), (x0, y0, z0), (x0, y1, z0), (x1, y1, z0)),  # -z
        ((x1, y0, z1), (x1, y0, z0), (x1, y1, z0), (x1, y1, z1)),  # +x
        ((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)),  # -x
        ((x0, y1, z1), (x1, y1, z1), (x1, y1, z0), (x0, y1, z0)),  # +y
        ((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)),  # -y
    ]
    return merge([quad(*f) for f in faces])


def icosphere(subdiv):
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = np.asarray(v, np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.asarray(f, np.int64)
    for _ in range(subdiv):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        e.sort(axis=1)
        ue, inv = np.unique(e, axis=0, return_inverse=True)
        mid = v[ue[:, 0]] + v[ue[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid])
        n = len(f)
        a, b, c = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], a, c], 1), np.stack([f[:, 1], b, a], 1), np.stack([f[:, 2], c, b], 1),
                            np.stack([a, b, c], 1)])
    return v, f


def _value_noise(p, seed, octaves=3):
    """Smooth pseudo-random displacement field on points p (n,3): sum of seeded sinusoid octaves."""
    rng = np.random.RandomState(seed)
    out = np.zeros(len(p))
    for o in range(octaves):
        k = rng.normal(size=(6, 3)) * (2.0 ** o) * 2.5
        ph = rng.uniform(0, 2 * math.pi, 6)
        out += (np.sin(p @ k.T + ph).sum(1) / 6.0) * (0.5 ** o)
    return out


def blob(subdiv, seed=1234, amplitude=0.18, radius=1.0):
    """Seeded displaced icosphere with smooth vertex normals (stand-in for the bunny asset)."""
    v, f = icosphere(subdiv)
    r = radius * (1.0 + amplitude * _value_noise(v, seed))
    p = v * r[:, None]
    fn = np.cross(p[f[:, 1]] - p[f[:, 0]], p[f[:, 2]] - p[f[:, 0]])
    n = np.zeros_like(p)
    for k in range(3):
        np.add.at(n, f[:, k], fn)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    theta = np.arccos(np.clip(v[:, 1], -1, 1)) / math.pi
    phi = (np.arctan2(v[:, 2], v[:, 0]) + math.pi) / (2 * math.pi)
    return make_vertices(p, np.stack([phi, theta], 1), n), f.astype(np.uint32).reshape(-1)


def grid_plane(n, size, y=0.0):
    """n x n quads in the xz-plane, facing +y."""
    xs = np.linspace(-size, size, n + 1)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    p = np.stack([X.ravel(), np.full(X.size, y), Z.ravel()], 1)
    uv = np.stack([(X.ravel() / size + 1) / 2, (Z.ravel() / size + 1) / 2], 1)
    nrm = np.tile([0, 1, 0], (len(p), 1))
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    a = (j * (n + 1) + i).ravel()
    idx = np.stack([a, a + n + 1, a + n + 2, a, a + n + 2, a + 1], 1)
    return make_vertices(p, uv, nrm), idx.astype(np.uint32).reshape(-1)


# ----------------------------------------------------------------------------- materials
def mat(type_, **kw):
    m = S.default_material()
    m["type"] = type_
    for k, v in kw.items():
        m[k] = v
    return m


def rect_light(position, v1, v2, radiance, double_side=False):
    """JSON rect light: three corner points (reference src/loader/loader.cpp:284-296)."""
    l = np.zeros((), S.Light)
    l["type"] = S.LIGHT_RECT
    l["radiance"] = radiance
    l["position"] = position
    l["u"] = np.asarray(v1, f32) - np.asarray(position, f32)
    l["v"] = np.asarray(v2, f32) - np.asarray(position, f32)
    l["area"] = f32(np.linalg.norm(np.cross(l["u"], l["v"]).astype(f32)))
    l["doubleSide"] = 1 if double_side else 0
    return l


def point_light(position, radiance):
    l = np.zeros((), S.Light)
    l["type"] = S.LIGHT_POINT
    l["position"], l["radiance"] = position, radiance
    return l


def distant_light(direction, radiance):
    l = np.zeros((), S.Light)
    l["type"] = S.LIGHT_DIRECTIONAL
    l["direction"], l["radiance"] = direction, radiance
    return l


# ----------------------------------------------------------------------------- textures / env
def procedural_envmap(w=2048, h=1024, seed=7):
    """Vertical gradient + grid lines + one bright Gaussian 'sun' blob."""
    rng = np.random.RandomState(seed)
    v = (np.arange(h) + 0.5) / h
    u = (np.arange(w) + 0.5) / w
    U, V = np.meshgrid(u, v)
    sky = np.stack([0.35 + 0.4 * (1 - V), 0.45 + 0.4 * (1 - V), 0.6 + 0.5 * (1 - V)], 2)
    ground = np.stack([0.25 + 0 * V, 0.22 + 0 * V, 0.2 + 0 * V], 2)
    img = np.where((V < 0.5)[..., None], sky, ground)
    lines = ((np.abs(((U * 16) % 1.0) - 0.5) > 0.47) | (np.abs(((V * 8) % 1.0) - 0.5) > 0.47))
    img = img * np.where(lines[..., None], 0.55, 1.0)
    su, sv = 0.2 + 0.6 * rng.rand(), 0.18 + 0.1 * rng.rand()
    du = np.minimum(np.abs(U - su), 1 - np.abs(U - su))
    sun = 60.0 * np.exp(-((du * 2) ** 2 + (V - sv) ** 2) / (2 * 0.012 ** 2))
    img = img + sun[..., None] * np.asarray([1.0, 0.93, 0.8])
    return np.concatenate([img, np.ones((h, w, 1))], 2).astype(f32)


def noise_texture(n, seed, channels=3, lo=0.05, hi=0.9, octaves=4):
    rng = np.random.RandomState(seed)
    img = np.zeros((n, n, channels))
    for o in range(octaves):
        m = 4 * 2 ** o
        g = rng.rand(m, m, channels)
        g = np.concatenate([g, g[:1]], 0)
        g = np.concatenate([g, g[:, :1]], 1)
        x = np.linspace(0, m, n, endpoint=False)
        i = x.astype(int)
        fr = x - i
        fr = fr * fr * (3 - 2 * fr)
        a = g[i][:, i] * (1 - fr)[None, :, None] + g[i][:, i + 1] * fr[None, :, None]
        b = g[i + 1][:, i] * (1 - fr)[None, :, None] + g[i + 1][:, i + 1] * fr[None, :, None]
        img += (a * (1 - fr)[:, None, None] + b * fr[:, None, None]) * 0.5 ** o
    img /= img.max()
    img = lo + (hi - lo) * img
    out = np.ones((n, n, 4), f32)
    out[..., :channels] = img
    if channels == 1:
        out[..., 1] = out[..., 2] = out[..., 0]
    return out


# ----------------------------------------------------------------------------- configs
def _cornell_geometry(sc, white="white", red="red", green="green", short="white", tall="white"):
    fl = quad((0, 0, 1), (1, 0, 1), (1, 0, 0), (0, 0, 0))
    ce = quad((0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1))
    ba = quad((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0))
    le = quad((0, 0, 1), (0, 0, 0), (0, 1, 0), (0, 1, 1))
    ri = quad((1, 0, 0), (1, 0, 1), (1, 1, 1), (1, 1, 0))
    sc.add_mesh("walls", *merge([fl, ce, ba]))
    sc.add_mesh("left", *le)
    sc.add_mesh("right", *ri)
    sc.add_mesh("unit_box", *box((-0.5, 0, -0.5), (0.5, 1, 0.5)))
    sc.add_instance("walls", white)
    sc.add_instance("left", red)
    sc.add_instance("right", green)
    xs = translation((0.67, 0, 0.62)) @ rotation_y(math.radians(-17)) @ scaling((0.3, 0.3, 0.3))
    xt = translation((0.34, 0, 0.33)) @ rotation_y(math.radians(19)) @ scaling((0.3, 0.6, 0.3))
    sc.add_instance("unit_box", short, xs)
    sc.add_instance("unit_box", tall, xt)


def cornell(width=512, height=512, spp=64, depth=5, channels=("diffuse", "normal", "position")):
    """C1: Cornell box, lambertian + one rectangle area light."""
    sc = Scene()
    sc.set_camera("perspective", width, height, fov=39.3)
    sc.state["spp"], sc.state["maxPathDepth"] = spp, depth
    sc.set_channels(list(channels))
    sc.add_material("white", mat(S.MAT_LAMBERTIAN, diffuse=(.725, .71, .68)))
    sc.add_material("red", mat(S.MAT_LAMBERTIAN, diffuse=(.63, .065, .05)))
    sc.add_material("green", mat(S.MAT_LAMBERTIAN, diffuse=(.14, .45, .091)))
    sc.add_light(rect_light((0.35, 0.999, 0.35), (0.65, 0.999, 0.35), (0.35, 0.999, 0.65), (17, 12, 4)))
    _cornell_geometry(sc)
    sc.shots.append(Shot((0.5, 0.5, 2.4), (0.5, 0.5, 0), (0, 1, 0)))
    return sc


def cornell_materials(width=256, height=256, spp=16, depth=5, env=False, lights="rect", textured=False, seed=3):
    """Small Cornell variant that exercises every in-scope BSDF, light type and texture slot."""
    sc = Scene()
    sc.set_camera("perspective", width, height, fov=39.3)
    sc.state["spp"], sc.state["maxPathDepth"] = spp, depth
    sc.set_channels(["diffuse", "normal", "specular", "tangent", "roughness", "position", "uv"])
    tex = {}
    if textured:
        tex["albedo"] = sc.add_texture("albedo", noise_texture(64, seed))
        tex["rough"] = sc.add_texture("rough", noise_texture(64, seed + 1, 1, 0.1, 0.8))
        tex["metal"] = sc.add_texture("metal", noise_texture(64, seed + 2, 1, 0.0, 1.0))
        nm = noise_texture(64, seed + 3, 3, 0.35, 0.65)
        nm[..., 2] = 1.0
        tex["normal"] = sc.add_texture("normal", nm)
        op = noise_texture(64, seed + 4, 1, 0.0, 0.6)
        tex["opacity"] = sc.add_texture("opacity", op)
    t = lambda k: tex.get(k, -1)
    sc.add_material("white", mat(S.MAT_LAMBERTIAN, diffuse=(.725, .71, .68), diffuseTextureId=t("albedo"),
                                 normalTextureId=t("normal")))
    sc.add_material("red", mat(S.MAT_PLASTIC, diffuse=(.63, .065, .05), ior=1.49))
    sc.materials[-1]["radiance"][0] = _fdr(1.49)
    sc.add_material("green", mat(S.MAT_ROUGH_PLASTIC, diffuse=(.14, .45, .091), ior=1.6, anisoAlpha=(0.15, 0.3)))
    sc.materials[-1]["radiance"][0] = _fdr(1.6)
    from .host import COMPLEX_IOR
    sc.add_material("gold", mat(S.MAT_CONDUCTOR, diffuse=(1, 1, 1), radiance=COMPLEX_IOR["Au"][0],
                                radianceFactor=COMPLEX_IOR["Au"][1]))
    sc.add_material("glass", mat(S.MAT_DIELECTRIC, ior=1.5))
    sc.add_material("pbr", mat(S.MAT_PBR, diffuse=(.8, .5, .3), metalness=0.6, roughness=0.35, specular=0.0,
                               diffuseTextureId=t("albedo"), roughnessTextureId=t("rough"),
                               metalnessTextureId=t("metal"), normalTextureId=t("normal"),
                               opacityTextureId=t("opacity")))
    # kang18 takes the diffuse texture *or* the constant (reference src/loader/material.cpp:166-169)
    sc.add_material("kang", mat(S.MAT_KANG18, diffuse=(0, 0, 0) if textured else (.3, .4, .6), rhoSpec=(.5, .5, .4),
                                anisoAlpha=(.2, .08), metalness=0.0, diffuseTextureId=t("albedo")))
    sc.add_material("glow", mat(S.MAT_EMISSIVE, radiance=(2.0, 1.5, 3.0)))
    if lights in ("rect", "all"):
        sc.add_light(rect_light((0.35, 0.999, 0.35), (0.65, 0.999, 0.35), (0.35, 0.999, 0.65), (17, 12, 4)))
    if lights in ("point", "all"):
        sc.add_light(point_light((0.5, 0.8, 0.8), (0.6, 0.6, 0.7)))
    if lights in ("distant", "all"):
        sc.add_light(distant_light((0.2, 0.4, 1.0), (0.8, 0.7, 0.6)))
    if lights in ("mesh", "all"):
        v, i = quad((0.1, 0.998, 0.1), (0.25, 0.998, 0.1), (0.25, 0.998, 0.25), (0.1, 0.998, 0.25))
        sc.add_mesh_light((9, 9, 12), v, i[:3])
    if env:
        sc.set_envmap(procedural_envmap(64, 32, seed))
    fl = quad((0, 0, 1), (1, 0, 1), (1, 0, 0), (0, 0, 0), 2.0)
    ce = quad((0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1))
    ba = quad((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0))
    le = quad((0, 0, 1), (0, 0, 0), (0, 1, 0), (0, 1, 1))
    ri = quad((1, 0, 0), (1, 0, 1), (1, 1, 1), (1, 1, 0))
    sc.add_mesh("floor", *fl), sc.add_mesh("ceil", *ce), sc.add_mesh("back", *ba)
    sc.add_mesh("left", *le), sc.add_mesh("right", *ri)
    sc.add_mesh("unit_box", *box((-0.5, 0, -0.5), (0.5, 1, 0.5)))
    sc.add_mesh("ball", *blob(2, seed, 0.1))
    sc.add_instance("floor", "pbr"), sc.add_instance("back", "kang")
    if not env:
        sc.add_instance("ceil", "white")
    sc.add_instance("left", "red"), sc.add_instance("right", "green")
    sc.add_instance("unit_box", "gold", translation((0.7, 0, 0.6)) @ rotation_y(math.radians(-17)) @ scaling((0.28, 0.3, 0.28)))
    sc.add_instance("unit_box", "white", translation((0.3, 0, 0.3)) @ rotation_y(math.radians(19)) @ scaling((0.28, 0.6, 0.28)))
    sc.add_instance("ball", "glass", translation((0.68, 0.47, 0.6)) @ scaling((0.15, 0.15, 0.15)))
    sc.add_instance("ball", "glow", translation((0.15, 0.08, 0.8)) @ scaling((0.07, 0.07, 0.07)))
    sc.shots.append(Shot((0.5, 0.5, 2.4), (0.5, 0.5, 0), (0, 1, 0)))
    return sc


def cornell_all_materials(width=256, height=256, spp=16, depth=5, env=False, lights="rect", textured=False, seed=3):
    """cornell_materials plus the four remaining closest-hit shaders of the reference (mirror, rough_conductor,
    phong, disney -- SURVEY.md 8f row 1): all twelve material types in one frame."""
    sc = cornell_materials(width, height, spp, depth, env, lights, textured, seed)
    from .host import COMPLEX_IOR
    t = lambda k: sc.texture_ids.get(k, -1)
    sc.add_material("mirror", mat(S.MAT_MIRROR, diffuse=(.9, .92, .95)))
    sc.add_material("copper", mat(S.MAT_ROUGH_CONDUCTOR, diffuse=(1, 1, 1), radiance=COMPLEX_IOR["Cu"][0],
                                  radianceFactor=COMPLEX_IOR["Cu"][1], anisoAlpha=(0.25, 0.1), roughnessTextureId=-1))
    sc.add_material("phong", mat(S.MAT_PHONG, diffuse=(.2, .3, .7), rhoSpec=(.6, .6, .5), specular=40.0,
                                 diffuseTextureId=t("albedo")))
    sc.add_material("disney", mat(S.MAT_DISNEY, diffuse=(.7, .25, .2), metalness=0.3, roughness=0.4, subsurface=0.2,
                                  specularTint=0.3, anisotropic=0.5, sheen=0.4, sheenTint=0.6, clearcoat=0.8,
                                  clearcoatGloss=0.7, ior=1.45, normalTextureId=t("normal"), opacityTextureId=t("opacity")))
    sc.materials[-1]["rhoSpec"][0] = 0.0 if textured else 0.15  # disney keeps its opacity constant in rhoSpec.x
    sc.add_instance("ball", "mirror", translation((0.22, 0.72, 0.32)) @ scaling((0.1, 0.1, 0.1)))
    sc.add_instance("ball", "copper", translation((0.45, 0.12, 0.78)) @ scaling((0.11, 0.11, 0.11)))
    sc.add_instance("unit_box", "phong", translation((0.82, 0.3, 0.25)) @ rotation_y(math.radians(30)) @ scaling((0.16, 0.22, 0.16)))
    sc.add_instance("ball", "disney", translation((0.5, 0.38, 0.45)) @ scaling((0.12, 0.12, 0.12)))
    return sc


_FDR_CACHE = {}


def _fdr(ior):
    from .host import compute_diffuse_fresnel
    if ior not in _FDR_CACHE:
        _FDR_CACHE[ior] = compute_diffuse_fresnel(ior, 1000)
    return _FDR_CACHE[ior]


def glass_blob(width=1920, height=1080, spp=256, depth=8, subdiv=6, env_size=(2048, 1024)):
    """C2: dielectric blob inside an open glass bowl on a lambertian ground, env-map lit."""
    sc = Scene()
    sc.set_camera("perspective", width, height, fov=45.0)
    sc.state["spp"], sc.state["maxPathDepth"], sc.state["envMapIntensity"] = spp, depth, 1.0
    sc.set_channels(["diffuse", "normal", "position"])
    sc.add_material("glass", mat(S.MAT_DIELECTRIC, ior=1.5))
    sc.add_material("ground", mat(S.MAT_LAMBERTIAN, diffuse=(.55, .5, .45)))
    sc.set_envmap(procedural_envmap(*env_size))
    sc.add_mesh("blob", *blob(subdiv, 1234))
    sc.add_mesh("bowl", *_bowl(128, 64))
    sc.add_mesh("ground", *grid_plane(8, 6.0, 0.0))
    sc.add_instance("blob", "glass", translation((0, 1.05, 0)))
    sc.add_instance("bowl", "glass", translation((0, 0.001, 0)))
    sc.add_instance("ground", "ground")
    sc.shots.append(Shot((0.0, 2.2, 5.2), (0, 0.9, 0), (0, 1, 0)))
    return sc


def _bowl(nseg, nring):
    """Open lathe bowl (outer + inner wall) around the y axis, ~4*nseg*nring triangles."""
    prof = []
    for k in range(nring + 1):  # outer wall, bottom -> rim
        a = (k / nring) * (math.pi / 2)
        prof.append((1.9 * math.sin(a), 1.9 * (1 - math.cos(a))))
    for k in range(nring, -1, -1):  # inner wall, rim -> bottom
        a = (k / nring) * (math.pi / 2)
        prof.append((1.8 * math.sin(a), 0.1 + 1.8 * (1 - math.cos(a))))
    prof = np.asarray(prof)
    ang = np.linspace(0, 2 * math.pi, nseg, endpoint=False)
    P = np.stack([np.outer(prof[:, 0], np.cos(ang)), np.repeat(prof[:, 1][:, None], nseg, 1),
                  np.outer(prof[:, 0], np.sin(ang))], 2)
    m = len(prof)
    pos = P.reshape(-1, 3)
    idx = []
    for r in range(m - 1):
        for s in range(nseg):
            a, b = r * nseg + s, r * nseg + (s + 1) % nseg
            c, d = a + nseg, b + nseg
            idx += [a, c, d, a, d, b]
    idx = np.asarray(idx, np.int64).reshape(-1, 3)
    fn = np.cross(pos[idx[:, 1]] - pos[idx[:, 0]], pos[idx[:, 2]] - pos[idx[:, 0]])
    keep = np.linalg.norm(fn, axis=1) > 1e-12
    idx, fn = idx[keep], fn[keep]
    n = np.zeros_like(pos)
    for k in range(3):
        np.add.at(n, idx[:, k], fn)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(ln > 0, n / np.maximum(ln, 1e-20), [0, 1, 0])
    uv = np.stack([np.tile(ang / (2 * math.pi), m), np.repeat(np.linspace(0, 1, m), nseg)], 1)
    return make_vertices(pos, uv, n), idx.astype(np.uint32).reshape(-1)


def pbr_spheres(width=1920, height=1080, spp=1024, depth=5, subdiv=7, tex_size=2048):
    """C3: textured pbr_metalness_roughness meshes with sun/sky + a point light."""
    sc = Scene()
    sc.set_camera("perspective", width, height, fov=40.0)
    sc.state["spp"], sc.state["maxPathDepth"] = spp, depth
    sc.set_channels(["diffuse", "normal"])
    ta = sc.add_texture("albedo", noise_texture(tex_size, 99, 3))
    tr = sc.add_texture("rough", noise_texture(tex_size, 100, 1, 0.1, 0.8))
    tm = sc.add_texture("metal", noise_texture(tex_size, 101, 1, 0.0, 1.0))
    nm = noise_texture(tex_size, 102, 3, 0.4, 0.6)
    nm[..., 2] = 1.0
    tn = sc.add_texture("normal", nm)
    sc.add_material("pbr_a", mat(S.MAT_PBR, diffuseTextureId=ta, roughnessTextureId=tr, metalnessTextureId=tm,
                                 normalTextureId=tn, specular=0.0))
    sc.add_material("pbr_b", mat(S.MAT_PBR, diffuse=(.9, .6, .2), roughness=0.3, metalness=1.0, specular=0.0))
    sc.add_material("floor", mat(S.MAT_PBR, diffuse=(.5, .5, .5), roughness=0.6, metalness=0.0, specular=0.0))
    sc.sunsky["in_use"] = 1
    sc.add_light(point_light((0.0, 3.0, 3.0), (6.0, 6.0, 6.5)))
    sc.add_mesh("ball_a", *blob(subdiv, 99, 0.08))
    sc.add_mesh("ball_b", *blob(subdiv, 98, 0.12))
    sc.add_mesh("floor", *grid_plane(16, 8.0, 0.0))
    sc.add_instance("ball_a", "pbr_a", translation((-1.15, 1.05, 0)))
    sc.add_instance("ball_b", "pbr_b", translation((1.15, 1.08, 0)))
    sc.add_instance("floor", "floor")
    sc.shots.append(Shot((0.0, 2.0, 6.0), (0, 1.0, 0), (0, 1, 0)))
    return sc


def instanced_field(width=3840, height=2160, spp=1024, depth=5, subdiv=7, grid=10, seed=42):
    """C4: one blob mesh x grid^2 instances (random yaw/scale), mixed BSDFs, env map + rect light."""
    from .host import COMPLEX_IOR
    sc = Scene()
    sc.set_camera("perspective", width, height, fov=50.0)
    sc.state["spp"], sc.state["maxPathDepth"] = spp, depth
    sc.set_channels(["diffuse", "normal", "position"])
    names = ["lam", "pbr", "plastic", "rplastic", "glass", "gold", "kang"]
    sc.add_material("lam", mat(S.MAT_LAMBERTIAN, diffuse=(.6, .3, .25)))
    sc.add_material("pbr", mat(S.MAT_PBR, diffuse=(.3, .55, .7), roughness=0.4, metalness=0.5, specular=0.0))
    sc.add_material("plastic", mat(S.MAT_PLASTIC, diffuse=(.2, .6, .3), ior=1.5))
    sc.materials[-1]["radiance"][0] = _fdr(1.5)
    sc.add_material("rplastic", mat(S.MAT_ROUGH_PLASTIC, diffuse=(.7, .65, .2), ior=1.5, anisoAlpha=(.2, .2)))
    sc.materials[-1]["radiance"][0] = _fdr(1.5)
    sc.add_material("glass", mat(S.MAT_DIELECTRIC, ior=1.5))
    sc.add_material("gold", mat(S.MAT_CONDUCTOR, diffuse=(1, 1, 1), radiance=COMPLEX_IOR["Au"][0],
                                radianceFactor=COMPLEX_IOR["Au"][1]))
    sc.add_material("kang", mat(S.MAT_KANG18, diffuse=(.35, .3, .5), rhoSpec=(.4, .4, .4), anisoAlpha=(.25, .1)))
    sc.add_material("ground", mat(S.MAT_LAMBERTIAN, diffuse=(.5, .5, .5)))
    sc.set_envmap(procedural_envmap(1024, 512, 11))
    g = float(grid)
    sc.add_light(rect_light((-2, 6.0, -2), (2, 6.0, -2), (-2, 6.0, 2), (30, 28, 25)))
    sc.add_mesh("blob", *blob(subdiv, 1234))
    sc.add_mesh("ground", *grid_plane(8, 2.0 * g, 0.0))
    rng = np.random.RandomState(seed)
    k = 0
    for i in range(grid):
        for j in range(grid):
            s = 0.55 + 0.35 * rng.rand()
            x = (i - (grid - 1) / 2) * 2.2 + 0.5 * (rng.rand() - 0.5)
            z = (j - (grid - 1) / 2) * 2.2 + 0.5 * (rng.rand() - 0.5)
            xf = translation((x, 1.25 * s, z)) @ rotation_y(rng.rand() * 2 * math.pi) @ scaling((s, s, s))
            sc.add_instance("blob", names[k % len(names)], xf)
            k += 1
    sc.add_instance("ground", "ground")
    sc.shots.append(Shot((0.0, 9.0, 1.6 * g + 4), (0, 0.5, 0), (0, 1, 0)))
    return sc


def ray_bench(width=1920, height=1080, subdiv=8, depth=4, spp=16):
    """C4': ~1.3 M-triangle single mesh, lambertian, uniform white environment: closest-hit rays at
    depth >= 2 are cosine-hemisphere bounces (the 'incoherent rays' of the headline metric)."""
    sc = Scene()
    sc.set_camera("perspective", width, height, fov=45.0)
    sc.state["spp"], sc.state["maxPathDepth"] = spp, depth
    sc.add_material("lam", mat(S.MAT_LAMBERTIAN, diffuse=(.7, .7, .7)))
    sc.set_envmap(np.ones((8, 16, 4), f32))
    sc.add_mesh("blob", *blob(subdiv, 1234, 0.25))
    sc.add_mesh("ground", *grid_plane(4, 6.0, -1.3))
    sc.add_instance("blob", "lam")
    sc.add_instance("ground", "lam")
    sc.shots.append(Shot((0.0, 1.2, 3.6), (0, 0, 0), (0, 1, 0)))
    return sc


def orbit_shots(sc, n, center, radius, height, seed=5):
    """C5: n look-at shots on a seeded orbit (poses only; camera type is per scene, loader.cpp:242-268)."""
    rng = np.random.RandomState(seed)
    sc.shots = []
    for k in range(n):
        a = 2 * math.pi * (k + 0.3 * rng.rand()) / n
        eye = (center[0] + radius * math.cos(a), height + 0.2 * rng.rand(), center[2] + radius * math.sin(a))
        sc.shots.append(Shot(eye, center, (0, 1, 0)))
    return sc
