"""Generates tests/golden/* from the CPU oracle (the reference ships no fixtures).

Run from the repo root:  python tools/make_golden.py
The vectors pin the oracle against drift and give the GPU tests a fixed target that does not
need the oracle to be rebuilt on the GPU box.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asuna_b200 import scenes, structs as S  # noqa: E402
from oracle.binding import OracleContext, library  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
L = library().lib

# ---- RNG streams (reference src/shaders/utils/math.glsl:20-44)
L.oracle_xxhash32.restype = C.c_uint32
L.oracle_pcg.restype = C.c_uint32
L.oracle_rand.restype = C.c_float
rng = {"xxhash32": [], "pcg": [], "rand": []}
for x, y, z in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (511, 511, 63), (1919, 1079, 1023), (123456, 654321, 4294967295)]:
    rng["xxhash32"].append([x, y, z, int(L.oracle_xxhash32(C.c_uint32(x), C.c_uint32(y), C.c_uint32(z)))])
for seed in (0, 1, 0xDEADBEEF, 0xFFFFFFFF):
    st = C.c_uint32(seed)
    rng["pcg"].append([seed, [int(L.oracle_pcg(C.byref(st))) for _ in range(8)]])
    st = C.c_uint32(seed)
    rng["rand"].append([seed, [float(L.oracle_rand(C.byref(st))) for _ in range(8)]])
json.dump(rng, open(os.path.join(OUT, "rng.json"), "w"), indent=1)

# ---- sun & sky samples (reference src/shaders/utils/sun_and_sky.glsl:405-533)
ss = S.default_sunsky()
ss["in_use"] = 1
dirs = []
r = np.random.RandomState(11)
for _ in range(64):
    d = r.normal(size=3)
    dirs.append((d / np.linalg.norm(d)).astype(np.float32))
sun = np.array(ss["sun_direction"], np.float32)
sun /= np.linalg.norm(sun)
for eps in (0.0, 0.01, 0.03):
    d = sun + np.array([eps, 0, 0], np.float32)
    dirs.append((d / np.linalg.norm(d)).astype(np.float32))
dirs = np.array(dirs, np.float32)
out = np.zeros_like(dirs)
for i, d in enumerate(dirs):
    o = (C.c_float * 3)()
    L.oracle_sun_and_sky(ss.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), o)
    out[i] = list(o)
np.savez(os.path.join(OUT, "sunsky.npz"), dirs=dirs, radiance=out)


# ---- small renders
def render(sc, name):
    ctx = OracleContext()
    sc.upload(ctx)
    imgs = sc.render_shot(ctx, 0)
    sc.begin_shot(ctx, 0)
    ids, t = ctx.trace_primary()
    st = ctx.stats()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), radiance=imgs[0], aov=np.array(imgs[1:]), ids=ids, t=t,
                        closest_rays=st["closest_rays"], shadow_rays_nonzero=ctx.traversal_counters()["shadow_rays_nonzero"])
    print(name, imgs[0][..., :3].mean(), st)


render(scenes.cornell(48, 48, spp=4, depth=5), "cornell_48_spp4")
render(scenes.cornell_materials(48, 36, spp=4, depth=5, env=False, lights="all", textured=True), "materials_48x36_spp4")
render(scenes.cornell_materials(48, 36, spp=4, depth=5, env=True, lights="rect", textured=True), "materials_env_48x36_spp4")
render(scenes.pbr_spheres(48, 27, spp=4, depth=4, subdiv=3, tex_size=32), "pbr_sunsky_48x27_spp4")
render(scenes.cornell_all_materials(48, 36, spp=4, depth=5, env=True, lights="all", textured=True), "all_materials_48x36_spp4")
