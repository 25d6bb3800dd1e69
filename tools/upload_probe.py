"""Where the serial part of a job goes: times every call of Scene.upload (context creation, film, textures, env
tables, meshes, instances, build) and the first frame (path-state allocation) for the benched scene.
usage: python tools/upload_probe.py [glass|field|rays]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asuna_b200 import capi, scenes, structs as S
from asuna_b200.host import colmajor

which = sys.argv[1] if len(sys.argv) > 1 else "glass"
sc = {"glass": lambda: scenes.glass_blob(1920, 1080, subdiv=6, env_size=(2048, 1024)),
      "field": lambda: scenes.instanced_field(3840, 2160, subdiv=7, grid=10),
      "rays": lambda: scenes.ray_bench(1920, 1080, subdiv=8, depth=4)}[which]()
T = {}


def timed(name, f):
    t0 = time.perf_counter()
    r = f()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    return r


t_all = time.perf_counter()
ctx = timed("create", lambda: capi.Context(gpu_id=0))
timed("set_film", lambda: ctx.set_film(sc.camera["width"], sc.camera["height"]))
for t in sc.textures:
    timed("textures", lambda: ctx.add_texture(t))
for m in sc.materials:
    timed("materials", lambda: ctx.add_material(m))
timed("lights", lambda: ctx.set_lights(np.array(sc.lights, S.Light)))
if sc.envmap is not None:
    timed("envmap", lambda: ctx.set_envmap(*sc.envmap))
for v, i in sc.meshes:
    timed("meshes", lambda: ctx.add_mesh(v, i))
for x, mesh, mat, light in sc.instances:
    timed("instances", lambda: ctx.add_instance(colmajor(x), mesh, mat, light))
timed("sunsky", lambda: ctx.set_sunsky(sc.sunsky))
build_ms = timed("build_accel", lambda: ctx.build_accel())
timed("begin_shot", lambda: sc.begin_shot(ctx, 0))
timed("first_frames(8)+sync", lambda: (ctx.render_frames(8), ctx.sync()))
timed("second_frames(8)+sync", lambda: (ctx.render_frames(8), ctx.sync()))
img = timed("pinned_alloc", lambda: ctx.pinned_image())
timed("read_channel", lambda: ctx.read_channel(0, out=img))
T["total"] = time.perf_counter() - t_all
print(json.dumps({"scene": which, "build_ms_device": build_ms, "seconds": {k: round(v, 4) for k, v in T.items()},
                  "bytes": {"env": 0 if sc.envmap is None else int(sum(a.nbytes for a in sc.envmap)),
                            "textures": int(sum(t.nbytes for t in sc.textures)),
                            "meshes": int(sum(v.nbytes + i.nbytes for v, i in sc.meshes))}}))
