"""C4' ray-throughput check: ~1.3 M-triangle mesh, lambertian, white environment, 1080p.
Incoherent rays = closest-hit rays at depth >= 2 (after a cosine-hemisphere bounce).
usage: python tools/ray_bench.py [subdiv=8] [frames=16] [scene=rays|glass|field|pbr|cornell]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asuna_b200 import capi, scenes

subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 16
which = sys.argv[3] if len(sys.argv) > 3 else "rays"
if which == "rays":
    sc = scenes.ray_bench(1920, 1080, subdiv=subdiv, depth=4)
elif which == "rays_merged":  # same geometry as one mesh / one instance: what a flattened scene would cost
    from asuna_b200 import host
    import numpy as np
    sc = scenes.ray_bench(1920, 1080, subdiv=subdiv, depth=4)
    (v0, i0), (v1, i1) = sc.meshes[0], sc.meshes[1]
    sc.meshes = [(np.concatenate([v0, v1]), np.concatenate([i0, i1 + len(v0)]))]
    sc.mesh_ids = {"blob": 0}
    sc.instances = sc.instances[:1]
elif which == "pbr":  # C3: textured PBR spheres, sun/sky + point light
    sc = scenes.pbr_spheres(1920, 1080, depth=5, subdiv=subdiv)
elif which == "cornell":  # C1: 36 triangles, 512 x 512, depth 5
    sc = scenes.cornell(512, 512, spp=64, depth=5)
elif which == "glass":
    sc = scenes.glass_blob(1920, 1080, subdiv=subdiv, env_size=(2048, 1024))
else:
    sc = scenes.instanced_field(1920, 1080, subdiv=subdiv, grid=10)
ctx = capi.Context(gpu_id=0)
ctx.set_profiling(True)
build_ms = sc.upload(ctx)
acc = ctx.accel_stats()
sc.begin_shot(ctx, 0)
ctx.set_counting(True)
ctx.render_frames(1)
s = ctx.stats()
npr, tpr = s["node_visits"] / s["closest_rays"], s["tri_tests"] / s["closest_rays"]
import ctypes as _C
_ls = (_C.c_uint64 * 6)()
lane = None
if hasattr(ctx.L.lib, "asuna_debug_lane_stats"):
    ctx.L.lib.asuna_debug_lane_stats(ctx.h, _ls)
    it, act, node, want, fired, tri = [float(x) for x in _ls]
    if it:
        lane = {"iters_per_32_rays": it / (s["closest_rays"] / 32), "live_lanes": act / it, "node_step_lanes": node / it,
                "tri_wanting_lanes": want / it, "tri_steps_per_iter": fired / it, "tri_step_lanes": tri / max(fired, 1)}
ctx.set_counting(False)
sc.begin_shot(ctx, 0)
ctx.render_frames(8)
ctx.sync()
ctx.reset_stats()
ctx.render_frames(frames)
ctx.sync()
s = ctx.stats()
print(json.dumps({
    "scene": which, "triangles": acc["leaf_prims"], "wide_nodes": acc["nodes"], "sah": acc["sah_cost"], "build_ms": build_ms,
    "nodes_per_ray": npr, "tris_per_ray": tpr, "lane_stats": lane,
    "incoherent_closest_Mrays_s": s["incoherent_closest_rays"] / s["closest_ms"] / 1e3 * (s["closest_rays"] / max(s["closest_rays"], 1)),
    "closest_Mrays_s": s["closest_rays"] / s["closest_ms"] / 1e3,
    "all_Mrays_s": (s["closest_rays"] + s["shadow_rays"]) / s["trace_ms"] / 1e3,
    "shadow_Mrays_s": s["shadow_rays"] / max(s["shadow_ms"], 1e-9) / 1e3,
    "samples_per_s_M": s["paths"] / s["total_ms"] / 1e3,
    "ms": {k: s[k] for k in ("closest_ms", "shadow_ms", "shade_ms", "total_ms")}}))
