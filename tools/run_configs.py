"""Renders every BASELINE.json config at its full resolution on one GPU for a bounded number of frames and prints one
JSON line per config: samples/s, rays/s (all / incoherent closest), BVH build ms, traversal statistics.
(bench.py measures configs[1] under the full contract; this is the survey table for DESIGN.md.)
usage: python tools/run_configs.py [frames_per_config=16]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asuna_b200 import capi, scenes

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16


def c5():
    sc = scenes.cornell(1920, 1080, spp=16, depth=5)
    sc.camera["aperture"], sc.camera["focal_distance"] = 0.05, 3.0
    scenes.orbit_shots(sc, 16, (0.5, 0.5, 0.5), 2.2, 0.5)
    return sc


CONFIGS = [
    ("C1 cornell 512x512 depth 5", lambda: scenes.cornell(512, 512, spp=64, depth=5)),
    ("C2 glass blob 1080p depth 8", lambda: scenes.glass_blob(1920, 1080, spp=256, depth=8, subdiv=6)),
    ("C3 PBR textured + sun/sky + point light 1080p depth 5", lambda: scenes.pbr_spheres(1920, 1080, spp=1024, depth=5)),
    ("C4 10M-triangle instanced field 4K depth 5", lambda: scenes.instanced_field(3840, 2160, spp=1024, depth=5, subdiv=7, grid=10)),
    ("C4' 1.3M-triangle ray bench 1080p depth 4", lambda: scenes.ray_bench(1920, 1080, subdiv=8, depth=4)),
    ("C5 multi-shot sweep (16 shots, thin lens) 1080p", c5),
]
for name, build in CONFIGS:
    t0 = time.perf_counter()
    sc = build()
    gen_s = time.perf_counter() - t0
    ctx = capi.Context(gpu_id=0)
    ctx.set_profiling(True)
    build_ms = sc.upload(ctx)
    acc = ctx.accel_stats()
    shots = range(len(sc.shots)) if name.startswith("C5") else [0]
    per_shot = frames if len(shots) == 1 else max(frames // 4, 2)
    sc.begin_shot(ctx, 0)
    ctx.render_frames(per_shot)  # warm-up with the timed batch shape (path buffers are sized on first use)
    ctx.sync()
    ctx.reset_stats()
    t0 = time.perf_counter()
    for s in shots:
        sc.begin_shot(ctx, s)
        ctx.render_frames(per_shot)
    ctx.sync()
    wall = time.perf_counter() - t0
    st = ctx.stats()
    inst_tris = sum(len(sc.meshes[m][1]) // 3 for _, m, _, _ in sc.instances)
    print(json.dumps({
        "config": name, "instanced_triangles": inst_tris, "unique_triangles": acc["leaf_prims"], "wide_nodes": acc["nodes"],
        "bvh_build_ms": round(build_ms, 2), "samples_per_s_M": round(st["paths"] / wall / 1e6, 1),
        "rays_per_sample": round((st["closest_rays"] + st["shadow_rays"]) / max(st["paths"], 1), 2),
        "all_Mrays_s": round((st["closest_rays"] + st["shadow_rays"]) / max(st["trace_ms"], 1e-9) / 1e3),
        "incoherent_closest_Mrays_s": round(st["incoherent_closest_rays"] / max(st["closest_ms"], 1e-9) / 1e3),
        "ms": {k: round(st[k], 1) for k in ("closest_ms", "shadow_ms", "shade_ms", "total_ms")},
        "stack_overflow": st.get("stack_overflow", 0), "scene_gen_s": round(gen_s, 1)}), flush=True)
    ctx.close()
