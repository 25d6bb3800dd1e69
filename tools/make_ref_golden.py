"""Generates tests/golden/ref_*.npz from oracle/_ref/libref.so -- the reference's own GLSL compiled as C++
(oracle/refbuild/build_ref.py).  Needs /root/reference (this container); the fixtures travel to the GPU box.

Run from the repo root:  python tools/make_ref_golden.py
  ref_frame_<scene>.npz : radiance, AOVs, primary-hit ids rendered by the reference GLSL
  ref_probes.npz        : 256 single shader invocations per material type, the emitter hit group and the miss
                          shader -- inputs are regenerated from the seeds below, outputs are the reference's
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from asuna_b200 import scenes, structs as S  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

FRAMES = {
    "cornell": lambda: scenes.cornell(48, 48, spp=4, depth=5),
    "materials": lambda: scenes.cornell_materials(48, 36, spp=4, depth=5, env=False, lights="all", textured=True),
    "materials_env": lambda: scenes.cornell_materials(48, 36, spp=4, depth=5, env=True, lights="rect", textured=True),
    "pbr_sunsky": lambda: scenes.pbr_spheres(48, 27, spp=4, depth=4, subdiv=3, tex_size=32),
    "all_materials": lambda: scenes.cornell_all_materials(48, 36, spp=4, depth=5, env=True, lights="all", textured=True),
}


def probe_cases(n=256):
    """(key, scene, probe arguments) -- deterministic, shared by the generator and tests/test_ref_pins.py."""
    import helpers as H
    from test_ref_pins import MATERIAL_NAMES, probe_scene
    for mtype in sorted(MATERIAL_NAMES):
        rng = np.random.RandomState(1000 + mtype)
        sc = probe_scene(rng, mtype, env=bool(mtype % 2), sunsky=(mtype % 3 == 0 and not mtype % 2))
        inst = len(sc.instances) - 1
        yield MATERIAL_NAMES[mtype], sc, H.random_probes(rng, n, len(sc.meshes[sc.instances[inst][1]][1]) // 3, inst)
    rng = np.random.RandomState(2000)
    sc = probe_scene(rng, S.MAT_LAMBERTIAN, env=True)
    yield "emitter_rect", sc, H.random_probes(rng, n, 2, 0)
    yield "miss_envmap", sc, H.random_probes(rng, n, 1, H.MISS)
    sc2 = probe_scene(rng, S.MAT_LAMBERTIAN, env=False, sunsky=True)
    yield "miss_sunsky", sc2, H.random_probes(rng, n, 1, H.MISS)


def main():
    import helpers as H
    from oracle.binding import RefContext
    os.makedirs(OUT, exist_ok=True)
    for name, make in FRAMES.items():
        sc, ctx = make(), RefContext()
        sc.upload(ctx)
        imgs = sc.render_shot(ctx, 0)
        sc.begin_shot(ctx, 0)
        ids, t = ctx.trace_primary()
        st = ctx.stats()
        np.savez_compressed(os.path.join(OUT, f"ref_frame_{name}.npz"), radiance=imgs[0], aov=np.array(imgs[1:]), ids=ids,
                            t=t, closest_rays=st["closest_rays"], shadow_rays=st["shadow_rays"])
        print(name, float(imgs[0][..., :3].mean()), st["closest_rays"], st["shadow_rays"])
        ctx.close()
    out = {}
    for key, sc, args in probe_cases():
        ctx = RefContext()
        sc.upload(ctx)
        sc.begin_shot(ctx, 0)
        out[key] = H.run_probes(ctx, *args).view(np.uint8)
        ctx.close()
    np.savez_compressed(os.path.join(OUT, "ref_probes.npz"), **out)
    print("probes", {k: v.size for k, v in out.items()})


if __name__ == "__main__":
    main()
