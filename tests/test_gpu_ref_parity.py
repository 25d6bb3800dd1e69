"""The CUDA path against the REFERENCE'S OWN SHADERS (oracle/_ref/libref.so: the reference GLSL compiled as C++,
oracle/refbuild/build_ref.py -- prebuilt where /root/reference exists, it travels to the GPU box) and against the
frames frozen from it (tests/golden/ref_frame_*.npz), with the BASELINE.json thresholds; then the same
comparison at BASELINE size on the benched scene itself and on the other full-size configurations.
Nothing here reads /root/reference at run time."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from asuna_b200 import metrics, scenes, structs as S
from test_gpu_parity import SCENES, silhouette_mask

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ref_ctx():
    """libref.so when it is there (kind 'reference'), else the hand restatement it is pinned to."""
    from oracle import binding
    ctx = binding.RefContext() if os.path.exists(binding.REF_LIB) else binding.OracleContext()
    yield ctx
    ctx.close()


def check(sc, gpu, cpu, shot=0, spp=None, aov_tol=1e-4, aov_outliers=0.0):
    """BASELINE.json: primary ids >= 99.9 %, AOVs <= 1e-4 off silhouettes, radiance mean relative error <= 1 %,
    FLIP <= 0.01 at equal spp.  aov_outliers: fraction of interior pixels allowed between aov_tol and 10 x aov_tol
    (full-size textured scenes only: 2048^2 normal maps under a 4x uv repeat turn 1 ulp of uv into 5e-4 texel)."""
    sc.upload(gpu), sc.upload(cpu)
    sc.begin_shot(gpu, shot), sc.begin_shot(cpu, shot)
    ig, tg = gpu.trace_primary()
    ic, tc = cpu.trace_primary()
    same = (ig == ic).all(axis=2)
    assert same.mean() >= 0.999, f"primary ids agree on {same.mean():.5f}"
    g, c = sc.render_shot(gpu, shot, spp), sc.render_shot(cpu, shot, spp)
    interior = same & ~silhouette_mask(ic)
    for k in range(1, len(g)):
        d = np.abs(g[k][..., :3] - c[k][..., :3]).max(axis=2)
        scale = max(1.0, float(np.abs(c[k][..., :3]).max()))
        over = float((d[interior] > aov_tol * scale).mean())
        assert over <= aov_outliers and d[interior].max() <= 10 * aov_tol * scale, f"AOV {k}: max {d[interior].max()}, {over:.2e} of pixels over {aov_tol}"
    rel, fl = metrics.mean_relative_error(g[0], c[0]), metrics.flip(g[0], c[0])
    assert rel <= 0.01 and fl <= 0.01, (rel, fl)
    return rel, fl, float(same.mean())


@pytest.mark.parametrize("name", list(SCENES))
def test_gpu_vs_reference_glsl(name, gpu_ctx, ref_ctx):
    check(SCENES[name](), gpu_ctx, ref_ctx)


@pytest.mark.parametrize("name", ["cornell", "materials", "materials_env", "pbr_sunsky", "all_materials"])
def test_gpu_vs_frozen_reference_glsl_frames(name, gpu_ctx):
    from tools.make_ref_golden import FRAMES
    g = np.load(os.path.join(GOLDEN, f"ref_frame_{name}.npz"))
    sc = FRAMES[name]()
    sc.upload(gpu_ctx)
    sc.begin_shot(gpu_ctx, 0)
    ids, _ = gpu_ctx.trace_primary()
    same = (ids == g["ids"]).all(axis=2)
    assert same.mean() >= 0.999
    imgs = sc.render_shot(gpu_ctx, 0)
    interior = same & ~silhouette_mask(g["ids"])
    for k, aov in enumerate(g["aov"]):
        d = np.abs(imgs[1 + k][..., :3] - aov[..., :3]).max(axis=2)
        assert d[interior].max() <= 1e-4
    assert metrics.mean_relative_error(imgs[0], g["radiance"]) <= 0.01 and metrics.flip(imgs[0], g["radiance"]) <= 0.01


# ----------------------------------------------------------------------------- BASELINE size
def test_full_size_benched_scene(gpu_ctx, ref_ctx):
    """BASELINE configs[1] exactly as bench.py renders it: glass blob, 1920x1080, depth 8, subdiv 6 (114 816
    triangles), 2048x1024 env map -- 2 spp against the reference shaders on the CPU (~2.1 M paths)."""
    sc = scenes.glass_blob(1920, 1080, spp=2, depth=8, subdiv=6, env_size=(2048, 1024))
    rel, fl, agree = check(sc, gpu_ctx, ref_ctx)
    print(f"configs[1] 1080p: primary ids {agree:.5f}, mean rel {rel:.2e}, FLIP {fl:.2e}")


def test_full_size_cornell_config0(gpu_ctx, ref_ctx):
    """BASELINE configs[0]: Cornell box 512x512, depth 5 -- 8 of the 64 spp (equal spp on both sides)."""
    check(scenes.cornell(512, 512, spp=8, depth=5), gpu_ctx, ref_ctx)


def test_full_size_pbr_config2(gpu_ctx, ref_ctx):
    """BASELINE configs[2]: two displaced spheres (655 872 triangles), 2048^2 albedo / roughness / metalness / normal
    textures, sun & sky + point light, 1080p, depth 5 -- 1 spp (AOVs, ids) + 1 jittered frame."""
    sc = scenes.pbr_spheres(1920, 1080, spp=2, depth=5, subdiv=7, tex_size=2048)
    check(sc, gpu_ctx, ref_ctx, aov_outliers=1e-5)


def test_million_triangle_ray_bench_config(gpu_ctx, ref_ctx):
    """C4' (SURVEY.md 8d): the 1.31 M-triangle ray-bench mesh, primary ids + position AOV at a reduced film."""
    sc = scenes.ray_bench(480, 270, subdiv=8, depth=4, spp=2)
    check(sc, gpu_ctx, ref_ctx)


def test_instanced_field_config3_reduced_film(gpu_ctx, ref_ctx):
    """BASELINE configs[3]: 100 instances of a 327 k-triangle mesh (32.8 M instanced triangles, flattened on the GPU,
    two-level on the CPU), mixed BSDFs, env + rect light -- full geometry, film reduced to 480x270, 2 spp."""
    sc = scenes.instanced_field(480, 270, spp=2, depth=5, subdiv=7, grid=10)
    check(sc, gpu_ctx, ref_ctx)


def test_multi_shot_sweep_config4(gpu_ctx, ref_ctx):
    """BASELINE configs[4] at 1080p: the 16-shot thin-lens orbit of tools/run_configs.py (three of its shots, 2 spp each,
    one with a per-shot spp / depth override) and an opencv pinhole shot, radiance + albedo / normal / depth AOVs against
    the reference shaders.  The thin lens draws two more random numbers per sample (rgen:64-100): ids are compared on the
    pinhole shot, where frame 0 has no jitter."""
    sc = scenes.cornell(1920, 1080, spp=2, depth=5)
    sc.camera["aperture"], sc.camera["focal_distance"] = 0.05, 3.0
    scenes.orbit_shots(sc, 16, (0.5, 0.5, 0.5), 2.2, 0.5)
    sc.shots[11].state = sc.state.copy()
    sc.shots[11].state["spp"], sc.shots[11].state["maxPathDepth"] = 3, 2
    sc.upload(gpu_ctx), sc.upload(ref_ctx)
    for shot in (0, 5, 11):
        g, c = sc.render_shot(gpu_ctx, shot), sc.render_shot(ref_ctx, shot)
        assert metrics.mean_relative_error(g[0], c[0]) <= 0.01 and metrics.flip(g[0], c[0]) <= 0.01, shot
        for k in range(1, len(g)):  # AOVs of a lens-sampled primary ray: a few pixels straddle an edge differently
            d = np.abs(g[k][..., :3] - c[k][..., :3]).max(axis=2)
            assert (d > 1e-4 * max(1.0, float(np.abs(c[k][..., :3]).max()))).mean() <= 1e-3, (shot, k)
    so = scenes.cornell(1920, 1080, spp=2, depth=5)
    so.set_camera("opencv", 1920, 1080, fxfycxcy=[1700.0, 1700.0, 960.0, 540.0])
    check(so, gpu_ctx, ref_ctx)


def test_frame_batches_beyond_one_internal_batch(gpu_ctx, ref_ctx):
    """24 frames of a 320x180 film = 3 internal batches of 8: the accumulation across batches against 24 reference
    frames (rgen:171-178)."""
    sc = scenes.cornell_materials(320, 180, spp=24, depth=5, env=True, lights="rect", textured=True)
    check(sc, gpu_ctx, ref_ctx)


def test_opacity_pass_through_scene(gpu_ctx, ref_ctx):
    """pbr / kang18 / disney opacity: `rand < opacity` continues the ray through the surface with depth-- (the host
    loop of render_batch re-launches until the queue drains); constant and textured opacity."""
    sc = scenes.cornell_all_materials(160, 120, spp=8, depth=5, env=False, lights="all", textured=True)
    for m in sc.materials:
        if int(m["type"]) == 3:
            m["specular"] = 0.35
        elif int(m["type"]) == 1:
            m["metalness"] = 0.35
    check(sc, gpu_ctx, ref_ctx)


def test_post_process_on_gpu_vs_reference_glsl(gpu_ctx, ref_ctx):
    """asuna_post_process (k_post_process) against post.idle.frag compiled as C++ (or the restatement pinned to it), all
    seven tone mappers incl. the custom one with grading and global auto-exposure.  The CPU side tone-maps the GPU's own
    HDR image, so only the post stage is compared."""
    from test_ref_pins import post_agreement, post_cases
    sc = scenes.cornell_materials(96, 72, spp=4, env=True, lights="rect", textured=True)
    sc.upload(gpu_ctx), sc.upload(ref_ctx)
    hdr = sc.render_shot(gpu_ctx, 0)[0]
    sc.render_shot(ref_ctx, 0, spp=1)
    # hand the GPU's radiance image to the CPU path: export/import round trip of (L * 1, 1)
    import ctypes as C
    ptr = C.cast(ref_ctx.export_partial(), C.POINTER(C.c_float))
    part = np.ctypeslib.as_array(ptr, shape=hdr.shape)
    part[..., :3], part[..., 3] = hdr[..., :3], 1.0
    ref_ctx.import_partial()
    assert np.array_equal(ref_ctx.read_channel(0)[..., :3], hdr[..., :3])
    for name, tm in post_cases():
        g, c = gpu_ctx.post_process(tm), ref_ctx.post_process(tm)
        frac, worst = post_agreement(g, c)
        custom = name.startswith("custom")
        assert frac <= (0.01 if custom else 0.0) and worst <= (1.5 / 255 if custom else 2e-5), (name, frac, worst)
    bad = S.default_post("custom")
    bad["autoExposure"] = 3
    with pytest.raises(Exception):
        gpu_ctx.post_process(bad)


# ----------------------------------------------------------------------------- one shade-kernel invocation at a time
def gpu_probes(ctx, inst, prim, b1, b2, probes):
    """asuna_debug_shade_probes: the payloads go through one regroup + k_shade<kind> pass of the CUDA path."""
    import ctypes as C
    q = probes.copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = ctx.L.lib.asuna_debug_shade_probes(ctx.h, C.c_uint32(len(q)), p(inst), p(prim), p(b1), p(b2), p(q))
    assert rc == 0, ctx.L.fn("last_error", C.c_char_p)(ctx.h)
    return q


def check_gpu_probes(G, R, name):
    """G: CUDA path, R: reference GLSL (live or frozen).  Integer outputs identical: stop, and for continuing paths the
    RNG state after the shader (= number and order of draws) and the depth; lobe flags on >= 99.5 %.  Float outputs within
    2e-4 of the vector's magnitude on >= 99.5 % of probes.  A path that stopped keeps no next ray / RNG state in the
    wavefront form, and a zero NEE contribution is never queued (A.3-4): those fields are compared where they exist."""
    assert np.array_equal(G["stop"], R["stop"]), (name, float((G["stop"] != R["stop"]).mean()))
    cont = R["stop"] == 0
    assert np.array_equal(G["seed"][cont], R["seed"][cont]), (name, "rng state")
    assert np.array_equal(G["depth"][cont], R["depth"][cont]), (name, "depth")
    assert (G["brec_flags"][cont] != R["brec_flags"][cont]).mean() <= 0.005 if cont.any() else True

    def frac_bad(a, b, sel, tol=2e-4):
        if not sel.any():
            return 0.0
        a, b = a[sel].reshape(int(sel.sum()), -1).astype(np.float64), b[sel].reshape(int(sel.sum()), -1).astype(np.float64)
        fa, fb = np.isfinite(a).all(axis=1), np.isfinite(b).all(axis=1)
        with np.errstate(invalid="ignore"):
            d = np.linalg.norm(np.where(np.isfinite(a - b), a - b, 0), axis=1)
            sc = np.maximum(1.0, np.linalg.norm(np.where(np.isfinite(b), b, 0), axis=1))
        return float(((fa != fb) | (fa & fb & (d > tol * sc))).mean())

    bad = {f: frac_bad(G[f], R[f], cont) for f in ("ray_o", "ray_d", "throughput", "brec_pdf")}
    bad["radiance"] = frac_bad(G["radiance"], R["radiance"], np.ones(len(R), bool))
    bad["channel"] = frac_bad(G["channel"][:, :7], R["channel"][:, :7], np.ones(len(R), bool), 1e-4)
    ref_skip = (R["drec_skip"] == 1) | (R["drec_radiance"] == 0).all(axis=1)
    bad["drec_skip"] = float(((G["drec_skip"] == 1) != ref_skip).mean())
    both = (G["drec_skip"] == 0) & ~ref_skip
    for f in ("drec_radiance", "drec_o", "drec_d", "drec_dist"):
        bad[f] = frac_bad(G[f], R[f], both)
    assert max(bad.values()) <= 0.005, (name, bad)
    return bad


def test_gpu_shade_kernels_vs_frozen_reference_glsl_probes(product_lib):
    """Every shade kernel (twelve materials, emitter hit, miss with env map / sun & sky) one invocation at a time against
    tests/golden/ref_probes.npz -- outputs of the reference's own closest-hit / miss shaders (libref.so), 256 per kind."""
    import helpers as H
    from asuna_b200 import capi
    from tools.make_ref_golden import probe_cases
    g = np.load(os.path.join(GOLDEN, "ref_probes.npz"))
    for key, sc, args in probe_cases():
        ctx = capi.Context(product_lib, 0)
        sc.upload(ctx)
        sc.begin_shot(ctx, 0)
        G = gpu_probes(ctx, *args)
        check_gpu_probes(G, g[key].view(H.PROBE).reshape(-1), key)
        ctx.close()


@pytest.mark.parametrize("mtype", list(range(12)))
def test_gpu_shade_kernels_vs_live_reference_glsl_probes(mtype, product_lib):
    """The same with 3 random materials x {lights, + env map, + sun/sky} x 1500 random hits per material type against
    libref.so on the box (skipped where the prebuilt library did not travel)."""
    import helpers as H
    from asuna_b200 import capi
    from oracle import binding
    from test_ref_pins import MATERIAL_NAMES, probe_scene
    if not os.path.exists(binding.REF_LIB):
        pytest.skip("oracle/_ref/libref.so not present")
    rng = np.random.RandomState(500 + mtype)
    for rnd in range(3):
        for env, sunsky in ((False, False), (True, False), (False, True)):
            sc = probe_scene(rng, mtype, env, sunsky)
            sc.camera["width"], sc.camera["height"] = 64, 32
            gpu, ref = capi.Context(product_lib, 0), binding.RefContext()
            sc.upload(gpu), sc.upload(ref)
            sc.begin_shot(gpu, 0), sc.begin_shot(ref, 0)
            inst = len(sc.instances) - 1
            args = H.random_probes(rng, 1500, len(sc.meshes[sc.instances[inst][1]][1]) // 3, inst)
            check_gpu_probes(gpu_probes(gpu, *args), H.run_probes(ref, *args), f"{MATERIAL_NAMES[mtype]} round {rnd} env {env} sunsky {sunsky}")
            gpu.close(), ref.close()
