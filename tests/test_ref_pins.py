"""Pins the CPU oracle (the hand restatement, oracle/oracle.cpp) to oracle/_ref/libref.so: the reference's OWN
GLSL -- raytrace.projective.rgen, default/shadow.rmiss, all twelve bxdf/*.rchit and utils/*.glsl -- compiled as
C++ against the GLM vendored in the reference tree (recipe: oracle/refbuild/build_ref.py; nothing of the
reference is stored in this repository).  Only traceRayEXT's geometric query (which the Vulkan driver answers in
the reference) comes from the oracle's BVH; RNG, camera, hit state, light sampling, BSDFs, MIS, miss, filter and
accumulation arithmetic are the reference's text.

Two layers:
  * live: oracle vs libref.so on >= 10 k random inputs per function / per shader (needs libref.so: built here from
    /root/reference, or prebuilt in oracle/_ref/);
  * frozen: oracle vs tests/golden/ref_*.npz, generated from libref.so by tools/make_ref_golden.py (always runs).

Tolerances: integer outputs (RNG words, RNG state after a shader = number of draws, flags, depth, stop/skip
decisions, offsetPositionAlongNormal's int-ULP arithmetic, primary-hit ids) bit-exact; fp32 outputs within a few
ulp per operation -- GLM's normalize() is v*inversesqrt(dot), its mix() is x+a(y-x), where the restatement follows
the GLSL specification's v/length(v) and x(1-a)+ya -- stated per test.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN
from asuna_b200 import host, metrics, scenes, structs as S
import helpers as H

N = 20000


@pytest.fixture(scope="session")
def ref_lib():
    from oracle import binding
    path = binding.build_ref()
    if path is None:
        pytest.skip("oracle/_ref/libref.so is not built and the reference tree is absent")
    return binding.ref_library()


@pytest.fixture()
def ref_ctx(ref_lib):
    from oracle.binding import RefContext
    ctx = RefContext()
    yield ctx
    ctx.close()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ulp_distance(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def test_libref_is_the_reference_glsl_and_liboracle_is_not(oracle_lib, ref_lib):
    assert C.CDLL(ref_lib.path).oracle_is_reference_glsl() == 1
    assert C.CDLL(oracle_lib.path).oracle_is_reference_glsl() == 0
    gen = os.path.join(os.path.dirname(ref_lib.path), "ref_glsl_gen.cpp")
    if os.path.exists(gen):  # the generated unit names every shader file it was made from
        text = open(gen).read()
        for f in ("raytrace.projective.rgen", "raytrace.default.rmiss", "utils/math.glsl", "utils/sample_light.glsl",
                  "utils/sun_and_sky.glsl", "bxdf/raytrace.brdf_disney.rchit", "bxdf/raytrace.bsdf_dielectric.rchit"):
            assert "======== src/shaders/" + f in text


# ----------------------------------------------------------------------------- integer / bit-exact functions
def test_rng_bit_exact(oracle_lib, ref_lib):
    """xxhash32Seed, pcg, rand (utils/math.glsl:20-44) on 20 k random inputs each: identical words and floats, and
    rand2 draws x before y (GLSL argument order)."""
    O, R = oracle_lib.lib, C.CDLL(ref_lib.path)
    for L, pre in ((O, "oracle_"), (R, "refglsl_")):
        getattr(L, pre + "xxhash32").restype = C.c_uint32
        getattr(L, pre + "pcg").restype = C.c_uint32
        getattr(L, pre + "rand").restype = C.c_float
    rng = np.random.RandomState(5)
    xyz = rng.randint(0, 2 ** 32, (N, 3), dtype=np.uint64)
    xyz[:64] = [[x, y, z] for x in (0, 1, 1919, 2 ** 32 - 1) for y in (0, 1079, 2 ** 32 - 1, 7) for z in (0, 1, 255, 2 ** 32 - 1)]
    for x, y, z in xyz:
        a = O.oracle_xxhash32(C.c_uint32(int(x)), C.c_uint32(int(y)), C.c_uint32(int(z)))
        b = R.refglsl_xxhash32(C.c_uint32(int(x)), C.c_uint32(int(y)), C.c_uint32(int(z)))
        assert a == b
    sa, sb = C.c_uint32(12345), C.c_uint32(12345)
    for _ in range(N):
        assert O.oracle_pcg(C.byref(sa)) == R.refglsl_pcg(C.byref(sb)) and sa.value == sb.value
    ones = 0
    for seed in rng.randint(0, 2 ** 32, N, dtype=np.uint64):
        sa, sb = C.c_uint32(int(seed)), C.c_uint32(int(seed))
        fa, fb = O.oracle_rand(C.byref(sa)), R.refglsl_rand(C.byref(sb))
        assert np.float32(fa).tobytes() == np.float32(fb).tobytes() and sa.value == sb.value
        ones += fa == 1.0
    # rand2: x is the first draw
    st, out = C.c_uint32(99), (C.c_float * 2)()
    R.refglsl_rand2(C.byref(st), out)
    s2 = C.c_uint32(99)
    assert out[0] == R.refglsl_rand(C.byref(s2)) and out[1] == R.refglsl_rand(C.byref(s2)) and st.value == s2.value


def test_offset_position_bit_exact(oracle_lib, ref_lib):
    """offsetPositionAlongNormal (utils/math.glsl:241-266): int-ULP branch, near-origin float branch, negative
    coordinates -- bit-exact on 20 k points."""
    O, R = oracle_lib.lib, C.CDLL(ref_lib.path)
    rng = np.random.RandomState(6)
    p = (rng.randn(N, 3) * rng.choice([1e-3, 0.02, 0.05, 1.0, 50.0], (N, 1))).astype(np.float32)
    n = rng.randn(N, 3).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n[::7] *= -1
    a, b = (C.c_float * 3)(), (C.c_float * 3)()
    for i in range(N):
        O.oracle_offset_position(_p(p[i]), _p(n[i]), a)
        R.refglsl_offset_position(_p(p[i]), _p(n[i]), b)
        assert bytes(a) == bytes(b), (p[i], n[i], list(a), list(b))


def test_texture_bilinear_bit_exact(oracle_lib, ref_lib):
    """texture() through the LINEAR / REPEAT sampler of core/texture.cpp:103-107, incl. wrap-around and negative
    coordinates.  (Both sides state the Vulkan filtering formula with fp32 weights; this checks they state it alike.)"""
    O, R = oracle_lib.lib, C.CDLL(ref_lib.path)
    rng = np.random.RandomState(7)
    tex = rng.rand(5, 7, 4).astype(np.float32)
    uv = (rng.rand(N // 4, 2) * 4 - 2).astype(np.float32)
    uv[:4] = [[0, 0], [1, 1], [0, 0.5], [-1e-8, 0.999999]]
    a, b = (C.c_float * 4)(), (C.c_float * 4)()
    for u, v in uv:
        O.oracle_texture_bilinear(_p(tex), 7, 5, C.c_float(u), C.c_float(v), a)
        R.refglsl_texture_bilinear(_p(tex), 7, 5, C.c_float(u), C.c_float(v), b)
        assert bytes(a) == bytes(b)


# ----------------------------------------------------------------------------- fp32 functions
def _pairs(oracle_lib, ref_lib, name, inputs, nout, in_types):
    O, R = oracle_lib.lib, C.CDLL(ref_lib.path)
    fo, fr = getattr(O, "oracle_" + name), getattr(R, "refglsl_" + name)
    A, B = np.zeros((len(inputs), nout), np.float32), np.zeros((len(inputs), nout), np.float32)
    a, b = (C.c_float * nout)(), (C.c_float * nout)()
    for i, x in enumerate(inputs):
        fo(_p(x), a), fr(_p(x), b)
        A[i], B[i] = list(a), list(b)
    return A, B


def test_sampling_helpers(oracle_lib, ref_lib):
    """concentricSampleDisk, cosineSampleHemisphere, uniformSampleSphere (utils/math.glsl:137-175), basis
    (:199-216), powerHeuristic (:187-192): <= 2 ulp (same libm; only a/b vs a*(1/b) style differences)."""
    rng = np.random.RandomState(8)
    u = rng.rand(N, 2).astype(np.float32)
    u[:5] = [[0.5, 0.5], [0, 0], [1, 1], [0, 1], [1, 0]]
    for name, nout in (("concentric_disk", 2), ("cosine_hemisphere", 3), ("uniform_sphere", 3)):
        A, B = _pairs(oracle_lib, ref_lib, name, u, nout, None)
        assert np.all(np.abs(A - B) <= 2.5e-7), name  # values are O(1): 2 ulp
    O, R = oracle_lib.lib, C.CDLL(ref_lib.path)
    n = rng.randn(N, 3).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n[:3] = [[0, 0, -1], [0, 0, 1], [1e-4, 0, -0.99999994]]
    fa, ra, fb, rb = ((C.c_float * 3)() for _ in range(4))
    for v in n:
        O.oracle_basis(_p(v), fa, ra), R.refglsl_basis(_p(v), fb, rb)
        assert bytes(fa) == bytes(fb) and bytes(ra) == bytes(rb)
    O.oracle_power_heuristic.restype = R.refglsl_power_heuristic.restype = C.c_float
    ab = (rng.rand(N, 2) * rng.choice([0, 1e-3, 1, 1e3], (N, 2))).astype(np.float32)
    for x, y in ab:
        assert O.oracle_power_heuristic(C.c_float(x), C.c_float(y)) == R.refglsl_power_heuristic(C.c_float(x), C.c_float(y))


def test_sun_and_sky(oracle_lib, ref_lib):
    """sun_and_sky (utils/sun_and_sky.glsl:405-533) for 5 settings x 4 k directions, above and below the horizon,
    into the disk and the glow: relative difference <= 2e-5 of the colour's magnitude outside the glow radius."""
    O, R = oracle_lib.lib, C.CDLL(ref_lib.path)
    rng = np.random.RandomState(9)
    a, b = (C.c_float * 3)(), (C.c_float * 3)()
    for k in range(5):
        ss = S.default_sunsky()
        ss["in_use"] = 1
        if k:
            ss["haze"], ss["redblueshift"], ss["saturation"] = rng.uniform(0, 8), rng.uniform(-0.5, 0.5), rng.uniform(0.2, 1.8)
            ss["horizon_height"], ss["horizon_blur"] = rng.uniform(-0.5, 0.5), rng.uniform(0, 1)
            sd = rng.randn(3)
            sd[1] = abs(sd[1]) * (1 if k < 4 else -0.2)  # k = 4: sun below the horizon (night branch)
            ss["sun_direction"] = sd / np.linalg.norm(sd)
            ss["sun_disk_scale"], ss["sun_glow_intensity"] = rng.uniform(0.5, 6), rng.uniform(0, 2)
            ss["physically_scaled_sun"], ss["y_is_up"] = k % 2, 1 if k != 3 else 0
        d = rng.randn(4000, 3).astype(np.float32)
        sun = np.asarray(ss["sun_direction"], np.float32)
        d[:400] = sun + rng.randn(400, 3) * 0.02  # disk and glow
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        worst = 0.0
        glow = 0.00465 * float(ss["sun_disk_scale"]) * 10.0
        for v in d:
            O.oracle_sun_and_sky(_p(ss), _p(v), a), R.refglsl_sun_and_sky(_p(ss), _p(v), b)
            x, y = np.array(list(a), np.float64), np.array(list(b), np.float64)
            assert np.array_equal(np.isfinite(x), np.isfinite(y))
            f = np.isfinite(x)
            if not f.any():
                continue
            err = float(np.abs(x - y)[f].max() / max(np.abs(y[f]).max(), 1e-3))
            if np.arccos(np.clip(np.dot(v, sun), -1, 1)) < 1.2 * glow + 0.01:
                # inside disk + glow the model takes acos(dot) of two nearly parallel unit vectors (one ulp of the
                # dot product moves the angle by up to 1e-3 relative), cubes (1 - angle / radius) and feeds a
                # smoothstep edge: ill-conditioned in fp32, so normalize() alone (GLM: v * inversesqrt(dot(v, v)),
                # restatement: v / length(v)) moves the result.  Measured <= 3.5e-3; bound 1e-2.
                assert err <= 1e-2, (k, v, err)
            else:
                worst = max(worst, err)
        assert worst <= 2e-5, (k, worst)


def test_light_samplers(oracle_lib, ref_lib):
    """sampleOneLight -> rect / triangle (non-uniform, A.3-2) / point (A.3-12) / distant (sample_light.glsl:10-82):
    radiance, direction, normal, distance within 4e-6 relative, pdf within 1e-4, flags equal."""
    O, R = oracle_lib.lib, C.CDLL(ref_lib.path)
    rng = np.random.RandomState(10)
    a, b = (C.c_float * 12)(), (C.c_float * 12)()
    for ltype in (S.LIGHT_RECT, S.LIGHT_TRIANGLE, S.LIGHT_POINT, S.LIGHT_DIRECTIONAL):
        for _ in range(N // 8):
            l = np.zeros((), S.Light)
            l["type"] = ltype
            l["position"], l["direction"], l["radiance"] = rng.randn(3), rng.randn(3), rng.rand(3) * 20
            l["u"], l["v"] = rng.randn(3), rng.randn(3)
            l["area"] = np.linalg.norm(np.cross(l["u"], l["v"])) * (0.5 if ltype == S.LIGHT_TRIANGLE else 1.0)
            r, pos = rng.rand(2).astype(np.float32), rng.randn(3).astype(np.float32)
            O.oracle_sample_one_light(_p(l), _p(r), _p(pos), a), R.refglsl_sample_one_light(_p(l), _p(r), _p(pos), b)
            x, y = np.array(list(a)), np.array(list(b))
            assert x[11] == y[11]
            tol = np.full(12, 4e-6)
            tol[10] = 1e-4  # pdf = d^2 / (A |n.d| + EPS): the cosine of a grazing direction amplifies 1 ulp of n
            assert np.all(np.abs(x - y) <= tol * np.maximum(1.0, np.abs(y))), (ltype, x, y)


def test_envmap_sampling(oracle_lib, ref_lib, cpu_ctx):
    """sampleEnvmap / evalEnvmap / pdfEnvmap (sample_light.glsl:84-129) over the tables of core/texture.cpp:144-226,
    with a rotated envTransform; reproduces the halved marginal fetch (A.3-15) on both sides.  Tolerance 1e-4: a
    1-ulp difference in the lat-long uv moves a bilinear lookup of a 64 x 32 map with sharp features by ~1e-5."""
    O, R = oracle_lib.lib, C.CDLL(ref_lib.path)
    rng = np.random.RandomState(11)
    env = scenes.procedural_envmap(64, 32, 3)
    marg, cond = host.envmap_tables(env)
    sc = scenes.cornell(8, 8, spp=1)
    sc.set_envmap(env)
    sc.state["envMapIntensity"] = 1.7
    sc.shots[0].env_transform = host.rotation_y(0.6) @ host.rotation_x(0.3)
    sc.upload(cpu_ctx)
    sc.begin_shot(cpu_ctx, 0)
    xf = np.ascontiguousarray(host.colmajor(sc.shots[0].env_transform), np.float32)
    a7, b7, a4, b4 = (C.c_float * 7)(), (C.c_float * 7)(), (C.c_float * 4)(), (C.c_float * 4)()
    env_, marg_, cond_ = (np.ascontiguousarray(t, np.float32) for t in (env, marg, cond))
    for _ in range(N // 2):
        r = rng.rand(2).astype(np.float32)
        O.oracle_envmap_sample(cpu_ctx.h, _p(r), a7)
        R.refglsl_envmap_sample(_p(env_), _p(marg_), _p(cond_), 64, 32, _p(xf), C.c_float(1.7), _p(r), b7)
        x, y = np.array(list(a7)), np.array(list(b7))
        assert np.all(np.abs(x - y) <= 1e-4 * np.maximum(1.0, np.abs(y))), (r, x, y)
        d = rng.randn(3).astype(np.float32)
        d /= np.linalg.norm(d)
        O.oracle_envmap_eval_pdf(cpu_ctx.h, _p(d), a4)
        R.refglsl_envmap_eval_pdf(_p(env_), _p(marg_), _p(cond_), 64, 32, _p(xf), C.c_float(1.7), _p(d), b4)
        x, y = np.array(list(a4)), np.array(list(b4))
        assert np.all(np.abs(x - y) <= 1e-4 * np.maximum(1.0, np.abs(y))), (d, x, y)


# ----------------------------------------------------------------------------- host-side reference arithmetic
def test_host_diffuse_fresnel_and_complex_ior(ref_lib):
    """computeDiffuseFresnel (loader/material.cpp:25-37; feeds fdrInt of plastic / rough_plastic) and the 40-row
    complex-IOR table (:250-291) against the scene loader's (asuna_b200/host.py, mirrored by host/scene.cpp)."""
    R = C.CDLL(ref_lib.path)
    R.refhost_diffuse_fresnel.restype = C.c_float
    for ior in (1.05, 1.33, 1.49, 1.5, 1.6, 1.9, 2.4, 0.8):
        want = R.refhost_diffuse_fresnel(C.c_float(ior), 1000)
        assert abs(host.compute_diffuse_fresnel(ior, 1000) - want) <= 2e-7 * max(1.0, abs(want)), ior
    names, buf = [], C.create_string_buffer(16)
    i = 0
    while R.refhost_complex_ior_name(i, buf):
        names.append(buf.value.decode())
        i += 1
    assert len(names) == 40 and set(names) == set(host.COMPLEX_IOR)
    eta, k = (C.c_float * 3)(), (C.c_float * 3)()
    for n in names:
        assert R.refhost_complex_ior(n.encode(), eta, k) == 1
        assert np.array_equal(np.array(list(eta), np.float32), np.asarray(host.COMPLEX_IOR[n][0], np.float32)), n
        assert np.array_equal(np.array(list(k), np.float32), np.asarray(host.COMPLEX_IOR[n][1], np.float32)), n


def test_host_camera_matrices(ref_lib):
    """rasterToCamera = invert(cameraToRaster) and cameraToWorld = invert_rot_trans(scale(1,-1,-1) look_at)
    (core/camera.cpp:28-99 on the reference's nvmath) against host.perspective_raster_to_camera / Scene.gpu_camera."""
    R = C.CDLL(ref_lib.path)
    rng = np.random.RandomState(13)
    out = (C.c_float * 16)()
    for w, h, fov in ((512, 512, 39.3), (1920, 1080, 45.0), (3840, 2160, 60.0), (64, 48, 0.5), (100, 300, 120.0)):
        R.refhost_raster_to_camera(w, h, C.c_float(fov), out)
        want = np.array(list(out), np.float32).reshape(4, 4).T  # column-major -> row-major
        got = host.perspective_raster_to_camera(w, h, fov)
        # both invert a matrix whose entries span 1e-3 .. 1e3: compare what the shader uses, transformPoint(pixel)
        for px in ((0.5, 0.5), (w - 0.5, h - 0.5), (w / 2, h / 3)):
            p = np.array([px[0], px[1], 0, 1], np.float64)
            a, b = got.astype(np.float64) @ p, want.astype(np.float64) @ p
            assert np.allclose(a[:3] / a[3], b[:3] / b[3], rtol=2e-5, atol=1e-6), (w, h, fov, px)
    for _ in range(200):
        eye, ctr = rng.randn(3).astype(np.float32) * 5, rng.randn(3).astype(np.float32)
        up = np.array([0, 1, 0], np.float32) if rng.rand() < 0.5 else rng.randn(3).astype(np.float32)
        R.refhost_camera_to_world(_p(eye), _p(ctr), _p(up), out)
        want = np.array(list(out), np.float32).reshape(4, 4).T
        got = host.invert_rot_trans(host.scaling((1, -1, -1)) @ host.look_at(eye, ctr, up))
        assert np.allclose(got, want, rtol=0, atol=2e-6 * max(1.0, float(np.abs(want).max()))), (eye, ctr, up)


def test_host_envmap_tables(ref_lib):
    """The marginal / conditional inverse-CDF tables of EnvMap::EnvMap (core/texture.cpp:149-226: fp32 running row
    sums, double division, lower-bound search) against host.envmap_tables, on maps with a bright spot, flat rows and
    a black row.  The sampled coordinates (.x) are integers / size and must be identical; pdfs (.y) within 1 ulp."""
    R = C.CDLL(ref_lib.path)
    rng = np.random.RandomState(14)
    maps = [scenes.procedural_envmap(64, 32, 3), scenes.procedural_envmap(128, 64, 9)]
    flat = np.ones((8, 16, 4), np.float32)
    flat[3] = 0.0
    maps.append(flat)
    maps.append(np.concatenate([rng.rand(16, 32, 3).astype(np.float32) ** 4 * 50, np.ones((16, 32, 1), np.float32)], -1))
    for env in maps:
        env = np.ascontiguousarray(env, np.float32)
        h, w = env.shape[:2]
        marg, cond = np.zeros_like(env), np.zeros_like(env)
        R.refhost_envmap_tables(_p(env), w, h, _p(marg), _p(cond))
        m2, c2 = host.envmap_tables(env)
        assert np.array_equal(m2[..., 0], marg[..., 0]) and np.array_equal(c2[..., 0], cond[..., 0])
        assert ulp_distance(m2[..., 1], marg[..., 1]).max() <= 1 and ulp_distance(c2[..., 1], cond[..., 1]).max() <= 1
        assert np.array_equal(m2[..., 2:], marg[..., 2:]) and np.array_equal(c2[..., 2:], cond[..., 2:])


# ----------------------------------------------------------------------------- one shader invocation at a time
MATERIAL_NAMES = {S.MAT_LAMBERTIAN: "lambertian", S.MAT_KANG18: "kang18", S.MAT_EMISSIVE: "emissive", S.MAT_PBR: "pbr",
                  S.MAT_PLASTIC: "plastic", S.MAT_ROUGH_PLASTIC: "rough_plastic", S.MAT_CONDUCTOR: "conductor",
                  S.MAT_ROUGH_CONDUCTOR: "rough_conductor", S.MAT_MIRROR: "mirror", S.MAT_DISNEY: "disney",
                  S.MAT_DIELECTRIC: "dielectric", S.MAT_PHONG: "phong"}


def probe_scene(rng, mtype, env, sunsky=False):
    """A tilted, scaled blob instance with a random material of one type, all four light kinds, textures."""
    sc = scenes.Scene()
    sc.set_camera("perspective", 16, 16, fov=40)
    sc.set_channels(["diffuse", "normal", "specular", "tangent", "roughness", "position", "uv"])
    tex = [sc.add_texture(f"t{k}", scenes.noise_texture(16, 40 + k, 3, 0.05, 0.95)) for k in range(3)]
    sc.add_material("m", H.random_material(rng, mtype, tex))
    sc.add_light(scenes.rect_light((-1, 3, -1), (1, 3, -1), (-1, 3, 1), (17, 12, 4), double_side=bool(rng.rand() < 0.5)))
    sc.add_light(scenes.point_light((2, 2, 2), (6, 6, 7)))
    sc.add_light(scenes.distant_light((0.2, 0.9, 0.3), (0.8, 0.7, 0.6)))
    v, i = scenes.quad((3, 0.5, 0), (4, 0.5, 0), (4, 0.5, 1), (3, 0.5, 1))
    sc.add_mesh_light((9, 9, 12), v, i[:3])
    if env:
        sc.set_envmap(scenes.procedural_envmap(32, 16, 5))
    if sunsky:
        sc.sunsky["in_use"] = 1
    sc.add_mesh("blob", *scenes.blob(1, 3, 0.2))
    x = host.translation((0.3, -0.2, 0.1)) @ host.rotation_y(0.7) @ host.scaling((1.3, 0.8, 1.1))
    sc.add_instance("blob", "m", x)
    sc.shots.append(host.Shot((0, 0, 5), (0, 0, 0), (0, 1, 0)))
    # useFaceNormal stays 0: the reference reads an unassigned field there (rchit_layouts.glsl:62) --
    # undefined in GLSL, zeros in libref.so (-ftrivial-auto-var-init=zero) -- and the oracle deliberately deviates.
    sc.state["ignoreEmissive"] = int(rng.rand() < 0.3)
    return sc


@pytest.mark.parametrize("mtype", sorted(MATERIAL_NAMES), ids=lambda t: MATERIAL_NAMES[t])
def test_closest_hit_shader_probes(mtype, cpu_ctx, ref_lib):
    """Every closest-hit shader's main() (bxdf/raytrace.*.rchit, incl. getHitState, texture fetches, opacity
    pass-through, AOV writes, sampleLights + eval + pdf + MIS, sampleBsdf, next ray), one invocation at a time:
    4 random materials x {lights, lights + env map, lights + sun/sky} x 1000 random hits = 12 k invocations per
    material type.  The RNG state after the shader (= number and order of draws), depth, stop, skip and the lobe
    flags must be identical on every probe; float outputs within 2e-4 of the vector's magnitude on >= 99.8 % of
    probes (a 1-ulp difference that flips a branch -- lobe choice against rand, a delta-lobe match -- moves the rest)."""
    from oracle.binding import OracleContext, RefContext
    rng = np.random.RandomState(100 + mtype)
    total, bad = 0, {}
    for rnd in range(4):
        for env, sunsky in ((False, False), (True, False), (False, True)):
            sc = probe_scene(rng, mtype, env, sunsky)
            a, b = OracleContext(), RefContext()
            sc.upload(a), sc.upload(b)
            sc.begin_shot(a, 0), sc.begin_shot(b, 0)
            inst = len(sc.instances) - 1
            args = H.random_probes(rng, 1000, len(sc.meshes[sc.instances[inst][1]][1]) // 3, inst)
            A, B = H.run_probes(a, *args), H.run_probes(b, *args)
            for f in ("seed", "depth", "stop", "drec_skip"):
                assert np.array_equal(A[f], B[f]), (MATERIAL_NAMES[mtype], f, float((A[f] != B[f]).mean()))
            flags_differ = float((A["brec_flags"] != B["brec_flags"]).mean())
            assert flags_differ <= 0.002, flags_differ
            for f, v in H.probe_mismatch(A, B).items():
                bad[f] = bad.get(f, 0.0) + v * len(A)
            total += len(A)
            a.close(), b.close()
    worst = max(bad.values(), default=0.0) / total
    assert worst <= 0.002, {k: v / total for k, v in bad.items()}


def test_emitter_hit_and_miss_shader_probes(ref_lib):
    """hitLight (brdf_lambertian.rchit:44-68: one-sided test, MIS against the area pdf) on rect and mesh-light
    instances, and raytrace.default.rmiss (bgColor / env map / sun-sky, MIS against the env pdf)."""
    from oracle.binding import OracleContext, RefContext
    rng = np.random.RandomState(77)
    for env, sunsky in ((False, False), (True, False), (False, True)):
        sc = probe_scene(rng, S.MAT_LAMBERTIAN, env, sunsky)
        sc.state["bgColor"] = (0.2, 0.3, 0.5)
        a, b = OracleContext(), RefContext()
        sc.upload(a), sc.upload(b)
        sc.begin_shot(a, 0), sc.begin_shot(b, 0)
        for inst, (_, mesh, _, light) in enumerate(sc.instances):
            if light < 0:
                continue
            args = H.random_probes(rng, 3000, len(sc.meshes[mesh][1]) // 3, inst)
            A, B = H.run_probes(a, *args), H.run_probes(b, *args)
            assert not H.probe_mismatch(A, B, 1e-5), H.probe_mismatch(A, B, 1e-5)
        args = H.random_probes(rng, 10000, 1, H.MISS)
        A, B = H.run_probes(a, *args), H.run_probes(b, *args)
        mm = H.probe_mismatch(A, B, 2e-5)
        assert not mm, mm
        a.close(), b.close()


# ----------------------------------------------------------------------------- whole frames
def image_agreement(A, B):
    """A, B: lists [radiance, aov...] from the two CPU paths fed identical RNG streams.  They differ only where an
    ulp-level difference flips a branch, so almost all pixels agree to ~1e-6."""
    x, y = A[0][..., :3].astype(np.float64), B[0][..., :3].astype(np.float64)
    rel = np.abs(x - y).max(axis=2) / (np.abs(y).max(axis=2) + 1e-3)
    out = {"radiance_frac_gt_1e-4": float((rel > 1e-4).mean()), "mean_rel": abs(x.mean() - y.mean()) / y.mean()}
    out["aov_frac_gt_1e-4"] = max([float((np.abs(p - q).max(axis=2) > 1e-4).mean()) for p, q in zip(A[1:], B[1:])], default=0.0)
    return out


def ref_scene_table():
    from test_gpu_parity import SCENES
    t = dict(SCENES)
    def dof():
        sc = scenes.cornell(64, 48, spp=6, depth=4)
        sc.camera["aperture"], sc.camera["focal_distance"] = 0.05, 2.0
        return sc
    def opencv():
        sc = scenes.cornell(64, 48, spp=6, depth=4)
        sc.set_camera("opencv", 64, 48, fxfycxcy=[56.0, 56.0, 32.0, 24.0])
        return sc
    def face_normal():
        sc = scenes.cornell_materials(64, 48, spp=6, env=False, lights="rect", textured=False)
        sc.state["ignoreEmissive"] = 1
        return sc
    t["thin_lens"], t["opencv"], t["ignore_emissive"] = dof, opencv, face_normal
    return t


@pytest.mark.parametrize("name", list(ref_scene_table()))
def test_frames_oracle_vs_reference_glsl(name, cpu_ctx, ref_ctx):
    """Whole frames through raytrace.projective.rgen (seed, jitter, camera models, bounce loop, shadow rays, Gaussian
    filter, clamp, running weighted mean, AOV stores on frame 0): every parity scene of tests/test_gpu_parity.py
    rendered by the restatement and by the reference GLSL.  >= 99.5 % of pixels within 1e-4 relative, image mean (a handful of branch-flipped pixels in a 2-7 k pixel image)
    within 5e-4, AOVs within 1e-4 on >= 99.9 % of pixels, identical ray counts to 0.1 %."""
    sc = ref_scene_table()[name]()
    sc.upload(cpu_ctx), sc.upload(ref_ctx)
    A, B = sc.render_shot(cpu_ctx, 0), sc.render_shot(ref_ctx, 0)
    r = image_agreement(A, B)
    assert r["radiance_frac_gt_1e-4"] <= 0.005 and r["mean_rel"] <= 5e-4 and r["aov_frac_gt_1e-4"] <= 0.001, r
    sa, sb = cpu_ctx.stats(), ref_ctx.stats()
    assert abs(sa["closest_rays"] - sb["closest_rays"]) <= 1e-3 * sb["closest_rays"]
    assert abs(sa["shadow_rays"] - sb["shadow_rays"]) <= 1e-3 * max(sb["shadow_rays"], 1)


def test_partitioned_accumulation_matches_reference_glsl(ref_lib):
    """Two frame-range partitions of the reference GLSL (rgen:171-178 accumulating onto zeroed planes) combine to
    the single-partition image: what the multi-GPU split relies on."""
    from oracle.binding import RefContext
    sc = scenes.cornell(48, 48, spp=6, depth=4)
    whole = RefContext()
    sc.upload(whole)
    W = sc.render_shot(whole, 0)[0]
    parts = []
    for r in range(2):
        c = RefContext()
        sc.upload(c)
        c.set_partition(r, 2)
        sc.render_shot(c, 0)
        ptr = C.cast(c.export_partial(), C.POINTER(C.c_float))  # host memory on the CPU paths
        parts.append(np.ctypeslib.as_array(ptr, shape=(48, 48, 4)).copy())
        c.close()
    s = parts[0] + parts[1]
    assert np.abs(s[..., :3] / s[..., 3:4] - W[..., :3]).max() <= 2e-5


def post_cases():
    """(name, GpuPushConstantPost) for all seven tone mappers (post.idle.frag:76-133), the custom one also with
    non-default grading and with global auto-exposure."""
    cases = [(n, S.default_post(n)) for n in S.TONE_MAPPERS]
    p = S.default_post("custom")
    p["contrast"], p["brightness"], p["saturation"], p["vignette"], p["avgLum"] = 1.2, 0.9, 0.7, 0.3, 1.4
    cases.append(("custom_graded", p))
    p = S.default_post("custom")
    p["autoExposure"], p["key"], p["Ywhite"] = 1, 0.4, 0.8
    cases.append(("custom_auto_exposure", p))
    return cases


def post_agreement(a, b):
    """Tone-mapped images: fp32 pow() chains agree to ~1e-6; the custom mapper's dither picks one of two 8-bit levels by
    a `<` on values an ulp apart, so a small fraction of channels may sit one quantisation step (1/255) away."""
    d = np.abs(a[..., :3].astype(np.float64) - b[..., :3])
    return float((d > 2e-5).mean()), float(d.max())


def test_post_process_oracle_vs_reference_glsl(cpu_ctx, ref_ctx):
    """PipelinePost: post.idle.frag compiled as C++ against the restatement (the one asuna_post_process is tested against
    on the GPU), every tone mapper, on a rendered HDR image with values from 0 to the clamp at 10."""
    sc = scenes.cornell_materials(64, 48, spp=4, env=True, lights="rect", textured=True)
    sc.upload(cpu_ctx), sc.upload(ref_ctx)
    A, B = sc.render_shot(cpu_ctx, 0)[0], sc.render_shot(ref_ctx, 0)[0]
    assert np.abs(A - B).max() < 0.2  # same picture (a few branch-flipped pixels aside)
    for name, tm in post_cases():
        a, b = cpu_ctx.post_process(tm), ref_ctx.post_process(tm)
        same_input = np.abs(A - B).max(axis=2) <= 1e-6
        frac, worst = post_agreement(a[same_input], b[same_input])
        assert frac <= (0.01 if name.startswith("custom") else 0.0) and worst <= (1.5 / 255 if name.startswith("custom") else 2e-5), (name, frac, worst)
        assert np.array_equal(a[..., 3], A[..., 3])


# ----------------------------------------------------------------------------- frozen vectors (always run)
def _golden(name):
    p = os.path.join(GOLDEN, name)
    if not os.path.exists(p):
        pytest.fail(f"{p} missing: run tools/make_ref_golden.py where /root/reference exists")
    return np.load(p)


@pytest.mark.parametrize("name", ["cornell", "materials", "materials_env", "pbr_sunsky", "all_materials"])
def test_oracle_against_frozen_reference_glsl_frames(name, cpu_ctx):
    """tests/golden/ref_frame_*.npz were rendered by libref.so (tools/make_ref_golden.py); the oracle must
    reproduce them without the reference being present."""
    from tools.make_ref_golden import FRAMES
    g = _golden(f"ref_frame_{name}.npz")
    sc = FRAMES[name]()
    sc.upload(cpu_ctx)
    A = sc.render_shot(cpu_ctx, 0)
    r = image_agreement(A, [g["radiance"]] + list(g["aov"]))
    assert r["radiance_frac_gt_1e-4"] <= 0.005 and r["mean_rel"] <= 5e-4 and r["aov_frac_gt_1e-4"] <= 0.001, r
    sc.begin_shot(cpu_ctx, 0)
    ids, _ = cpu_ctx.trace_primary()
    assert np.array_equal(ids, g["ids"])


def test_oracle_against_frozen_reference_glsl_probes(cpu_ctx):
    """tests/golden/ref_probes.npz: 256 shader invocations per material type + emitter + miss, inputs and the
    reference GLSL's outputs."""
    from oracle.binding import OracleContext
    from tools.make_ref_golden import probe_cases
    g = _golden("ref_probes.npz")
    for key, sc, args in probe_cases():
        ctx = OracleContext()
        sc.upload(ctx)
        sc.begin_shot(ctx, 0)
        A = H.run_probes(ctx, *args)
        B = g[key].view(H.PROBE).reshape(-1)
        for f in ("seed", "depth", "stop", "drec_skip"):
            assert np.array_equal(A[f], B[f]), (key, f)
        mm = H.probe_mismatch(A, B)
        assert max(mm.values(), default=0.0) <= 0.01, (key, mm)
        ctx.close()
