"""Host-side logic (scene model, JSON loader) and the C-ABI surface, runnable without a GPU."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from asuna_b200 import capi, host, scenes, structs as S


def test_library_loads_and_exports_every_declared_symbol():
    lib = C.CDLL(capi.PRODUCT_LIB)
    header = open(os.path.join(ROOT, "include", "asuna_b200.h")).read()
    declared = set(re.findall(r"\b(asuna_[a-z_]+)\s*\(", header))
    assert declared == {"asuna_" + s for s in capi.ABI_SYMBOLS}
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/asuna_b200.h but not exported"
    sizes = (C.c_uint32 * 6)()
    lib.asuna_abi_sizes(sizes)
    assert tuple(sizes) == S.EXPECTED_SIZES


def test_oracle_mirrors_the_same_abi(oracle_lib):
    for s in capi.ABI_SYMBOLS:
        if s in ("channel_device_ptr", "accel_stats", "stream_handle", "set_counting", "set_profiling", "host_alloc", "host_free"):
            continue  # device-only introspection
        assert oracle_lib.has(s), f"oracle_{s} missing"


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.AsunaError):
        capi.Context(gpu_id=0)


def test_scene_ids_follow_reference_insertion_order():
    sc = scenes.cornell(32, 32)
    # dummies at index 0 (scene.cpp:93-111); the rect light owns mesh 0 and instance 0 (lights before meshes)
    assert sc.texture_ids["add_by_default_dummy_texture"] == 0 and sc.material_ids["white"] == 1
    assert len(sc.lights) == 2 and sc.lights[0]["type"] == S.LIGHT_DIRECTIONAL
    assert sc.mesh_ids["__rectLight:1"] == 0 and sc.instances[0][3] == 1
    l = sc.lights[1]
    n = np.cross(l["u"], l["v"])
    assert n[1] < 0 and np.isclose(l["area"], 0.09, atol=1e-6)
    st = sc.shot_state(0)
    assert st["numLights"] == 1 and st["positionOutChannel"] == 2 and st["nMultiChannel"] == 3


def test_toworld_chain_and_rotate_order():
    m = host._parse_toworld([{"type": "scale", "value": [2, 2, 2]}, {"type": "translate", "value": [1, 0, 0]}])
    assert np.allclose(m @ [1, 1, 1, 1], [3, 2, 2, 1])  # applied in order: scale, then translate (loader.cpp:394)
    r = host._parse_toworld([{"type": "rotate", "value": [90, 0, 90]}])
    want = host.rotation_z(np.radians(90)) @ host.rotation_y(0) @ host.rotation_x(np.radians(90))
    assert np.allclose(r, want, atol=1e-6)
    with pytest.raises(ValueError):  # env_toworld bans translation: the reference falls into its error branch (loader.cpp:374,389-392,454)
        host._parse_toworld([{"type": "translate", "value": [1, 2, 3]}], True)


def test_json_round_trip_matches_procedural_scene(tmp_path):
    from tools import gen_scenes
    sc = scenes.cornell_materials(40, 30, spp=3, env=True, lights="all", textured=True)
    path = gen_scenes.write_scene(sc, str(tmp_path), "mat")
    js = json.load(open(path))
    assert {"state", "camera", "meshes", "instances"} <= set(js)
    sc2 = host.load_scene_json(path)
    assert len(sc2.meshes) == len(sc.meshes) and len(sc2.instances) == len(sc.instances)
    assert len(sc2.lights) == len(sc.lights) and len(sc2.materials) == len(sc.materials)
    for a, b in zip(sc.lights, sc2.lights):
        for k in S.Light.names:
            assert np.allclose(a[k], b[k], atol=1e-6), k
    for a, b in zip(sc.materials, sc2.materials):
        for k in S.Material.names:
            assert np.allclose(a[k], b[k], atol=1e-6), k
    for (v, i), (v2, i2) in zip(sc.meshes, sc2.meshes):
        # the OBJ loader unrolls vertices per face corner (mesh.cpp:123-143)
        assert np.allclose(v["pos"][i], v2["pos"][i2], atol=1e-6)
        assert np.allclose(v["uv"][i], v2["uv"][i2], atol=1e-6)
        assert np.allclose(v["normal"][i], v2["normal"][i2], atol=1e-6)
    for a, b in zip(sc.instances, sc2.instances):
        assert np.allclose(a[0], b[0], atol=1e-6) and a[1:] == b[1:]
    for a, b in zip(sc.textures, sc2.textures):
        assert np.allclose(a, b, atol=1e-6)
    assert np.allclose(sc.envmap[0], sc2.envmap[0], atol=1e-6)
    for k in S.State.names:
        assert np.allclose(sc.shot_state(0)[k], sc2.shot_state(0)[k]), k
    ca, cb = sc.gpu_camera(sc.shots[0]), sc2.gpu_camera(sc2.shots[0])
    for k in S.Camera.names:
        assert np.allclose(ca[k], cb[k], atol=1e-6), k


def test_per_shot_state_override_takes_only_six_fields(tmp_path):
    from tools import gen_scenes
    sc = scenes.cornell(16, 16, spp=2)
    path = gen_scenes.write_scene(sc, str(tmp_path), "c")
    js = json.load(open(path))
    js["shots"][0]["state"] = {"path_tracing": {"spp": 9, "max_path_depth": 2, "multi_channel": ["uv"]}}
    json.dump(js, open(path, "w"))
    sc2 = host.load_scene_json(path)
    st = sc2.shot_state(0)
    assert st["spp"] == 9 and st["maxPathDepth"] == 2
    assert st["nMultiChannel"] == 3 and st["uvOutChannel"] == -1  # multi_channel stays global (scene.cpp:439-453)


def test_opencv_and_toworld_shots(tmp_path):
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [1, 2, 3]
    w2c = np.linalg.inv(c2w)
    base = {"state": {}, "camera": {"type": "opencv", "film": {"resolution": [8, 8]}, "fx": 10, "fy": 10, "cx": 4, "cy": 4},
            "meshes": [], "instances": [],
            "shots": [{"type": "opencv", "matrix": w2c.reshape(-1).tolist()}, {"type": "toworld", "matrix": c2w.reshape(-1).tolist()}]}
    p = tmp_path / "s.json"
    json.dump(base, open(p, "w"))
    sc = host.load_scene_json(str(p))
    s0, s1 = sc.shots
    assert np.allclose(s0.eye, [1, 2, 3]) and np.allclose(s0.up, [0, -1, 0]) and np.allclose(s0.lookat, [1, 2, 4])
    assert np.allclose(s1.eye, [1, 2, 3]) and np.allclose(s1.up, [0, 1, 0]) and np.allclose(s1.lookat, [1, 2, 4])
    cam = sc.gpu_camera(s0)
    assert cam["type"] == S.CAMERA_OPENCV and np.allclose(cam["fxfycxcy"], [10, 10, 4, 4])
    # OpenCV convention: camera y is down, so world "up" (0,-1,0) maps to camera -y
    assert np.allclose(cam["cameraToWorld"].reshape(4, 4).T[:3, :3], np.eye(3), atol=1e-6)


def test_flip_metric_behaves():
    from asuna_b200 import metrics
    rng = np.random.RandomState(0)
    a = rng.rand(32, 32, 3)
    assert metrics.flip(a, a) == 0.0
    assert metrics.flip(a, a + 0.0005) < 0.01 < metrics.flip(a, a + 0.01) < metrics.flip(a, np.zeros_like(a))
    assert metrics.mean_relative_error(a * 1.005, a) == pytest.approx(0.005, rel=1e-6)


def test_header_is_c_and_every_symbol_links_from_c(tmp_path):
    """tests/abi_smoke.c: include/asuna_b200.h must parse as C11 (-pedantic -Werror) and every declared entry point must
    resolve against libasuna_b200.so from a plain C program (extern "C", no C++ types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    src = os.path.join(ROOT, "tests", "abi_smoke.c")
    text = open(src).read()
    for s in capi.ABI_SYMBOLS:
        assert f"TAKE(asuna_{s})" in text, f"tests/abi_smoke.c does not reference asuna_{s}"
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(capi.PRODUCT_LIB)
    subprocess.check_call([gcc, "-std=c11", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", libdir, "-lasuna_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    fields = [int(x) for x in out.stdout.split()]
    assert fields[0] == len(capi.ABI_SYMBOLS) and tuple(fields[1:7]) == S.EXPECTED_SIZES == tuple(fields[7:13])
