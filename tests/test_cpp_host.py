"""The C++ host (host/asuna_b200): scene JSON loader, image readers / writers, tone mappers.
CPU-only: `--dump-scene` stops after loading, `--convert` exercises the image code; both are compared
with the Python mirror (asuna_b200/host.py), PIL / OpenCV and closed forms."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from asuna_b200 import host, scenes, structs as S  # noqa: E402
import gen_scenes  # noqa: E402

CLI = os.path.join(ROOT, "host", "asuna_b200")


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")])
    return CLI


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    p = 0
    while p < len(data):
        (nl,) = struct.unpack_from("<I", data, p)
        name = data[p + 4:p + 4 + nl].decode()
        (nb,) = struct.unpack_from("<Q", data, p + 4 + nl)
        p += 12 + nl
        out[name] = data[p:p + nb]
        p += nb
    return out


SCENES = {
    "cornell": lambda: scenes.cornell(32, 24, spp=3, depth=4),
    "materials": lambda: scenes.cornell_materials(24, 16, spp=2, depth=4, env=True, lights="all", textured=True),
    "all_materials": lambda: scenes.cornell_all_materials(24, 16, spp=2, depth=4, env=True, lights="all", textured=True),
    "pbr": lambda: scenes.pbr_spheres(24, 16, spp=2, depth=3, subdiv=2, tex_size=8),
    "field": lambda: scenes.instanced_field(16, 16, spp=1, depth=2, subdiv=1, grid=2),
}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_loader_matches_python_mirror(cli, tmp_path, name):
    path = gen_scenes.write_scene(SCENES[name](), str(tmp_path), name)
    py = host.load_scene_json(path)
    dump = str(tmp_path / "scene.bin")
    subprocess.check_call([cli, "--scene", path, "--dump-scene", dump])
    d = read_dump(dump)
    assert np.frombuffer(d["film"], np.int32).tolist() == [py.camera["width"], py.camera["height"]]
    assert d["materials"] == np.array(py.materials, S.Material).tobytes()
    assert d["lights"] == np.array(py.lights, S.Light).tobytes()
    assert d["sunsky"] == py.sunsky.tobytes()
    for i, t in enumerate(py.textures):
        assert np.frombuffer(d[f"texture_size:{i}"], np.int32).tolist() == [t.shape[1], t.shape[0]]
        assert d[f"texture:{i}"] == np.ascontiguousarray(t, np.float32).tobytes()
    if py.envmap is not None:
        assert d["envmap"] == py.envmap[0].tobytes()
        assert np.array_equal(np.frombuffer(d["env_marginal"], np.float32), py.envmap[1].reshape(-1))
        assert np.array_equal(np.frombuffer(d["env_conditional"], np.float32), py.envmap[2].reshape(-1))
    assert sum(k.startswith("mesh_vertices:") for k in d) == len(py.meshes)
    for i, (v, idx) in enumerate(py.meshes):
        assert d[f"mesh_vertices:{i}"] == v.tobytes() and d[f"mesh_indices:{i}"] == idx.tobytes()
    assert sum(k.startswith("instance_ids:") for k in d) == len(py.instances)
    for i, (x, mesh, mat, light) in enumerate(py.instances):
        assert np.frombuffer(d[f"instance_ids:{i}"], np.int32).tolist() == [mesh, mat, light]
        assert np.allclose(np.frombuffer(d[f"instance_xform:{i}"], np.float32), host.colmajor(x), rtol=0, atol=0)
    for i, shot in enumerate(py.shots):
        cam_c = np.frombuffer(d[f"shot_camera:{i}"], S.Camera)[0]
        cam_p = py.gpu_camera(shot)
        for k in S.Camera.names:
            assert np.allclose(cam_c[k], cam_p[k], rtol=2e-6, atol=2e-6), k
        assert d[f"shot_state:{i}"] == py.shot_state(i).tobytes()


def test_loader_transform_chain_shots_and_errors(cli, tmp_path):
    """toworld singletons, the three shot types, per-shot state, and the reference's error messages."""
    path = gen_scenes.write_scene(scenes.cornell(16, 16, spp=2, depth=3), str(tmp_path), "c")
    import json
    js = json.load(open(path))
    js["instances"][0]["toworld"] = [{"type": "scale", "value": [1, 2, 3]}, {"type": "rotx", "value": 30}, {"type": "roty", "value": -40},
                                     {"type": "rotz", "value": 50}, {"type": "rotate", "value": [10, 20, 30]},
                                     {"type": "translate", "value": [0.5, -0.25, 2]}]
    c2w = host.invert_rot_trans(host.look_at((1, 2, 3), (0, 0.5, 0), (0, 1, 0)))
    js["shots"] += [{"type": "toworld", "matrix": [float(v) for v in c2w.reshape(-1)]},
                    {"type": "opencv", "matrix": [float(v) for v in host.look_at((1, 2, 3), (0, 0.5, 0), (0, 1, 0)).reshape(-1)],
                     "state": {"path_tracing": {"spp": 7, "max_path_depth": 2, "background_color": [0.1, 0.2, 0.3]}},
                     "env_toworld": [{"type": "roty", "value": 90}, {"type": "rotx", "value": 15}]}]
    js["state"]["post_processing"] = {"tone_mapping": "Aces"}
    json.dump(js, open(path, "w"))
    py = host.load_scene_json(path)
    dump = str(tmp_path / "scene.bin")
    subprocess.check_call([cli, "--scene", path, "--dump-scene", dump])
    d = read_dump(dump)
    assert np.allclose(np.frombuffer(d["instance_xform:1"], np.float32), host.colmajor(py.instances[1][0]), atol=1e-6)
    assert d["tone_mapping"] == b"Aces"
    for i, shot in enumerate(py.shots):
        cam_c, cam_p = np.frombuffer(d[f"shot_camera:{i}"], S.Camera)[0], py.gpu_camera(shot)
        for k in ("cameraToWorld", "envTransform"):
            assert np.allclose(cam_c[k], cam_p[k], atol=2e-6), (i, k)
        assert d[f"shot_state:{i}"] == py.shot_state(i).tobytes()
    assert np.frombuffer(d["shot_state:2"], S.State)[0]["spp"] == 7
    # errors: same conditions and wording as the reference's loader, as a non-zero exit instead of exit(1) deep inside
    for mutate, msg in ((lambda j: j.pop("camera"), 'missing key ["camera"]'),
                        (lambda j: j["materials"].append({"type": "brdf_unknown", "name": "x"}), "unrecognized material type [brdf_unknown]"),
                        (lambda j: j["lights"].append({"type": "laser", "radiance": [1, 1, 1]}), "unrecognized light type [laser]"),
                        (lambda j: j["instances"].append({"mesh": "nope", "material": "white"}), "mesh [nope] does not exist"),
                        (lambda j: j["shots"].append({"type": "orbit"}), "unrecognized shot type [orbit]"),
                        # env_toworld bans translations: the singleton falls through to the error branch (loader.cpp:374,389-392)
                        (lambda j: j["shots"][0].update(env_toworld=[{"type": "translate", "value": [1, 2, 3]}]),
                         "unrecognized toworld singleton type [translate]")):
        bad = json.load(open(path))
        mutate(bad)
        bp = str(tmp_path / "bad.json")
        json.dump(bad, open(bp, "w"))
        r = subprocess.run([cli, "--scene", bp, "--dump-scene", dump], capture_output=True, text=True)
        assert r.returncode != 0 and msg in r.stderr, r.stderr


def test_autofit_camera_when_no_shots(cli, tmp_path):
    path = gen_scenes.write_scene(scenes.cornell(16, 16, spp=1, depth=2), str(tmp_path), "c")
    import json
    js = json.load(open(path))
    js.pop("shots")
    json.dump(js, open(path, "w"))
    dump = str(tmp_path / "scene.bin")
    subprocess.check_call([cli, "--scene", path, "--dump-scene", dump])
    cam = np.frombuffer(read_dump(dump)["shot_camera:0"], S.Camera)[0]
    c2w = cam["cameraToWorld"].reshape(4, 4).T
    eye, fwd = c2w[:3, 3], c2w[:3, 2]
    assert np.allclose(fwd, -np.ones(3) / np.sqrt(3), atol=1e-5)  # looking down the (10,10,10) -> centre diagonal
    centre = np.array([0.5, 0.5, 0.5])
    assert np.linalg.norm(np.cross(centre - eye, fwd)) < 1e-3 and np.dot(centre - eye, fwd) > 0.8  # the box centre is on the axis


def _convert(cli, src, dst, *extra):
    subprocess.check_call([cli, "--scene", "-", "--convert", src, dst, *[str(e) for e in extra]])


def test_png_reader_and_writer(cli, tmp_path):
    from PIL import Image
    rng = np.random.RandomState(0)
    cases = {"rgb": rng.randint(0, 256, (5, 7, 3), np.uint8), "rgba": rng.randint(0, 256, (6, 4, 4), np.uint8),
             "gray": rng.randint(0, 256, (3, 9), np.uint8)}
    for name, a in cases.items():
        src = str(tmp_path / f"{name}.png")
        Image.fromarray(a).save(src)
        for gamma in (1.0, 2.2):
            dst = str(tmp_path / f"{name}.npy")
            _convert(cli, src, dst, gamma)
            got = np.load(dst)
            want = host.load_image(src, gamma)
            assert got.shape == want.shape and np.allclose(got, want, atol=2e-7), name
    pal = Image.fromarray(cases["rgb"]).convert("P", palette=Image.ADAPTIVE, colors=16)
    pal.save(str(tmp_path / "pal.png"))
    _convert(cli, str(tmp_path / "pal.png"), str(tmp_path / "pal.npy"))
    assert np.allclose(np.load(str(tmp_path / "pal.npy"))[..., :3], np.asarray(pal.convert("RGB"), np.float32) / 255, atol=2e-7)
    a16 = rng.randint(0, 65536, (4, 4)).astype(np.uint16)
    Image.fromarray(a16).save(str(tmp_path / "g16.png"))
    _convert(cli, str(tmp_path / "g16.png"), str(tmp_path / "g16.npy"))
    assert np.allclose(np.load(str(tmp_path / "g16.npy"))[..., 0], (a16 >> 8) / 255.0, atol=2e-7)  # stb keeps the high byte
    # writer: stb's hdr_to_ldr with gamma 1 -- clamp(int(x * 255 + 0.5))
    f = rng.rand(5, 6, 4).astype(np.float32) * 1.4 - 0.2
    np.save(str(tmp_path / "f.npy"), f)
    _convert(cli, str(tmp_path / "f.npy"), str(tmp_path / "f.png"))
    back = np.asarray(Image.open(str(tmp_path / "f.png")))
    assert back.shape == (5, 6, 4) and np.array_equal(back, np.clip((f * 255 + 0.5).astype(np.int64), 0, 255).astype(np.uint8))
    # a film-sized image: the IDAT stream is deflated in bands by several threads and must read as one zlib stream
    # (PIL, cv2 and the repository's own reader), smooth content and noise alike
    yy, xx = np.mgrid[0:540, 0:960].astype(np.float32)
    big = np.stack([xx / 960, yy / 540, 0.5 + 0.5 * np.sin(xx * 0.05) * np.cos(yy * 0.03), np.ones_like(xx)], axis=2).astype(np.float32)
    big[200:300, 300:500] = rng.rand(100, 200, 4)
    np.save(str(tmp_path / "big.npy"), big)
    _convert(cli, str(tmp_path / "big.npy"), str(tmp_path / "big.png"))
    want = np.clip((big * 255 + 0.5).astype(np.int64), 0, 255).astype(np.uint8)
    assert np.array_equal(np.asarray(Image.open(str(tmp_path / "big.png"))), want)
    import cv2
    assert np.array_equal(cv2.imread(str(tmp_path / "big.png"), cv2.IMREAD_UNCHANGED)[..., [2, 1, 0, 3]], want)
    _convert(cli, str(tmp_path / "big.png"), str(tmp_path / "big_back.npy"))
    assert np.allclose(np.load(str(tmp_path / "big_back.npy")), want / 255.0, atol=2e-7)


def test_hdr_exr_pfm_round_trips(cli, tmp_path):
    rng = np.random.RandomState(1)
    f = np.ones((7, 9, 4), np.float32)
    f[..., :3] = np.exp(rng.randn(7, 9, 3) * 2).astype(np.float32)
    np.save(str(tmp_path / "f.npy"), f)
    # Radiance RGBE: 8-bit mantissa shared exponent -> 1/128 relative to the largest channel
    _convert(cli, str(tmp_path / "f.npy"), str(tmp_path / "f.hdr"))
    _convert(cli, str(tmp_path / "f.hdr"), str(tmp_path / "f_hdr.npy"))
    got = np.load(str(tmp_path / "f_hdr.npy"))
    assert np.all(np.abs(got[..., :3] - f[..., :3]) <= f[..., :3].max(axis=2, keepdims=True) / 128 + 1e-6)
    import cv2
    ref = cv2.imread(str(tmp_path / "f.hdr"), cv2.IMREAD_UNCHANGED)[..., ::-1]
    assert np.allclose(got[..., :3], ref, rtol=1e-6)  # an independent RGBE decoder agrees with ours
    # EXR: half precision RGB, alpha dropped (WRITE_RGB): values are exactly float16(f)
    _convert(cli, str(tmp_path / "f.npy"), str(tmp_path / "f.exr"))
    _convert(cli, str(tmp_path / "f.exr"), str(tmp_path / "f_exr.npy"))
    got = np.load(str(tmp_path / "f_exr.npy"))
    assert np.array_equal(got[..., :3], f[..., :3].astype(np.float16).astype(np.float32)) and np.all(got[..., 3] == 1)
    head = open(str(tmp_path / "f.exr"), "rb").read(4)
    assert head == bytes([0x76, 0x2F, 0x31, 0x01])
    # an independent OpenEXR implementation (the one inside OpenCV) reads our file, and ours reads its file
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    ext = cv2.imread(str(tmp_path / "f.exr"), cv2.IMREAD_UNCHANGED)
    if ext is not None:  # OpenCV builds without OpenEXR return None
        assert ext.dtype == np.float32 and np.array_equal(ext[..., ::-1], f[..., :3].astype(np.float16).astype(np.float32))
        for flags in ([cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT]):
            if cv2.imwrite(str(tmp_path / "cv.exr"), np.ascontiguousarray(f[..., 2::-1]), flags):
                _convert(cli, str(tmp_path / "cv.exr"), str(tmp_path / "cv_exr.npy"))
                back = np.load(str(tmp_path / "cv_exr.npy"))[..., :3]
                want = f[..., :3].astype(np.float16).astype(np.float32) if flags[1] == cv2.IMWRITE_EXR_TYPE_HALF else f[..., :3]
                assert np.array_equal(back, want)
    _convert(cli, str(tmp_path / "f.npy"), str(tmp_path / "f.pfm"))
    _convert(cli, str(tmp_path / "f.pfm"), str(tmp_path / "f_pfm.npy"))
    assert np.array_equal(np.load(str(tmp_path / "f_pfm.npy"))[..., :3], f[..., :3])


def test_jpeg_reader_against_libjpeg(cli, tmp_path):
    """The reference reads jpg textures through stb_image (src/core/texture.cpp:307-336); the host's baseline JPEG decoder
    against libjpeg (PIL) on 4:4:4 / 4:2:0 / 4:2:2 / greyscale / optimised-Huffman / restart-interval files: within 3 of
    255 (IDCT and chroma-filter rounding), and gamma applied like the PNG path.  Progressive files are refused by name."""
    from PIL import Image
    import cv2
    rng = np.random.RandomState(3)
    y, x = np.mgrid[0:37, 0:53]
    img = np.stack([(np.sin(x / 7.0) * 0.5 + 0.5) * 255, (np.cos(y / 5.0) * 0.5 + 0.5) * 255, ((x + y) % 64) * 4], -1).astype(np.uint8)
    img[10:20, 10:30] = rng.randint(0, 255, (10, 20, 3))
    cases = {"444": dict(quality=95, subsampling=0), "420": dict(quality=90, subsampling=2), "422": dict(quality=85, subsampling=1),
             "opt": dict(quality=92, subsampling=0, optimize=True), "grey": dict(quality=90)}
    for name, kw in cases.items():
        path = str(tmp_path / f"{name}.jpg")
        Image.fromarray(img[..., 0] if name == "grey" else img).save(path, **kw)
        _convert(cli, path, str(tmp_path / f"{name}.npy"))
        got = np.load(str(tmp_path / f"{name}.npy"))
        ref = np.asarray(Image.open(path).convert("RGB"), np.float32) / 255
        assert got.shape == (37, 53, 4) and np.all(got[..., 3] == 1)
        assert np.abs(got[..., :3] - ref).max() <= 3.01 / 255, name
    path = str(tmp_path / "rst.jpg")
    if cv2.imwrite(path, img[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 4]):
        _convert(cli, path, str(tmp_path / "rst.npy"))
        ref = np.asarray(Image.open(path).convert("RGB"), np.float32) / 255
        assert np.abs(np.load(str(tmp_path / "rst.npy"))[..., :3] - ref).max() <= 3.01 / 255
    _convert(cli, str(tmp_path / "444.jpg"), str(tmp_path / "g.npy"), 2.2)
    ref = (np.asarray(Image.open(str(tmp_path / "444.jpg")).convert("RGB"), np.float32) / 255) ** 2.2
    assert np.abs(np.load(str(tmp_path / "g.npy"))[..., :3] - ref).max() <= 0.03
    Image.fromarray(img).save(str(tmp_path / "prog.jpg"), progressive=True)
    r = subprocess.run([cli, "--scene", "-", "--convert", str(tmp_path / "prog.jpg"), str(tmp_path / "p.npy")], capture_output=True, text=True)
    assert r.returncode != 0 and "progressive" in r.stderr


def test_tone_mappers_match_closed_forms(cli, tmp_path):
    """reference src/shaders/post.idle.frag:76-133"""
    x = np.ones((1, 64, 4), np.float32)
    x[0, :, :3] = np.linspace(0, 4, 64, dtype=np.float32)[:, None]
    np.save(str(tmp_path / "x.npy"), x)
    v = x[0, :, 0].astype(np.float64)
    want = {
        "none": v,
        "gamma": (v / (1 + v / 1.5)) ** (1 / 2.2),
        "filmic": (lambda t: (t * (6.2 * t + 0.5)) / (t * (6.2 * t + 1.7) + 0.06))(np.maximum(0, v - 0.004)),
        "Aces": np.clip((v * (2.51 * v + 0.03)) / (v * (2.43 * v + 0.59) + 0.14), 0, 1) ** (1 / 2.2),
        "pbrt": np.where(v < 0.0031308, 12.92 * v, 1.055 * np.maximum(v, 1e-30) ** (1 / 2.4) - 0.055),
    }
    want["reinhard"] = want["filmic"]
    for name, w in want.items():
        _convert(cli, str(tmp_path / "x.npy"), str(tmp_path / "y.npy"), 1.0, name)
        assert np.allclose(np.load(str(tmp_path / "y.npy"))[0, :, 0], w, atol=2e-6), name
    r = subprocess.run([cli, "--scene", "-", "--convert", str(tmp_path / "x.npy"), str(tmp_path / "y.npy"), "1.0", "custom"], capture_output=True)
    assert r.returncode != 0
