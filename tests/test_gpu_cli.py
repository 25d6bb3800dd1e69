"""The headless C++ command line (host/asuna_b200) end to end on the GPU: scene JSON in, images out, compared with
the same scene rendered through the Python mirror + C ABI and with the CPU oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from asuna_b200 import host, scenes  # noqa: E402
import gen_scenes  # noqa: E402

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "host", "asuna_b200")


def _run(args):
    r = subprocess.run([CLI] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r


def test_cli_renders_cornell_like_the_python_path(gpu_ctx, cpu_ctx, tmp_path):
    sc = scenes.cornell(96, 64, spp=16, depth=5)
    path = gen_scenes.write_scene(sc, str(tmp_path), "cornell")
    js = json.load(open(path))
    js["shots"].append({"type": "lookat", "eye": [0.5, 0.5, 2.0], "lookat": [0.5, 0.4, 0.0], "up": [0, 1, 0],
                        "state": {"path_tracing": {"spp": 4, "max_path_depth": 2}}})
    json.dump(js, open(path, "w"))
    out = str(tmp_path / "img")
    report = str(tmp_path / "report.json")
    _run(["--offline", "--scene", path, "--out", out, "--output_f32", "--report", report])
    py = host.load_scene_json(path)
    py.upload(gpu_ctx)
    py.upload(cpu_ctx)
    for shot in range(2):
        ref = py.render_shot(gpu_ctx, shot)
        orc = py.render_shot(cpu_ctx, shot)
        got = np.load(f"{out}_shot_{shot:04d}.exr.npy")
        assert got.shape == (64, 96, 4)
        # same wire structs up to 1 ulp in the camera matrices: a handful of paths may take another branch
        assert np.mean(np.abs(got - ref[0]).max(axis=2) > 1e-4) < 2e-3
        assert abs(got[..., :3].mean() - orc[0][..., :3].mean()) / orc[0][..., :3].mean() < 0.01
        for cid in range(3):
            aov = np.load(f"{out}_shot_{shot:04d}_channel_{cid:04d}.exr.npy")
            assert np.mean(np.abs(aov - ref[1 + cid]).max(axis=2) > 1e-4) < 2e-3
        # the EXR itself: half-precision RGB of the same plane
        back = str(tmp_path / "back.npy")
        subprocess.check_call([CLI, "--scene", "-", "--convert", f"{out}_shot_{shot:04d}.exr", back])
        assert np.array_equal(np.load(back)[..., :3], got[..., :3].astype(np.float16).astype(np.float32))
    rep = json.load(open(report))
    assert [s["spp"] for s in rep["shots"]] == [16, 4] and rep["bvh_build_ms"] > 0


def test_cli_ldr_output_and_scanline_mode(tmp_path):
    sc = scenes.cornell(48, 32, spp=4, depth=3)
    path = gen_scenes.write_scene(sc, str(tmp_path), "c")
    js = json.load(open(path))
    js["state"]["output_hdr"] = False
    js["state"]["path_tracing"]["multi_channel"] = ["position", "normal"]
    js["state"]["path_tracing"]["multi_channel_ldr"] = [False, True]
    json.dump(js, open(path, "w"))
    out = str(tmp_path / "o")
    _run(["--offline", "--scene", path, "--out", out, "--output_f32"])
    from PIL import Image
    png = np.asarray(Image.open(out + "_shot_0000.png"))
    tm = np.load(out + "_shot_0000.png.npy")  # the tone-mapped (filmic, default) plane the PNG was quantised from
    assert png.shape == (32, 48, 4) and np.array_equal(png, np.clip((tm * 255 + 0.5).astype(np.int64), 0, 255).astype(np.uint8))
    assert os.path.exists(out + "_shot_0000_channel_0000.exr") and os.path.exists(out + "_shot_0000_channel_0001.png")
    # --output_scanline: channels >= 1 become (n_valid, 3) float32 arrays of the pixels flagged in channel 1
    out2 = str(tmp_path / "s")
    _run(["--offline", "--scene", path, "--out", out2, "--output_scanline"])
    a = np.load(out2 + "_shot_0000_channel_0000.exr")
    assert a.ndim == 2 and a.shape[1] == 3


def test_cli_two_gpus_match_one(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    path = gen_scenes.write_scene(scenes.cornell(64, 64, spp=8, depth=4), str(tmp_path), "c")
    _run(["--offline", "--scene", path, "--out", str(tmp_path / "one"), "--output_f32"])
    _run(["--offline", "--scene", path, "--out", str(tmp_path / "two"), "--output_f32", "--gpus", "2"])
    a, b = np.load(str(tmp_path / "one_shot_0000.exr.npy")), np.load(str(tmp_path / "two_shot_0000.exr.npy"))
    assert np.allclose(a[..., :3], b[..., :3], rtol=3e-5, atol=3e-6)  # same samples, other summation order
    assert np.array_equal(np.load(str(tmp_path / "one_shot_0000_channel_0000.exr.npy")), np.load(str(tmp_path / "two_shot_0000_channel_0000.exr.npy")))


def test_cli_two_gpus_split_by_shots(tmp_path):
    """--split shots (replicas): every GPU renders whole shots, nothing is exchanged, so each image is
    the one a single GPU writes, bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sc = scenes.cornell(48, 40, spp=4, depth=4)
    scenes.orbit_shots(sc, 3, (0.5, 0.5, 0.5), 2.2, 0.5)
    path = gen_scenes.write_scene(sc, str(tmp_path), "c")
    _run(["--offline", "--scene", path, "--out", str(tmp_path / "one"), "--output_f32"])
    _run(["--offline", "--scene", path, "--out", str(tmp_path / "two"), "--output_f32", "--gpus", "2", "--split", "shots"])
    for shot in range(3):
        for suffix in (".exr.npy", "_channel_0000.exr.npy"):
            a = np.load(str(tmp_path / f"one_shot_{shot:04d}{suffix}"))
            b = np.load(str(tmp_path / f"two_shot_{shot:04d}{suffix}"))
            assert np.array_equal(a, b, equal_nan=True), (shot, suffix)
