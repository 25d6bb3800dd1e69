"""Small workloads for compute-sanitizer (tools/sanitize.sh): every kernel family of the library on inputs that finish
in seconds under memcheck / racecheck.
  smoke      : 64x64 Cornell, 4 frames, depth 4 (flattened single-level kernels, shade kernels, accumulate)
  two_level  : instanced field with flattening off (instance-level traversal) + all twelve materials + env map
  sort       : the builder's radix sort on 100 k random 64-bit keys, checked against numpy
  build      : a 9 k-triangle scene (PLOC grid rounds on shared-memory tiles, hand-over to the single-block tail, several
               emit levels, two sort tiles per pass) + the asynchronous read-back
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asuna_b200 import capi, scenes

which = sys.argv[1]
ctx = capi.Context(gpu_id=0)
if which == "smoke":
    sc = scenes.cornell(64, 64, spp=4, depth=4)
    sc.upload(ctx)
    img = sc.render_shot(ctx, 0)[0]
    print("smoke mean", float(img[..., :3].mean()))
elif which == "two_level":
    os.environ["ASUNA_FLATTEN"] = "0"
    sc = scenes.instanced_field(48, 32, spp=2, depth=4, subdiv=2, grid=3)
    sc.upload(ctx)
    print("field mean", float(sc.render_shot(ctx, 0)[0][..., :3].mean()), ctx.accel_stats())
    ctx.close()
    ctx = capi.Context(gpu_id=0)
    sc = scenes.cornell_all_materials(48, 36, spp=2, depth=4, env=True, lights="all", textured=True)
    sc.upload(ctx)
    print("materials mean", float(sc.render_shot(ctx, 0)[0][..., :3].mean()))
    ids, _ = ctx.trace_primary()
    print("primary hits", int((ids[..., 0] != 0xFFFFFFFF).sum()))
elif which == "build":
    sc = scenes.glass_blob(48, 32, spp=2, depth=4, subdiv=4, env_size=(16, 8))
    sc.upload(ctx)
    print("accel", ctx.accel_stats())
    sc.begin_shot(ctx, 0)
    ctx.render_frames(2)
    a, b = ctx.pinned_image(), ctx.pinned_image()
    ctx.read_channel_async(0, a)
    sc.begin_shot(ctx, 0)
    ctx.render_frames(2)
    ctx.read_channel_async(0, b)
    ctx.wait_reads()
    assert np.array_equal(a, b)
    print("build mean", float(a[..., :3].mean()))
elif which == "sort":
    rng = np.random.RandomState(1)
    keys = rng.randint(0, 2 ** 63, 100000, dtype=np.int64).astype(np.uint64)
    vals = np.arange(len(keys), dtype=np.uint32)
    k2, v2 = keys.copy(), vals.copy()
    rc = ctx.L.lib.asuna_debug_radix_sort(ctx.h, k2.ctypes.data_as(C.c_void_p), v2.ctypes.data_as(C.c_void_p), C.c_uint32(len(keys)))
    order = np.argsort(keys, kind="stable")
    assert rc == 0 and np.array_equal(k2, keys[order]) and np.array_equal(v2, vals[order])
    print("sort ok", len(keys))
ctx.close()
