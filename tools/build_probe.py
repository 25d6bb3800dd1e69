"""Builds the acceleration structure of one scene and prints asuna_build_accel's device time (use under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list).  usage: python tools/build_probe.py rays|field|glass [repeats]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asuna_b200 import capi, scenes

which = sys.argv[1] if len(sys.argv) > 1 else "rays"
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = {"rays": lambda: scenes.ray_bench(64, 64, subdiv=8, depth=4), "glass": lambda: scenes.glass_blob(64, 64, subdiv=6, env_size=(16, 8)),
      "field": lambda: scenes.instanced_field(64, 64, subdiv=7, grid=10)}[which]()
ctx = capi.Context(gpu_id=0)
ms = [sc.upload(ctx)]
for _ in range(rep - 1):
    ms.append(ctx.build_accel())
if ctx.L.has("debug_build_profile"):  # a -DASUNA_BUILD_PROFILE build: phase stamps of the LAST build
    import ctypes
    import numpy as np
    buf = np.zeros((2048, 2), np.uint64)
    k = ctx.L.fn("debug_build_profile")(buf.ctypes.data_as(ctypes.c_void_p), 2048)
    buf = buf[:k]
    # the reader drains the log: keep the stamps of the last build only (tag 1 = PLOC start)
    tags, vals, t = (buf[:, 1] >> np.uint64(32)).astype(int), (buf[:, 1] & np.uint64(0xFFFFFFFF)).astype(int), buf[:, 0].astype(np.int64)
    starts = np.flatnonzero(tags == 1)
    start = int(starts[vals[starts] == vals[starts].max()][-1])  # the largest BVH of the last build
    first = [int(x) for x in np.flatnonzero((tags == 20) | (tags == 21)) if x < start]  # ... from its first kernel on
    start = first[-1] if first else start
    while start > 0 and tags[start - 1] == 20:
        start -= 1
    k = next((int(x) for x in np.flatnonzero(tags == 28) if x > start), k - 1) + 1
    names = {1: "ploc init", 5: "tail start", 2: "nn", 3: "flags+scan", 4: "merge", 8: "emit init", 9: "emit level", 10: "emit prims", 20: "K world", 21: "K morton", 22: "K hist", 23: "K scan", 24: "K scatter", 25: "K ploc", 26: "K tail", 27: "K emit", 28: "end"}
    for i in range(start, k):
        dt = (t[i] - t[i - 1]) / 1e3 if i > start else 0.0
        print(f"{names[tags[i]]:12s} {vals[i]:9d} {dt:9.1f} us", file=sys.stderr)
print(json.dumps({"scene": which, "triangles": int(sum(len(sc.meshes[m][1]) // 3 for _, m, _, _ in sc.instances)), "build_ms": ms,
                  "accel": ctx.accel_stats()}))
