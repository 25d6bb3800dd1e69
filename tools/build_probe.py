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
print(json.dumps({"scene": which, "triangles": int(sum(len(sc.meshes[m][1]) // 3 for _, m, _, _ in sc.instances)), "build_ms": ms,
                  "accel": ctx.accel_stats()}))
