"""Occupancy of the compressed 8-wide BVH of a scene: children per node, inner / leaf children, primitives per leaf child.
usage: python tools/bvh_stats.py rays|glass|field"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from asuna_b200 import capi, scenes
import helpers

which = sys.argv[1] if len(sys.argv) > 1 else "glass"
sc = {"rays": lambda: scenes.ray_bench(64, 64, subdiv=8, depth=4), "glass": lambda: scenes.glass_blob(64, 64, subdiv=6, env_size=(16, 8)),
      "field": lambda: scenes.instanced_field(64, 64, subdiv=5, grid=10), "cornell": lambda: scenes.cornell(64, 64)}[which]()
ctx = capi.Context(gpu_id=0)
sc.upload(ctx)
nodes, tris, root = helpers.download_accel(ctx)
n_used = ctx.accel_stats()["nodes"]
meta = nodes["meta"][:n_used].astype(np.int64)
used = meta != 0
inner = used & ((meta & 31) >= 24)
leaf = used & ~inner
prims = np.where(leaf, np.array([bin(x).count("1") for x in range(8)])[(meta >> 5) & 7], 0)
print(json.dumps({"scene": which, "nodes": int(n_used), "children_per_node": float(used.sum(1).mean()),
                  "hist_children": np.bincount(used.sum(1), minlength=9).tolist(), "inner_per_node": float(inner.sum(1).mean()),
                  "leaf_children_per_node": float(leaf.sum(1).mean()), "prims_per_leaf_child": float(prims.sum() / max(leaf.sum(), 1)),
                  "hist_leaf_prims": np.bincount(prims[leaf], minlength=4).tolist()}))
