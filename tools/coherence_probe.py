"""How much would re-ordering the ray queue buy?  Traces secondary-like rays (origins = primary hit points in
image order, directions uniform on the sphere) in several queue orders and prints the kernel rate
(ASUNA_TIME_USER_RAYS=1 makes asuna_trace_rays report its kernel time on stderr).
usage: ASUNA_TIME_USER_RAYS=1 python tools/coherence_probe.py [glass|rays]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asuna_b200 import capi, scenes

which = sys.argv[1] if len(sys.argv) > 1 else "glass"
W, H = 1920, 1080
sc = scenes.glass_blob(W, H, subdiv=6, env_size=(64, 32)) if which == "glass" else scenes.ray_bench(W, H, subdiv=8, depth=4)
ctx = capi.Context(gpu_id=0)
ctx.set_profiling(True)
sc.upload(ctx)
sc.begin_shot(ctx, 0)
shot = sc.shots[0]
eye, look, up = (np.asarray(v, np.float64) for v in (shot.eye, shot.lookat, shot.up))
f = look - eye
f /= np.linalg.norm(f)
r = np.cross(f, up)
r /= np.linalg.norm(r)
u = np.cross(r, f)
fov = np.deg2rad(45.0)
xs = (np.arange(W) + 0.5) / W * 2 - 1
ys = 1 - (np.arange(H) + 0.5) / H * 2
X, Y = np.meshgrid(xs, ys)
t = np.tan(fov / 2)
d = f[None, None] + X[..., None] * t * r + Y[..., None] * t * H / W * u
d /= np.linalg.norm(d, axis=2, keepdims=True)
rays = np.zeros((H * W, 8), np.float32)
rays[:, :3], rays[:, 4:7], rays[:, 3], rays[:, 7] = eye, d.reshape(-1, 3), 1e-5, 1e10
print("primary:", file=sys.stderr)
tuv, ip = ctx.trace_rays(rays)
hit = ip[:, 0] != 0xFFFFFFFF
print("primary hit fraction", hit.mean())
rng = np.random.RandomState(1)
o = rays[:, :3].astype(np.float64) + tuv[:, :1] * rays[:, 4:7]
v = rng.normal(size=(H * W, 3))
v /= np.linalg.norm(v, axis=1, keepdims=True)
sec = np.zeros_like(rays)
sec[:, :3], sec[:, 4:7], sec[:, 3], sec[:, 7] = o + 1e-3 * v, v, 1e-5, 1e10
sec = sec[hit]
n = len(sec)
octant = (sec[:, 4] < 0).astype(np.int64) | ((sec[:, 5] < 0).astype(np.int64) << 1) | ((sec[:, 6] < 0).astype(np.int64) << 2)
lo, hi = sec[:, :3].min(0), sec[:, :3].max(0)
q = np.clip(((sec[:, :3] - lo) / (hi - lo + 1e-9) * 1024).astype(np.int64), 0, 1023)


def spread(x):
    x = (x | (x << 16)) & 0x030000FF
    x = (x | (x << 8)) & 0x0300F00F
    x = (x | (x << 4)) & 0x030C30C3
    x = (x | (x << 2)) & 0x09249249
    return x


morton = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
# direction quantised on a 4x4x4 grid as a finer direction key
dq = np.clip(((sec[:, 4:7] + 1) * 2).astype(np.int64), 0, 3)
dkey = dq[:, 0] | (dq[:, 1] << 2) | (dq[:, 2] << 4)
orders = {
    "image order (today)": np.arange(n),
    "random shuffle": rng.permutation(n),
    "octant bins, stable": np.argsort(octant, kind="stable"),
    "octant bins per 64k-ray block": np.concatenate([b + np.argsort(octant[b:b + 65536], kind="stable") for b in range(0, n, 65536)]),
    "dir 4x4x4 bins, stable": np.argsort(dkey, kind="stable"),
    "octant, then morton(origin)": np.lexsort((morton, octant)),
    "morton(origin) >> 12, then octant": np.lexsort((octant, morton >> 12)),
}
for name, order in orders.items():
    print(f"{name}:", file=sys.stderr)
    for _ in range(3):
        ctx.trace_rays(np.ascontiguousarray(sec[order]))
