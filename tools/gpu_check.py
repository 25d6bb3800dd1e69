"""Ad-hoc GPU bring-up check: CUDA path vs CPU oracle on one scene.  Usage: python tools/gpu_check.py [scene] [w] [h] [spp]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from asuna_b200 import capi, scenes
from oracle.binding import OracleContext

name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 256
h = int(sys.argv[3]) if len(sys.argv) > 3 else 256
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 16
if name == "cornell":
    sc = scenes.cornell(w, h, spp=spp)
elif name == "materials":
    sc = scenes.cornell_materials(w, h, spp=spp, env=False, lights="all", textured=True)
elif name == "materials_env":
    sc = scenes.cornell_materials(w, h, spp=spp, env=True, lights="rect", textured=True)
elif name == "glass":
    sc = scenes.glass_blob(w, h, spp=spp, subdiv=4, env_size=(256, 128))
elif name == "pbr":
    sc = scenes.pbr_spheres(w, h, spp=spp, subdiv=4, tex_size=128)
elif name == "field":
    sc = scenes.instanced_field(w, h, spp=spp, subdiv=3, grid=4)
else:
    raise SystemExit("unknown scene")
gpu = capi.Context(gpu_id=0)
cpu = OracleContext()
print("gpu build ms", sc.upload(gpu), "accel", gpu.accel_stats())
print("cpu build ms", sc.upload(cpu))
sc.begin_shot(gpu, 0); sc.begin_shot(cpu, 0)
ig, tg = gpu.trace_primary(); ic, tc = cpu.trace_primary()
agree = (ig == ic).all(axis=2)
print("primary agreement", agree.mean(), "max |dt| where agree", np.abs(tg - tc)[agree].max())
t = time.time(); g = sc.render_shot(gpu, 0); tgpu = time.time() - t
t = time.time(); c = sc.render_shot(cpu, 0); tcpu = time.time() - t
print("gpu s", tgpu, gpu.stats()); print("cpu s", tcpu, cpu.stats(), cpu.traversal_counters())
for k, (a, b) in enumerate(zip(g, c)):
    d = np.abs(a[..., :3] - b[..., :3])
    print("channel", k, "max abs diff", d.max(), "frac>1e-4", (d.max(axis=2) > 1e-4).mean(), "mean", a[..., :3].mean(), b[..., :3].mean())
rel = np.abs(g[0][..., :3] - c[0][..., :3]).mean() / c[0][..., :3].mean()
print("radiance mean abs diff / mean", rel)
np.save("gpurun_out/check_gpu.npy", g[0]); np.save("gpurun_out/check_cpu.npy", c[0])
