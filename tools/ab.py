"""A/B harness for kernel variants on the GPU box: runs tools/ray_bench.py for each (library, ASUNA_TUNE) pair and
prints one compact line per run.  usage: python tools/ab.py lib1[@tune] lib2[@tune] ... [-- scene:subdiv:frames ...]"""
import json
import os
import subprocess
import sys

args = sys.argv[1:]
scenes = ["glass:6:16", "rays:8:16", "field:6:8"]
if "--" in args:
    k = args.index("--")
    args, scenes = args[:k], args[k + 1:]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in args:
    lib, _, tune = spec.partition("@")
    path = lib if os.path.exists(lib) else os.path.join(root, "build", "variants", f"lib_{lib}.so")
    env = dict(os.environ, ASUNA_B200_LIB=os.path.abspath(path))
    if tune:
        env["ASUNA_TUNE"] = tune
    out = []
    for sc in scenes:
        name, subdiv, frames = sc.split(":")
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "ray_bench.py"), subdiv, frames, name],
                           env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            out.append(f"{name}: closest {d['closest_Mrays_s']:.0f} incoh {d['incoherent_closest_Mrays_s']:.0f} "
                       f"shadow {d['shadow_Mrays_s']:.0f} samples {d['samples_per_s_M']:.1f}M")
        except Exception:
            out.append(f"{name}: FAILED {r.stderr[-200:]}")
    print(f"{spec:28s} | " + " | ".join(out), flush=True)
