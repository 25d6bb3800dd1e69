import sys, os, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import helpers as H
from asuna_b200 import capi
from tools.make_ref_golden import probe_cases
from test_gpu_ref_parity import gpu_probes, check_gpu_probes
g = np.load("/root/repo/tests/golden/ref_probes.npz")
for key, sc, args in probe_cases():
    ctx = capi.Context(gpu_id=0); sc.upload(ctx); sc.begin_shot(ctx, 0)
    G = gpu_probes(ctx, *args); R = g[key].view(H.PROBE).reshape(-1)
    bad = check_gpu_probes(G, R, key)
    print(f"{key:16s} cont {float((R['stop']==0).mean()):.2f} nee_gpu {float((G['drec_skip']==0).mean()):.2f} nee_ref {float((R['drec_skip']==0).mean()):.2f} depth1 {float((R['depth']==1).mean()):.2f} chan_nonzero {float((np.abs(G['channel']).sum(axis=(1,2))>0).mean()):.2f} worst {max(bad.values()):.4f}")
    ctx.close()
