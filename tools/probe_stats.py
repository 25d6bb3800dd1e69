"""Per hit kind: how the CUDA shade kernels compare with the frozen reference-GLSL probes (tests/golden/ref_probes.npz) and how
many probes continue / queue a shadow ray / write AOVs -- a sanity view of tests/test_gpu_ref_parity.py on the GPU box."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from asuna_b200 import capi
from tools.make_ref_golden import probe_cases
from test_gpu_ref_parity import gpu_probes, check_gpu_probes
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_probes.npz"))
for key, sc, args in probe_cases():
    ctx = capi.Context(gpu_id=0); sc.upload(ctx); sc.begin_shot(ctx, 0)
    G = gpu_probes(ctx, *args); R = g[key].view(H.PROBE).reshape(-1)
    bad = check_gpu_probes(G, R, key)
    print(f"{key:16s} cont {float((R['stop']==0).mean()):.2f} nee_gpu {float((G['drec_skip']==0).mean()):.2f} nee_ref {float((R['drec_skip']==0).mean()):.2f} depth1 {float((R['depth']==1).mean()):.2f} chan_nonzero {float((np.abs(G['channel']).sum(axis=(1,2))>0).mean()):.2f} worst {max(bad.values()):.4f}")
    ctx.close()
