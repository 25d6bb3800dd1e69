"""World-size-2 `gloo` test of the multi-GPU sample-range split: rank r renders the
frames f with f % world == r, the (sum w*L, sum w) planes are summed to rank 0 and resolved there.
Runs on the CPU oracle through the same ABI calls bench.py uses on GPUs (partition / export / import)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import ctypes as C
    from asuna_b200 import scenes
    from oracle.binding import OracleContext
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.cornell(24, 20, spp=6, depth=4)
    ctx = OracleContext(threads=1)
    sc.upload(ctx)
    ctx.set_partition(rank, world)
    tot = sc.begin_shot(ctx, 0)
    ctx.render_frames(tot)
    ptr = ctx.export_partial()
    n = 24 * 20 * 4
    buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,))
    t = torch.from_numpy(buf)  # shares memory with the library-owned partial buffer
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        ctx.import_partial()
        np.save(out_path, ctx.read_channel(0))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_frame_split_equals_single_rank(tmp_path, cpu_ctx):
    from asuna_b200 import scenes
    out = str(tmp_path / "combined.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    combined = np.load(out)
    sc = scenes.cornell(24, 20, spp=6, depth=4)
    sc.upload(cpu_ctx)
    single = sc.render_shot(cpu_ctx, 0)[0]
    # identical sample set, different fp32 summation order
    assert np.allclose(combined[..., :3], single[..., :3], rtol=2e-5, atol=2e-6)
    assert np.allclose(combined[..., 3], 1.0)
