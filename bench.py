#!/usr/bin/env python
"""Benchmark of the asuna_b200 hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (CPU oracle on the host cores, same config)

Metric (BASELINE.json): 1080p samples/sec (pixel-samples per second) on config[1] -- the glass blob:
dielectric BSDF + env-map light, 1920x1080, max path depth 8.  One *step* is one pass of the hot path
over one batch of `--frames-per-step` 1-spp frames per GPU (raygen -> [trace, shade, shadow] x depth ->
accumulate).  Multi-GPU is the sample-range split of SURVEY.md 8e: rank r renders the frames
f % N == r of the shot (weak scaling: frames per GPU fixed), the (sum w*L, sum w) planes are summed to
rank 0 with one NCCL reduce and resolved there.

`value`   : device-resident throughput -- scene + BVH already in HBM, K steps bracketed by barrier +
            synchronize, CUDA events on the library's stream, max over ranks (+ the one NCCL reduce at N>1).
`e2e`     : the same metric through the reference-facing C ABI with HOST buffers: every step uploads the
            step's camera / state / sun-sky structs from host memory, renders, resolves (NCCL reduce at
            N > 1) and reads the radiance image back into page-locked host memory (a complete mini-shot); the copy
            of step k runs on the library's read stream while step k + 1 renders (asuna_read_channel_async) and
            the last one is awaited before the clock stops.
`roofline`: closest-hit trace kernel (dominant): algorithmic bytes per ray (32 B ray in +
            16 B hit out + visited nodes x 80 B + tested triangles x 48 B, counts from an untimed
            instrumented pass over the same BVH) x rays per launch / mean launch time (CUDA events).  The contract's
            fraction is against the HBM copy peak; the kernel's working set is L2-resident and what binds it is SM issue
            / the ALU pipe, so `issue_frac`, `alu_pipe_frac` and `lanes_per_inst` from the committed single-launch ncu
            capture (profiles/traffic.json) are carried beside it.
`multi_gpu_check` (N > 1): rank 0 renders the same frames alone and compares with the NCCL-reduced image.
`time_to_image_s`: the WHOLE 256-spp job of BASELINE configs[1] at N ranks (strong scaling): context creation, scene
            upload, BVH build, 256 / N frames per rank, NCCL reduce, resolve, read-back into host memory.
`ray_bench` (rank 0): the north star's other two figures -- incoherent closest-hit Mrays/s on the 1.31 M-triangle C4'
            scene with nodes / triangles per ray, and BVH build ms at 1.31 M and 32.8 M triangles.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "1080p samples/sec"
UNIT = "samples/s"
WORKLOAD = "glass blob (dielectric + env-map), 1920x1080, depth 8 [BASELINE.json configs[1]]"
NODE_BYTES, TRI_BYTES, RAY_IN_BYTES, HIT_OUT_BYTES = 80, 48, 32, 16


WORKLOAD_C4 = "instanced field: 100 instances of a 327 k-triangle mesh, mixed BSDFs, env map + rect light, 3840x2160, depth 5 [BASELINE.json configs[3]]"


def build_scene(args):
    from asuna_b200 import scenes
    if args.workload == "field4k":  # BASELINE.json configs[3]: the multi-GPU scaling scene (not the default bench line)
        return scenes.instanced_field(3840, 2160, spp=1024, depth=5, subdiv=7, grid=10)
    return scenes.glass_blob(args.width, args.height, spp=256, depth=8, subdiv=args.subdiv, env_size=(2048, 1024))


def n_triangles(args):
    if args.workload == "field4k":
        return 100 * 20 * 4 ** 7 + 2 * 8 * 8
    return 20 * 4 ** args.subdiv + 2 * 128 * (2 * 65 - 1) - 2 * 128 + 128  # blob + lathe bowl (minus pole slivers) + ground


def config_dict(args, n_gpus):
    if args.workload == "field4k":
        return {"workload": WORKLOAD_C4, "width": 3840, "height": 2160, "max_path_depth": 5, "triangles": n_triangles(args),
                "frames_per_step_per_gpu": args.frames_per_step, "parallelism": f"sample-range split x{n_gpus}, scene replicated",
                "cache_policy": "inputs larger than L2: 2 GB flattened BVH + path state"}
    return {"workload": WORKLOAD, "width": args.width, "height": args.height, "max_path_depth": 8,
            "triangles": n_triangles(args), "frames_per_step_per_gpu": args.frames_per_step,
            "parallelism": f"sample-range split x{n_gpus}, scene replicated",
            "cache_policy": "inputs larger than L2: per-step path state ~%.1f GB per GPU" %
                            (args.width * args.height * args.frames_per_step * 148 / 1e9)}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_capture():
    """Numbers of the dominant kernel from the committed single-launch `ncu --set full` capture (profiles/traffic.json:
    DRAM bytes per launch, issue-slot and ALU-pipe fractions, live lanes per instruction), if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def ncu_traffic():
    return ncu_capture().get("k_trace_closest_dram_bytes_per_launch")


# ----------------------------------------------------------------------------- CPU arm
def cpu_run(args, steps, warmup, budget_s=None):
    """Times the reference's per-pixel program on the host cores, same workload; a step is one 1-spp frame.
    kind "reference": oracle/_ref/libref.so -- the reference's own GLSL (rgen, rmiss, all rchit, utils) compiled as
    C++ against the GLM vendored in its tree (oracle/refbuild/build_ref.py), with the oracle's SAH BVH2 answering
    traceRayEXT (the reference leaves that to the Vulkan driver and cannot run without an RT device).
    kind "port": oracle/liboracle.so, the hand restatement, when libref.so was not built."""
    from oracle import binding
    sc = build_scene(args)
    kind = "reference" if os.path.exists(binding.REF_LIB) else "port"
    ctx = binding.RefContext() if kind == "reference" else binding.OracleContext()
    cores = os.cpu_count() or 1
    sc.upload(ctx)
    sc.begin_shot(ctx, 0)
    for _ in range(warmup):
        ctx.render_frames(1)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        ctx.render_frames(1)
        done += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    samples = done * args.width * args.height
    st = ctx.stats()
    ctx.close()
    what = "reference GLSL compiled as C++ (GLM) + oracle BVH2" if kind == "reference" else "CPU oracle port"
    return {"value": samples / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{done} x 1-spp {args.width}x{args.height} frames of the same scene ({dt:.1f} s, BVH build excluded); {what}",
            "ms_per_step": 1e3 * dt / max(done, 1), "steps": done,
            "mrays_per_s": (st["closest_rays"] + st["shadow_rays"]) / max(st["total_ms"], 1e-9) / 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_run(args, max(args.steps, 1), args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, args.gpus),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mrays_per_s": r["mrays_per_s"], "gpu_launches": 0,
            "note": "the reference's per-pixel program on all host threads (see cpu_baseline.kind / sample); one step = one 1-spp frame"}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="asuna_b200", choices=["asuna_b200", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--subdiv", type=int, default=6)
    ap.add_argument("--frames-per-step", type=int, default=8)
    ap.add_argument("--workload", default="glass", choices=["glass", "field4k"],
                    help="glass = BASELINE configs[1] (the bench line); field4k = configs[3], for scaling runs of the instanced 4K scene")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ray-bench", action="store_true", help="skip the C4' ray bench / 32.8 M-triangle build / time-to-image extras")
    ap.add_argument("--tti-spp", type=int, default=256, help="samples per pixel of the time-to-image job (BASELINE configs[1]: 256)")
    ap.add_argument("--cpu-budget-s", type=float, default=12.0)
    args = ap.parse_args()
    if args.workload == "field4k":
        args.width, args.height = 3840, 2160
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    # stdout carries exactly one JSON line: native libraries (NCCL prints its version banner there under
    # NCCL_DEBUG=VERSION/WARN/INFO) get stderr as their fd 1 for the whole run, the JSON goes to the saved descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from asuna_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    sc = build_scene(args)
    ctx = capi.Context(gpu_id=local)  # raises without the CUDA library / a device: there is no fallback
    ctx.set_profiling(True)  # per-kernel CUDA-event times feed the roofline block (the e2e leg below runs without)
    t0 = time.perf_counter()
    build_ms = sc.upload(ctx)
    upload_s = time.perf_counter() - t0
    ctx.set_partition(rank, world)
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
    n_px = args.width * args.height
    fps = args.frames_per_step
    global_frames_per_step = fps * world  # every rank renders fps of them

    class _Ptr:
        def __init__(self, p, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (p, False), "version": 3}

    def resolve():
        """Multi-GPU combine: one NCCL reduce of the (sum w*L, sum w) plane to rank 0, then L = sum/w."""
        if world == 1:
            return
        # no host synchronisation: export / import are ordered on the library's stream, the reduce on torch's; each
        # side waits for the other through stream events (wait_stream)
        t = torch.as_tensor(_Ptr(ctx.export_partial(), n_px * 4), device=dev)
        lib_stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(lib_stream)
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        lib_stream.wait_stream(cur)
        if rank == 0:
            ctx.import_partial()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed instrumented pass: mean nodes / triangles per closest-hit ray on this BVH
    sc.begin_shot(ctx, 0)
    ctx.set_counting(True)
    ctx.reset_stats()
    ctx.render_frames(world)  # one frame per rank
    s = ctx.stats()
    nodes_per_ray = s["node_visits"] / max(s["closest_rays"], 1)
    tris_per_ray = s["tri_tests"] / max(s["closest_rays"], 1)
    bytes_per_ray = RAY_IN_BYTES + HIT_OUT_BYTES + nodes_per_ray * NODE_BYTES + tris_per_ray * TRI_BYTES
    ctx.set_counting(False)

    # ---- device-resident throughput
    sc.begin_shot(ctx, 0)
    for _ in range(args.warmup):
        ctx.render_frames(global_frames_per_step)
    ctx.sync()
    ctx.reset_stats()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record()
    for _ in range(args.steps):
        ctx.render_frames(global_frames_per_step)
    with torch.cuda.stream(stream):
        ev1.record()
    ctx.sync()
    resolve()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - w0)
    dev_ms = ev0.elapsed_time(ev1)
    st = ctx.stats()
    t = torch.tensor([wall_ms, dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_ms, dev_ms = t.tolist()
    samples = args.steps * global_frames_per_step * n_px
    value = samples / (wall_ms / 1e3)

    # ---- end to end through the C ABI with host buffers (mini-shot per step)
    cam_host = sc.gpu_camera(sc.shots[0])
    h2d = cam_host.nbytes + sc.shot_state(0).nbytes + sc.sunsky.nbytes
    d2h = n_px * 16
    # two page-locked read-back buffers (asuna_host_alloc), used in turn: the copy of step k travels on the context's
    # read stream while step k + 1 renders (asuna_read_channel_async); the last one is awaited inside the timed region
    imgs = [ctx.pinned_image(), ctx.pinned_image()] if rank == 0 else None
    ctx.set_profiling(False)  # what an integration gets by default: no per-launch event records
    for _ in range(2):
        sc.begin_shot(ctx, 0)
        ctx.render_frames(global_frames_per_step)
        resolve()
    barrier()
    e0 = time.perf_counter()
    for k in range(args.steps):
        sc.begin_shot(ctx, 0)  # camera + state + sun/sky structs from host memory
        ctx.render_frames(global_frames_per_step)
        resolve()
        if rank == 0:
            ctx.read_channel_async(0, imgs[k & 1])  # radiance image back into (pinned) host memory
    if rank == 0:
        ctx.wait_reads()
    ctx.sync()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - e0)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()
    e2e_value = samples / (e2e_ms / 1e3)
    clock_info = clocks.stop() if rank == 0 else None

    # ---- multi-GPU correctness through the real NCCL path: the reduced image of N partitions against the same
    # frames rendered by rank 0 alone (same seeds per (pixel, frame); only the fp32 summation order differs)
    multi_gpu_check = None
    if world > 1:
        sc.begin_shot(ctx, 0)
        ctx.render_frames(global_frames_per_step)
        resolve()
        if rank == 0:
            reduced = ctx.read_channel(0).copy()
            ctx.set_partition(0, 1)
            sc.begin_shot(ctx, 0)
            ctx.render_frames(global_frames_per_step)
            ctx.sync()
            alone = ctx.read_channel(0)
            ctx.set_partition(rank, world)
            a, b = reduced[..., :3].astype(np.float64), alone[..., :3].astype(np.float64)
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
            multi_gpu_check = {"frames": global_frames_per_step, "max_rel_err": float(rel.max()), "rtol": 3e-5,
                               "mean_reduced": float(a.mean()), "mean_single_gpu": float(b.mean()),
                               "pass": bool(rel.max() <= 3e-5)}
            if not multi_gpu_check["pass"]:
                raise SystemExit(f"multi-GPU image differs from the single-GPU image: {multi_gpu_check}")
        barrier()

    # ---- time to image: the whole 256-spp job of BASELINE configs[1] at `world` ranks (strong scaling), everything a
    # user waits for after the scene description is in host memory: context, upload, BVH build, 256 / N frames per
    # rank, reduce + resolve, read-back
    tti = None
    if not args.no_ray_bench and args.workload == "glass":
        ctx.close()
        # the scene description waits in PAGE-LOCKED host memory (as the e2e inputs do): the 100 MB of env-map texels and
        # importance tables and the meshes then travel as plain DMA on the library's upload stream, under the BVH build
        keep = []

        def pin(a):
            t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
            keep.append(t)
            b = t.numpy().view(a.dtype).reshape(a.shape)
            b[...] = a
            return b
        if sc.envmap is not None:
            sc.envmap = tuple(pin(np.ascontiguousarray(a, np.float32)) for a in sc.envmap)
        sc.meshes = [(pin(v), pin(i)) for v, i in sc.meshes]
        # the whole job three times, the median reported (every run listed): a run right after a 3 GB context was torn down
        # occasionally pays tens of milliseconds inside the driver (context creation, the first copies) that a fresh
        # process does not
        runs = []
        for rep in range(3):
            if rep:
                ctx.close()
            barrier()
            t0 = time.perf_counter()
            ctx = capi.Context(gpu_id=local)
            t1 = time.perf_counter()
            tti_build_ms = sc.upload(ctx)
            ctx.set_partition(rank, world)
            t2 = time.perf_counter()
            sc.begin_shot(ctx, 0)
            ctx.render_frames(args.tti_spp)  # asynchronous: the launches are queued, the host goes on
            img = ctx.pinned_image() if rank == 0 else None  # page-locked read-back buffer, allocated while the GPU renders
            ctx.sync()
            t3 = time.perf_counter()
            resolve()
            if rank == 0:
                ctx.read_channel(0, out=img)
            t4 = time.perf_counter()
            barrier()
            t5 = time.perf_counter()
            tt = torch.tensor([t5 - t0, t1 - t0, t2 - t1, t3 - t2, t4 - t3], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            runs.append((tt.tolist(), tti_build_ms))
        tt, tti_build_ms = sorted(runs, key=lambda r: r[0][0])[1]
        tti = {"value": tt[0], "unit": "s", "spp": args.tti_spp, "scaling": "strong",
               "create_context_s": tt[1], "upload_and_build_s": tt[2], "bvh_build_ms": tti_build_ms, "render_s": tt[3],
               "reduce_resolve_readback_s": tt[4],
               "samples_per_s": args.tti_spp * n_px / tt[0], "runs_s": [r[0][0] for r in runs], "reported": "median of 3"}

    # ---- the north star's other figures, rank 0 only: C4' (1.31 M triangles) incoherent rays/s, BVH build at 1.31 M and
    # 32.8 M triangles
    ray_bench = None
    if not args.no_ray_bench and rank == 0:
        from asuna_b200 import scenes
        ctx.close()
        rb = scenes.ray_bench(1920, 1080, subdiv=8, depth=4)
        c2 = capi.Context(gpu_id=local)
        c2.set_profiling(True)
        rb_build_ms = rb.upload(c2)
        rb.begin_shot(c2, 0)
        c2.set_counting(True)
        c2.render_frames(1)
        s1 = c2.stats()
        c2.set_counting(False)
        rb.begin_shot(c2, 0)
        c2.render_frames(8)
        c2.sync()
        c2.reset_stats()
        c2.render_frames(16)
        s2 = c2.stats()
        ray_bench = {"scene": "C4': 1.31 M-triangle blob + ground, lambertian, white environment, 1920x1080, depth 4, 16 frames",
                     "triangles": int(sum(len(i) // 3 for _, i in rb.meshes)),
                     "incoherent_mrays_per_s": s2["incoherent_closest_rays"] / max(s2["closest_ms"], 1e-9) / 1e3 *
                                               1.0,
                     "closest_mrays_per_s": s2["closest_rays"] / max(s2["closest_ms"], 1e-9) / 1e3,
                     "all_mrays_per_s": (s2["closest_rays"] + s2["shadow_rays"]) / max(s2["trace_ms"], 1e-9) / 1e3,
                     "samples_per_s": s2["paths"] / max(s2["total_ms"], 1e-9) * 1e3,
                     "nodes_per_ray": s1["node_visits"] / max(s1["closest_rays"], 1),
                     "tris_per_ray": s1["tri_tests"] / max(s1["closest_rays"], 1),
                     "bvh_build_ms": rb_build_ms,
                     "note": "incoherent = closest-hit rays at depth >= 2 (after a cosine-hemisphere bounce) over the closest-hit kernel's device time"}
        c2.close()
        del rb
        fld = scenes.instanced_field(256, 144, spp=1, depth=5, subdiv=7, grid=10)
        c3 = capi.Context(gpu_id=local)
        ray_bench["bvh_build_ms_32M"] = fld.upload(c3)
        ray_bench["triangles_32M"] = int(sum(len(fld.meshes[m][1]) // 3 for _, m, _, _ in fld.instances))
        c3.close()
        del fld
    barrier()

    if rank == 0:
        peak, peak_src = measured_peak()
        cap = ncu_capture()
        rays_per_launch = st["closest_rays"] / max(st["closest_launches"], 1)
        launch_ms = st["closest_ms"] / max(st["closest_launches"], 1)
        achieved = bytes_per_ray * rays_per_launch / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
        rays = st["closest_rays"] + st["shadow_rays"]
        line = {
            "metric": METRIC if args.workload == "glass" else "2160p samples/sec", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args, world),
            "device_ms_per_step": dev_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(st["kernel_launches"]),
            "clocks": clock_info,
            "roofline": {"kernel": "k_trace_closest", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                         # what actually binds this kernel (its working set is L1/L2-resident): from the committed
                         # single-launch ncu --set full capture of this build
                         "actual_bound": "sm_issue / alu_pipe", "issue_frac": cap.get("k_trace_closest_issue_frac"),
                         "alu_pipe_frac": cap.get("k_trace_closest_alu_pipe_frac"),
                         "lanes_per_inst": cap.get("k_trace_closest_lanes_per_inst"), "ncu_source": cap.get("source"),
                         "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                         "rays_per_launch": rays_per_launch, "launch_ms": launch_ms,
                         "kernel_share_of_step": st["closest_ms"] / max(st["total_ms"], 1e-9),
                         "note": "BVH + triangles (~%.0f MB) %s; the HBM fraction is reported as the contract asks" %
                                 (n_triangles(args) * (0.16 * NODE_BYTES + TRI_BYTES) / 1e6,
                                  "are L2-resident, so this kernel is latency/issue-bound" if args.workload == "glass"
                                  else "exceed the 126 MB L2: top levels from L1/L2, leaves from HBM")},
            "mrays_per_s_rank0": rays / max(st["trace_ms"], 1e-9) / 1e3,
            "incoherent_mrays_per_s_rank0": st["incoherent_closest_rays"] / max(st["closest_ms"], 1e-9) / 1e3,
            "rays_per_sample": rays / max(st["paths"], 1),
            "bvh_build_ms": build_ms, "scene_upload_s": upload_s,
            "kernel_ms_rank0": {k: st[k] for k in ("closest_ms", "shadow_ms", "shade_ms", "total_ms")},
        }
        if multi_gpu_check is not None:
            line["multi_gpu_check"] = multi_gpu_check
        if tti is not None:
            line["time_to_image_s"] = tti["value"]
            line["time_to_image"] = tti
        if ray_bench is not None:
            line["ray_bench"] = ray_bench
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_run(args, 64, 1, budget_s=args.cpu_budget_s)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    ctx.close()  # idempotent
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
