"""Per-kernel differential tests: radix sort bit-exact, traversal against the
oracle on random rays, any-hit occlusion, watertightness on shared edges / vertices, BVH build stats."""
import ctypes as C

import numpy as np
import pytest

from asuna_b200 import host, scenes, structs as S

pytestmark = pytest.mark.gpu


def _sort(ctx, keys, vals):
    k = np.ascontiguousarray(keys, np.uint64).copy()
    v = np.ascontiguousarray(vals, np.uint32).copy()
    f = ctx.L.lib.asuna_debug_radix_sort
    f.restype = C.c_int
    rc = f(ctx.h, k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_uint32(k.size))
    assert rc == 0
    return k, v


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 4095, 4096, 4097, 100003, 1 << 20])
def test_radix_sort_is_exact_and_stable(gpu_ctx, n):
    rng = np.random.RandomState(n)
    keys = rng.randint(0, 1 << 62, size=n, dtype=np.int64).astype(np.uint64)
    keys[rng.rand(n) < 0.3] &= np.uint64(0xFF)  # many duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    k, v = _sort(gpu_ctx, keys, vals)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


def test_radix_sort_edge_patterns(gpu_ctx):
    n = 10000
    for keys in (np.zeros(n, np.uint64), np.full(n, 0xFFFFFFFFFFFFFFFF, np.uint64), np.arange(n, dtype=np.uint64)[::-1].copy(),
                 (np.arange(n, dtype=np.uint64) << np.uint64(56))):
        k, v = _sort(gpu_ctx, keys, np.arange(n, dtype=np.uint32))
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k, keys[order]) and np.array_equal(v, order.astype(np.uint32))


def _scene_bounds(sc):
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for x, mesh, _, _ in sc.instances:
        p = sc.meshes[mesh][0]["pos"].astype(np.float64)
        w = p @ np.asarray(x, np.float64)[:3, :3].T + np.asarray(x, np.float64)[:3, 3]
        lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
    return lo, hi


def _random_rays(sc, n, seed):
    """Origins in the scene box inflated by 50 %, aimed at random points inside the box."""
    lo, hi = _scene_bounds(sc)
    c, e = 0.5 * (lo + hi), 0.5 * (hi - lo)
    rng = np.random.RandomState(seed)
    o = c + 1.5 * e * (2 * rng.rand(n, 3) - 1)
    tgt = c + e * (2 * rng.rand(n, 3) - 1)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, :3], rays[:, 4:7] = o, d
    rays[:, 3], rays[:, 7] = 1e-5, 1e10
    return rays


def _mixed_scene():
    """An instanced mesh next to several single-use meshes: the world BLAS (flattened singles) and per-mesh
    BLASes meet under one top level."""
    sc = scenes.instanced_field(32, 32, subdiv=3, grid=4)
    v, idx = scenes.blob(3, 5, 0.3)
    sc.add_mesh("solo", v, idx)
    sc.add_instance("solo", "lam", scenes.translation((0.3, 3.0, 0.2)) @ scenes.rotation_y(0.7) @ scenes.scaling((1.5, 0.8, 1.2)))
    return sc


SCENE_BUILDERS = [
    lambda: scenes.cornell_materials(32, 32, env=False, lights="all"),  # all single-use: single-level traversal
    lambda: scenes.instanced_field(32, 32, subdiv=3, grid=5),           # instanced + one single: plain two-level
    lambda: scenes.glass_blob(32, 32, subdiv=5, env_size=(16, 8)),
    _mixed_scene,                                                       # instanced + world BLAS
]


# acceleration-structure policies of asuna_build_accel: plain two-level / single-use meshes flattened (instanced ones
# keep their BLAS under the instance level next to the world BLAS) / everything flattened (default under the budget)
ACCEL_POLICIES = {"two-level": {"ASUNA_FLATTEN": "0"}, "singles": {"ASUNA_FLATTEN": "1", "ASUNA_FLATTEN_MAX_TRIS": "0"},
                  "all": {"ASUNA_FLATTEN": "1"}}


def _set_policy(monkeypatch, name):
    monkeypatch.delenv("ASUNA_FLATTEN", raising=False)
    monkeypatch.delenv("ASUNA_FLATTEN_MAX_TRIS", raising=False)
    for k, v in ACCEL_POLICIES[name].items():
        monkeypatch.setenv(k, v)


@pytest.mark.parametrize("builder", SCENE_BUILDERS)
def test_flattening_does_not_change_hits(product_lib, builder, monkeypatch):
    """Instances are pre-transformed into one world-space BLAS (api.cu); against the plain two-level structure
    (ASUNA_FLATTEN=0) the nearest hits must be the same primitives at the same distance, whichever policy applies."""
    from asuna_b200 import capi
    sc = builder()
    rays = _random_rays(sc, 100000, 9)
    out = {}
    for name in ACCEL_POLICIES:
        _set_policy(monkeypatch, name)
        ctx = capi.Context(product_lib, 0)
        sc.upload(ctx)
        out[name] = ctx.trace_rays(rays) + (ctx.occlusion_rays(rays), ctx.accel_stats())
        ctx.close()
    t0, i0, o0, a0 = out["two-level"]
    assert a0["tlas_nodes"] >= 1
    for name in ("singles", "all"):
        t1, i1, o1, _ = out[name]
        same = (i0 == i1).all(axis=1)
        tie = ~same & (np.abs(t0[:, 0] - t1[:, 0]) <= 1e-5 * np.maximum(1.0, np.abs(t0[:, 0])))
        assert (same | tie).mean() >= 0.9995, (name, same.mean(), tie.mean())
        hit = same & (i0[:, 0] != 0xFFFFFFFF)
        assert hit.mean() > 0.2
        assert (np.abs(t0[hit, 0] - t1[hit, 0]) / np.maximum(1.0, t0[hit, 0])).max() <= 1e-5
        assert (o0 == o1).mean() >= 0.9995
    n_inst_tris = sum(len(sc.meshes[m][1]) // 3 for _, m, _, _ in sc.instances)
    assert out["all"][3]["leaf_prims"] == n_inst_tris  # one world-space copy per instance


@pytest.mark.parametrize("policy", ["all", "singles"])
@pytest.mark.parametrize("builder", SCENE_BUILDERS)
def test_traversal_matches_oracle_on_random_rays(gpu_ctx, cpu_ctx, builder, policy, monkeypatch):
    _set_policy(monkeypatch, policy)  # read by asuna_build_accel (sc.upload)
    sc = builder()
    sc.upload(gpu_ctx)
    sc.upload(cpu_ctx)
    rays = _random_rays(sc, 200000, 5)
    tg, ig = gpu_ctx.trace_rays(rays)
    tc, ic = cpu_ctx.trace_rays(rays)
    same = (ig == ic).all(axis=1)
    hit = ic[:, 0] != 0xFFFFFFFF
    assert hit.mean() > 0.2
    tie = ~same & (np.abs(tg[:, 0] - tc[:, 0]) <= 1e-5 * np.maximum(1.0, np.abs(tc[:, 0])))  # same t, other primitive
    assert (same | tie).mean() >= 0.999, (same.mean(), tie.mean())
    ok = same & hit
    assert (np.abs(tg[ok, 0] - tc[ok, 0]) / np.maximum(1.0, tc[ok, 0])).max() <= 1e-4
    assert np.quantile(np.abs(tg[ok, 1:] - tc[ok, 1:]).max(axis=1), 0.999) <= 1e-3  # barycentrics (grazing hits are ill-conditioned)
    # any-hit kernel agrees with "closest hit exists below tmax"
    rays[:, 7] = np.where(hit, np.maximum(tc[:, 0] * 1.5, 1e-3), 5.0)
    og, oc = gpu_ctx.occlusion_rays(rays), cpu_ctx.occlusion_rays(rays)
    assert (og == oc).mean() >= 0.999
    rays[:, 7] = tc[:, 0] * 0.5  # stop well short of the first hit: nothing may be reported
    assert gpu_ctx.occlusion_rays(rays)[hit].sum() <= 1e-4 * hit.sum()


def test_watertight_on_shared_edges_and_vertices(gpu_ctx):
    """Rays aimed exactly at mesh vertices and edge midpoints of a closed mesh from outside must hit it."""
    v, idx = scenes.blob(4, 77, 0.2)
    sc = host.Scene()
    sc.set_camera("perspective", 8, 8)
    sc.add_material("m", scenes.mat(0, diffuse=(.5, .5, .5)))
    sc.add_mesh("blob", v, idx)
    sc.add_instance("blob", "m")
    sc.shots.append(host.Shot((0, 0, 5), (0, 0, 0), (0, 1, 0)))
    sc.upload(gpu_ctx)
    from helpers import rays_toward, transversal_targets
    for origin in ((0.003, -0.002, 6.0), (5.0, 3.0, -2.0), (-4.0, -4.0, 4.0)):
        targets = transversal_targets(v["pos"], idx, origin)
        assert len(targets) > 3000
        _, ip = gpu_ctx.trace_rays(rays_toward(origin, targets))
        assert (ip[:, 0] != 0xFFFFFFFF).all(), f"{(ip[:, 0] == 0xFFFFFFFF).sum()} rays leaked through shared edges"


def test_bvh_build_stats(gpu_ctx):
    sc = scenes.glass_blob(16, 16, subdiv=6, env_size=(16, 8))
    ms = sc.upload(gpu_ctx)
    st = gpu_ctx.accel_stats()
    n_tris = sum(len(i) // 3 for _, i in sc.meshes)
    # every triangle sits in exactly one primitive slot; an 8-wide node with <= 3 triangles per leaf needs far
    # fewer nodes than the binary hierarchy it was collapsed from
    assert st["leaf_prims"] == n_tris and len(sc.meshes) <= st["nodes"] <= n_tris // 4 and st["tlas_nodes"] >= 1
    assert 0 < ms < 1000 and st["sah_cost"] > 0


@pytest.mark.parametrize("which", ["blob_single_mesh", "glass_flattened", "instanced_flattened"])
def test_bvh_is_a_valid_tree(gpu_ctx, which):
    """The builder's output walked on the host (asuna_debug_download_accel): every triangle slot is referenced by
    exactly one leaf child, every triangle lies inside the decoded quantised box of every ancestor slot (conservative at
    every level), every slot holds each (instance, primitive) pair once, and the SAH cost the builder reports matches the
    one recomputed from the decoded boxes."""
    from helpers import download_accel, walk_wide_bvh
    if which == "blob_single_mesh":
        sc = scenes.Scene()
        sc.set_camera("perspective", 16, 16)
        sc.add_material("m", scenes.mat(0))
        sc.add_mesh("blob", *scenes.blob(5, 7, 0.2))
        sc.add_instance("blob", "m")
        sc.shots.append(host.Shot((0, 0, 4), (0, 0, 0), (0, 1, 0)))
    elif which == "glass_flattened":
        sc = scenes.glass_blob(16, 16, subdiv=5, env_size=(16, 8))
    else:
        sc = scenes.instanced_field(16, 16, subdiv=4, grid=5)
    sc.upload(gpu_ctx)
    nodes, tris, root = download_accel(gpu_ctx)
    assert root != 0xFFFFFFFF, "scene was expected to be flattened into one world-space BVH"
    r = walk_wide_bvh(nodes, tris, root)
    n_inst_tris = sum(len(sc.meshes[m][1]) // 3 for _, m, _, _ in sc.instances)
    assert len(tris) == n_inst_tris and (r["refs"] == 1).all(), (len(tris), n_inst_tris, np.bincount(r["refs"]))
    assert r["outside"] == 0
    inst = tris["inst"] & 0x0FFFFFFF  # the top four bits of the word carry the instance's hit kind (miss 0, emitter 1, 2 + material type)
    kinds = tris["inst"] >> 28
    want_kind = np.array([1 if light >= 0 else 2 + int(sc.materials[mat]["type"]) for _, _, mat, light in sc.instances])
    assert np.array_equal(kinds, want_kind[inst])
    pairs = inst.astype(np.uint64) << np.uint64(32) | tris["prim"].astype(np.uint64)
    assert len(np.unique(pairs)) == len(pairs)
    per_inst = np.bincount(inst, minlength=len(sc.instances))
    assert [int(c) for c in per_inst] == [len(sc.meshes[m][1]) // 3 for _, m, _, _ in sc.instances]
    st = gpu_ctx.accel_stats()
    assert r["nodes_used"] == st["nodes"] and r["max_depth"] <= 24
    assert abs(r["sah"] / st["sah_cost"] - 1.0) <= 0.05, (r["sah"], st["sah_cost"])  # decoded boxes are a little larger


def soup_scene(n, mode, seed=3):
    """n small triangles: 'random' in the unit cube, 'line' with all centroids on one axis, 'same' all at one place
    (every Morton key equal), 'two_clusters' far apart (empty space inside the root)."""
    rng = np.random.RandomState(seed + n)
    c = rng.rand(n, 3)
    if mode == "line":
        c[:, 1:] = 0.5
    elif mode == "same":
        c[:] = 0.5
    elif mode == "two_clusters":
        c = c * 0.01 + np.where(rng.rand(n, 1) < 0.5, 0.0, 100.0)
    v = np.zeros(3 * n, S.Vertex)
    v["pos"] = (c[:, None, :] + 0.02 * (rng.rand(n, 3, 3) - 0.5)).reshape(-1, 3)
    v["normal"] = (0, 0, 1)
    sc = scenes.Scene()
    sc.set_camera("perspective", 16, 16)
    sc.add_material("m", scenes.mat(0))
    sc.add_mesh("soup", v, np.arange(3 * n, dtype=np.uint32))
    sc.add_instance("soup", "m")
    sc.shots.append(host.Shot((0.5, 0.5, 4), (0.5, 0.5, 0.5), (0, 1, 0)))
    return sc, v["pos"].reshape(n, 3, 3).mean(axis=1).astype(np.float64)  # the triangles' centroids


@pytest.mark.parametrize("n,mode", [(1, "random"), (2, "random"), (3, "random"), (4, "random"), (9, "random"), (33, "random"),
                                    (255, "same"), (1000, "line"), (2047, "random"), (2048, "random"), (2049, "random"),
                                    (2305, "two_clusters"), (4097, "random"), (70001, "random"), (70001, "same")])
def test_builder_at_its_thresholds(gpu_ctx, cpu_ctx, n, mode):
    """Primitive counts around every switch of the builder (one-primitive BVH, <= 3 primitives = a single leaf child,
    the 2048-cluster hand-over from the grid rounds to the shared-memory tail, one / several sort tiles) and degenerate
    placements (all Morton keys equal, collinear centroids, two far clusters): the tree is a valid tree over exactly the
    n primitives, and rays aimed at the triangles find the same nearest hit as the oracle's BVH2."""
    from helpers import download_accel, walk_wide_bvh
    sc, c = soup_scene(n, mode)
    sc.upload(gpu_ctx), sc.upload(cpu_ctx)
    nodes, tris, root = download_accel(gpu_ctx)
    r = walk_wide_bvh(nodes, tris, root)
    assert len(tris) == n and (r["refs"] == 1).all() and r["outside"] == 0
    assert sorted(tris["prim"].tolist()) == list(range(n))
    st = gpu_ctx.accel_stats()
    assert st["leaf_prims"] == n and r["nodes_used"] == st["nodes"]
    rng = np.random.RandomState(n)
    k = max(min(n, 4000), 64)
    tgt = c[rng.randint(0, n, k)] + 0.0004 * (rng.rand(k, 3) - 0.5)
    org = tgt + (rng.rand(k, 3) - 0.5) * (300.0 if mode == "two_clusters" else 3.0)
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, np.zeros((k, 1)), d, np.full((k, 1), 1e30)], axis=1).astype(np.float32)
    tg, ig = gpu_ctx.trace_rays(rays)
    tc, ic = cpu_ctx.trace_rays(rays)
    hit_g, hit_c = ig[:, 0] != 0xFFFFFFFF, ic[:, 0] != 0xFFFFFFFF
    assert np.array_equal(hit_g, hit_c) and hit_c.mean() > 0.3
    same = (ig == ic).all(axis=1)
    tie = ~same & (np.abs(tg[:, 0] - tc[:, 0]) <= 1e-5 * np.maximum(1.0, np.abs(tc[:, 0])))  # same t, other primitive
    assert (same | tie).mean() >= 0.999 and np.allclose(tg[same & hit_c, 0], tc[same & hit_c, 0], rtol=1e-5, atol=1e-6)


def test_bvh_sah_quality_against_cpu_binned_sah(gpu_ctx, cpu_ctx, oracle_lib):
    """SURVEY.md section 4 gate: SAH cost of the GPU tree (PLOC + optimal 8-wide collapse) <= 1.15 x the cost of the
    oracle's 16-bin SAH BVH2 collapsed by the same dynamic programme (oracle/bvh.h wide_sah_cost), same cost model
    (c_node 1, c_triangle 1, <= 3 triangles per leaf), on the 82 k-triangle blob of the benched scene."""
    import ctypes as C
    sc = scenes.Scene()
    sc.set_camera("perspective", 16, 16)
    sc.add_material("m", scenes.mat(0))
    sc.add_mesh("blob", *scenes.blob(6, 1234, 0.18))
    sc.add_instance("blob", "m")
    sc.shots.append(host.Shot((0, 0, 4), (0, 0, 0), (0, 1, 0)))
    sc.upload(gpu_ctx), sc.upload(cpu_ctx)
    f = oracle_lib.lib.oracle_wide_sah_cost
    f.restype = C.c_double
    cpu = f(cpu_ctx.h, C.c_uint32(0), C.c_double(1.0), C.c_double(1.0), C.c_uint32(3))
    gpu = gpu_ctx.accel_stats()["sah_cost"]
    print(f"wide-BVH SAH cost: GPU builder {gpu:.2f}, CPU binned SAH + collapse {cpu:.2f}, ratio {gpu / cpu:.3f}")
    assert cpu > 0 and gpu <= 1.15 * cpu, (gpu, cpu)
