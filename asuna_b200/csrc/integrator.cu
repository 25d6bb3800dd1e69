// Wavefront path-tracing integrator, hand-written for sm_100a.
//
// Replaces the Vulkan ray-tracing pipeline dispatch of the reference (one vkCmdTraceRaysKHR(w,h,1)
// per frame, reference src/pipeline/pipeline_raytrace.cpp:66-73, running
// src/shaders/raytrace.projective.rgen with the closest-hit / miss shaders of src/shaders/**).
// The implicit RT-pipeline scheduler (traceRayEXT -> SBT lookup -> shader) becomes explicit:
//
//   k_raygen            one thread per (frame-in-batch, pixel): RNG seed, jitter, camera ray
//   loop over bounces   k_trace_closest -> k_shade -> k_trace_shadow
//   k_accumulate        clamp, Gaussian filter weight, running weighted mean (rgen:134-178)
//
// Path state lives in float4 SoA arrays (16-byte accesses even when the compacted queue makes
// them a gather); queues are compacted with warp ballots + one atomic per warp; the two trace
// kernels are persistent with warp-granular dynamic work fetch.  No tensor cores: nothing here
// is a dense contraction.  B200 has no RT cores, so traversal is a software stack walk over the
// BVH built by bvh_build.cu with a watertight ray/triangle test (Woop, Benthin, Wald 2013).
#include <algorithm>
#include <utility>

#include "integrator.cuh"
#include "scan.cuh"
#include "shading.cuh"
#ifndef ASUNA_TRACE_STAGE
#define ASUNA_TRACE_STAGE 1  // prepared-ray staging in the single-level trace kernels (traverse.cuh)
#endif
#include "traverse.cuh"

namespace asuna {

namespace {

constexpr int kShadeThreads = 128;

// ---- trace kernels: persistent warps over the ray queues (traverse.cuh) -------------------------
struct ClosestPolicy {  // rgen:108-109: tmin 1e-5, tmax 1e10; the hit record goes back into the path slot
  PathState ps;
  const uint32_t* queue;
  const DInstance* instances;
  ADEV uint32_t load(uint32_t i, float3& o, float3& d, float& tmin, float& tmax) const {
    const uint32_t slot = queue[i];
    o = f3(ps.ray_o[slot]), d = f3(ps.ray_d[slot]);
    tmin = kMinimum, tmax = kInfinity;
    return slot;
  }
  ADEV void commit(uint32_t i, uint32_t slot, bool, const HitRec& h, uint32_t kind_hint) const {
    ps.hit[slot] = make_uint4(__float_as_uint(h.b1), __float_as_uint(h.b2), h.inst, h.prim);
    uint32_t kind = kKindMiss;  // the key the hit queue is regrouped by before shading
    if (h.inst != 0xFFFFFFFFu) {
      if (kind_hint != kKindUnknown) {
        kind = kind_hint;  // single-level kernels: came with the triangle slot
      } else {
        const uint32_t mt = __ldg(&instances[h.inst].mat_type);
        kind = mt == 0xFFFFFFFFu ? (uint32_t)kKindLight : kKindMaterial0 + mt;
      }
    }
    ps.kind[i] = (uint8_t)kind;
  }
};

struct ShadowPolicy {  // rgen:117-125: tmin 0, tmax = dist - 2 EPS, first hit ends it; unoccluded adds dRec.radiance
  PathState ps;
  ADEV uint32_t load(uint32_t i, float3& o, float3& d, float& tmin, float& tmax) const {
    const float4 a = ps.sh_o[i];
    o = f3(a), d = f3(ps.sh_d[i]);
    tmin = 0.0f, tmax = a.w;
    return 0u;
  }
  ADEV void commit(uint32_t i, uint32_t, bool occluded, const HitRec&, uint32_t) const {
    if (occluded) return;
    const uint32_t slot = __float_as_uint(ps.sh_d[i].w);
    const float4 L = ps.sh_l[i], r = ps.rad[slot];
    ps.rad[slot] = make_float4(r.x + L.x, r.y + L.y, r.z + L.z, r.w);
  }
};

struct UserPolicy {  // asuna_trace_rays / asuna_occlusion_rays / asuna_trace_primary (parity tests)
  const float4* rays;
  float* tuv;
  uint32_t* inst_prim;
  uint8_t* occluded;
  ADEV uint32_t load(uint32_t i, float3& o, float3& d, float& tmin, float& tmax) const {
    const float4 a = rays[2 * i], b = rays[2 * i + 1];
    o = f3(a), d = f3(b), tmin = a.w, tmax = b.w;
    return 0u;
  }
  ADEV void commit(uint32_t i, uint32_t, bool found, const HitRec& h, uint32_t) const {
    if (occluded) {
      occluded[i] = found ? 1 : 0;
      return;
    }
    if (tuv) tuv[3 * i] = found ? h.t : 0.f, tuv[3 * i + 1] = h.b1, tuv[3 * i + 2] = h.b2;
    inst_prim[2 * i] = h.inst, inst_prim[2 * i + 1] = h.prim;
  }
};

template <bool COUNT, bool SINGLE>
__global__ void __launch_bounds__(kTraceThreads, SINGLE ? ASUNA_TRACE_MIN_BLOCKS_SINGLE : ASUNA_TRACE_MIN_BLOCKS)
k_trace_closest(const __grid_constant__ SceneView sc, PathState ps, Counters* cnt, int iter, int qsel) {
  ClosestPolicy pol{ps, ps.queue[qsel], sc.instances};
  constexpr bool kStage = SINGLE && ASUNA_TRACE_STAGE;
  __shared__ __align__(16) uint32_t stage[kStage ? (kTraceThreads / 32) * kStageWords * 32 : 4];
#if ASUNA_TRI_POOL
  if constexpr (SINGLE) {
    __shared__ __align__(16) uint32_t pool[(kTraceThreads / 32) * kPoolWordsPerWarp];
    trace_persistent_pool<COUNT>(sc, pol, cnt->queue[iter], &cnt->ticket_closest[iter], &cnt->stack_overflow, &cnt->node_visits,
                                 &cnt->tri_tests, stage, pool);
    return;
  }
#endif
  trace_persistent<false, COUNT, SINGLE, kStage>(sc, pol, cnt->queue[iter], &cnt->ticket_closest[iter], &cnt->stack_overflow,
                                                 &cnt->node_visits, &cnt->tri_tests, cnt->lane_stats, stage);
}

template <bool SINGLE>
__global__ void __launch_bounds__(kTraceThreads, SINGLE ? ASUNA_SHADOW_MIN_BLOCKS_SINGLE : ASUNA_TRACE_MIN_BLOCKS)
k_trace_shadow(const __grid_constant__ SceneView sc, PathState ps, Counters* cnt, int iter) {
  ShadowPolicy pol{ps};  // (prepared-ray staging measured -3 % on the short any-hit traversals: not used here)
  trace_persistent<true, false, SINGLE, false>(sc, pol, cnt->shadow[iter], &cnt->ticket_shadow[iter], &cnt->stack_overflow,
                                               nullptr, nullptr);
}

template <bool ANY, bool SINGLE>
__global__ void __launch_bounds__(kTraceThreads)
k_trace_user(const __grid_constant__ SceneView sc, UserPolicy pol, uint32_t n, uint32_t* ticket, Counters* cnt) {
  trace_persistent<ANY, false, SINGLE, false>(sc, pol, n, ticket, &cnt->stack_overflow, nullptr, nullptr);
}

// Sun & sky: everything that depends on the setting only, once per setting instead of once per lookup.
__global__ void k_sky_prepare(AsunaSunSky ss, SkyPre* out) {
  SkyPre p;
  sky_prepare(ss, p);
  *out = p;
}

// Adds one batch's per-iteration counters into the persistent totals (one thread; a few hundred words).
__global__ void k_fold_counters(const Counters* cnt, Totals* tot, int iters) {
  unsigned long long c = 0, s = 0, inc = 0;
  for (int i = 0; i < iters; i++) c += cnt->queue[i], s += cnt->shadow[i], inc += cnt->incoherent[i];
  tot->closest_rays += c;
  tot->shadow_rays += s;
  tot->incoherent_rays += inc;
  tot->node_visits += cnt->node_visits;
  tot->tri_tests += cnt->tri_tests;
  tot->stack_overflow += cnt->stack_overflow;
  for (int k = 0; k < 6; k++) tot->lane_stats[k] += cnt->lane_stats[k];
}

// ---- camera: raytrace.projective.rgen:41-87 ---------------------------------------------------
ADEV void camera_ray(const AsunaCamera& cam, uint32_t x, uint32_t y, float2 jitter, uint32_t& seed, float3& o,
                     float3& d) {
  float3 origin = mat4_point(cam.cameraToWorld, f3(0.0f));
  float px = (float)x + jitter.x, py = (float)y + jitter.y;
  o = origin;
  d = f3(0.0f);
  if (cam.type == ASUNA_CAMERA_PERSPECTIVE) {
    float3 r = make_normal(mat4_point(cam.rasterToCamera, f3(px, py, 0.f)));
    if (cam.aperture > 0.f) {
      float2 lens = concentric_sample_disk(rnd2(seed));
      lens.x *= cam.aperture;
      lens.y *= cam.aperture;
      float ft = cam.focalDistance / r.z;
      float3 focus = ft * r;
      float3 lo = f3(lens.x, lens.y, 0.f);
      o = mat4_point(cam.cameraToWorld, lo);
      r = focus - lo;
    }
    d = make_normal(mat4_vector(cam.cameraToWorld, r));
  } else if (cam.type == ASUNA_CAMERA_OPENCV) {
    float3 r = f3((px - cam.fxfycxcy[2]) / cam.fxfycxcy[0], (py - cam.fxfycxcy[3]) / cam.fxfycxcy[1], 1.f);
    d = make_normal(mat4_vector(cam.cameraToWorld, r));
  }
}

ADEV float pack_depth_flags(int depth, uint32_t flags) { return __uint_as_float((uint32_t)depth | (flags << 16)); }

__global__ void __launch_bounds__(256)
k_raygen(const __grid_constant__ FrameParams fp, PathState ps, OutputImages out, Counters* cnt) {
  uint32_t total = fp.n_pixels * fp.n_frames;
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) cnt->queue[0] = total;
  if (s >= total) return;
  uint32_t fi = s / fp.n_pixels, pixel = s - fi * fp.n_pixels;
  uint32_t x = pixel % fp.width, y = pixel / fp.width;
  int frame = fp.frame_ids[fi];
  uint32_t seed = xxhash32_seed(x, y, (uint32_t)frame);
  float2 jitter = frame == 0 ? make_float2(0.5f, 0.5f) : rnd2(seed);
  float3 o, d;
  camera_ray(fp.cam, x, y, jitter, seed, o, d);
  ps.ray_o[s] = make_float4(o.x, o.y, o.z, __uint_as_float(seed));
  ps.ray_d[s] = make_float4(d.x, d.y, d.z, 0.0f);
  ps.thr[s] = make_float4(1.f, 1.f, 1.f, pack_depth_flags(1, kBsdfNull));
  ps.rad[s] = make_float4(0.f, 0.f, 0.f, 0.f);
  ps.queue[0][s] = s;
  if (frame == 0)  // rgen:159-161 writes every declared channel; misses leave zeros
    for (uint32_t c = 0; c < fp.pc.nMultiChannel && c < ASUNA_NUM_OUTPUT_IMAGES - 1; c++)
      out.img[c + 1][pixel] = make_float4(0.f, 0.f, 0.f, 1.f);
}

// rchit_layouts.glsl:67-95
ADEV void load_surface(const SceneView& sc, const AsunaState& pc, const DInstance& in, uint4 hit, float3 ray_d,
                       Surface& s) {
  const DMesh mesh = sc.meshes[in.mesh];
  const uint32_t* id = mesh.indices + 3 * (size_t)hit.w;
  const AsunaVertex* v0 = mesh.vertices + __ldg(id + 0);
  const AsunaVertex* v1 = mesh.vertices + __ldg(id + 1);
  const AsunaVertex* v2 = mesh.vertices + __ldg(id + 2);
  float b1 = __uint_as_float(hit.x), b2 = __uint_as_float(hit.y), b0 = 1.0f - b1 - b2;
  float3 p0 = f3(v0->pos), p1 = f3(v1->pos), p2 = f3(v2->pos);
  s.uv = make_float2(v0->uv[0] * b0 + v1->uv[0] * b1 + v2->uv[0] * b2, v0->uv[1] * b0 + v1->uv[1] * b1 + v2->uv[1] * b2);
  s.pos = xf_point(in.o2w, p0 * b0 + p1 * b1 + p2 * b2);
  float3 n = f3(v0->normal) * b0 + f3(v1->normal) * b1 + f3(v2->normal) * b2;
  s.N = make_normal(xf_normal(in.w2o, n));
  s.geoN = make_normal(xf_normal(in.w2o, cross(p1 - p0, p2 - p0)));
  s.ffN = s.geoN;
  s.V = make_normal(-ray_d);
  configure_frame(pc, s);
}

// ---- regroup the hit queue by kind: one counting-sort pass (histogram / scan / stable scatter) ----
// The RT pipeline of the reference picks the closest-hit shader per instance through the shader binding
// table; the wavefront equivalent is a queue per shader.  Keys are the bytes the trace kernel left in
// ps.kind, values the path slots; kNumKinds bins.
// Both kernels fetch everything a thread will look at in one batch before the ranking loops: a loop that loads a key,
// ranks it and goes on waits one DRAM latency per round, 32 rounds per block (the same finding as the builder's radix
// scatter: 30 + 60 us per bounce for 10 M paths that move in 15).
__global__ void __launch_bounds__(256) k_bin_count(PathState ps, const Counters* cnt, int iter, uint32_t n_blocks) {
  constexpr int kRounds = kBinTile / 256;
  __shared__ uint32_t h[kNumKinds];
  if (threadIdx.x < kNumKinds) h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t count = cnt->queue[iter], base = blockIdx.x * kBinTile;
  if (base < count) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t k[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
      const uint32_t i = base + r * 256 + threadIdx.x;
      k[r] = i < count ? ps.kind[i] : kNumKinds + lane;
    }
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, k[r]);
      if (k[r] < kNumKinds && lane == (uint32_t)__ffs((int)peers) - 1u) atomicAdd(&h[k[r]], (uint32_t)__popc(peers));
    }
  }
  __syncthreads();
  if (threadIdx.x < kNumKinds) ps.bin_hist[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(256) k_bin_scatter(PathState ps, const Counters* cnt, int iter, int qsel, uint32_t n_blocks) {
  constexpr int kWarps = 8, kRounds = kBinTile / kWarps / 32;
  __shared__ uint32_t wh[kWarps][kNumKinds];
  const uint32_t count = cnt->queue[iter], base = blockIdx.x * kBinTile;
  if (base >= count) return;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  if (threadIdx.x < kWarps * kNumKinds) (&wh[0][0])[threadIdx.x] = 0;
  const uint32_t* queue = ps.queue[qsel];
  const uint32_t wbase = base + warp * (kRounds * 32);
  uint32_t kk[kRounds / 4], q[kRounds];  // four kind bytes per register
#pragma unroll
  for (int r = 0; r < kRounds / 4; r++) kk[r] = 0;
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const uint32_t i = wbase + r * 32 + lane;
    const uint32_t kr = i < count ? ps.kind[i] : kNumKinds + lane;
    kk[r >> 2] |= kr << (8 * (r & 3));
    q[r] = i < count ? queue[i] : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const uint32_t kr = (kk[r >> 2] >> (8 * (r & 3))) & 0xFFu;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, kr);
    if (kr < kNumKinds && lane == (uint32_t)__ffs((int)peers) - 1u) wh[warp][kr] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < kNumKinds) {  // global offset of this block per kind, then exclusive over its warps
    uint32_t run = ps.bin_hist[threadIdx.x * n_blocks + blockIdx.x];
    for (int w = 0; w < kWarps; w++) {
      uint32_t c = wh[w][threadIdx.x];
      wh[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const uint32_t kr = (kk[r >> 2] >> (8 * (r & 3))) & 0xFFu;
    const bool valid = kr < kNumKinds;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, kr);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid) ps.sorted[wh[warp][kr] + rank] = q[r];
    __syncwarp();
    if (valid && lane == (uint32_t)__ffs((int)peers) - 1u) wh[warp][kr] += __popc(peers);
    __syncwarp();
  }
}

// ---- shading: one kernel per hit kind (≙ one closest-hit / miss shader each) -----------------------
// Resident blocks per SM the shade kernels are compiled for.  They are latency-bound (dependent gathers: instance ->
// indices -> vertices -> material -> texels), so occupancy competes with register-resident state, and what a capped
// kernel spills is not free: local-memory lines of finished threads are dirty and travel to L2 / DRAM (ncu: 20 GB of
// local sectors in one 16 M-path launch of the PBR shader at 40 registers, more than its path state and texels
// together).  Round 1 (per-thread copies of the instance / material records and an AOV pointer array on the stack):
// 8 blocks, 12 for the shaders that fetch up to five textures.  Round 2, with those stack objects gone: 6 blocks (80
// registers) and 8 (64) -- measured 4 / 5 / 6 / 7 / 8 / 10 / 12 and 4 / 5 / 6 / 8 / 10 / 12 / 16 on C1 / C2 / C3 / C4.
#ifndef ASUNA_SHADE_MIN_BLOCKS
#define ASUNA_SHADE_MIN_BLOCKS 6
#endif
#ifndef ASUNA_SHADE_MIN_BLOCKS_TEXTURED
#define ASUNA_SHADE_MIN_BLOCKS_TEXTURED 8
#endif
constexpr bool shade_kind_textured(uint32_t kind) {
  return kind == kKindMaterial0 + ASUNA_MAT_PBR_METALNESS_ROUGHNESS || kind == kKindMaterial0 + ASUNA_MAT_KANG18 ||
         kind == kKindMaterial0 + ASUNA_MAT_DISNEY;
}
template <uint32_t KIND>
__global__ void __launch_bounds__(kShadeThreads, shade_kind_textured(KIND) ? ASUNA_SHADE_MIN_BLOCKS_TEXTURED : ASUNA_SHADE_MIN_BLOCKS)
k_shade(const __grid_constant__ SceneView sc, const __grid_constant__ FrameParams fp, const __grid_constant__ PathState ps,
        const __grid_constant__ OutputImages out, Counters* cnt, int iter,
        int qsel, uint32_t n_blocks) {
  const uint32_t begin = ps.bin_hist[KIND * n_blocks];
  const uint32_t end = KIND + 1 < kNumKinds ? ps.bin_hist[(KIND + 1) * n_blocks] : cnt->queue[iter];
  const uint32_t* queue = ps.sorted;
  uint32_t* next_queue = ps.queue[qsel ^ 1];
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;

  ShadeEnv se;
  se.scene = &sc;
  se.fp = &fp;
  se.env.env = sc.env;
  se.env.env_transform = fp.cam.envTransform;
  se.env.res_x = fp.pc.envMapResolution[0];
  se.env.res_y = fp.pc.envMapResolution[1];
  se.env.intensity = fp.pc.envMapIntensity;
  se.out = &out;

  // The loop is a chain of dependent gathers per path (queue -> slot -> state + hit -> instance -> indices -> vertices ->
  // material), ~70 paths per thread one after the other.  The first hop of the NEXT path (its slot index) is fetched
  // before the current one is shaded.  Measured: fetching the next path's state and hit record ahead as well costs 20
  // live registers through the whole shader and is 2.7 % (C2) to 5 % (C1) slower than no prefetch at all; a
  // `prefetch.global.L2` of those lines instead (no registers) is 1-3 % slower too.
  constexpr bool kRadiance = !(KIND == kKindMaterial0 + ASUNA_MAT_DIELECTRIC || KIND == kKindMaterial0 + ASUNA_MAT_CONDUCTOR ||
                               KIND == kKindMaterial0 + ASUNA_MAT_MIRROR);  // the delta shaders neither read nor change the
                                                                              // path radiance: 32 B less per hit
  const uint32_t stride = n_warps * 32, first = begin + warp_global * 32 + lane;
  uint32_t slot_ahead = first < end ? queue[first] : 0u;
  for (uint32_t base = begin + warp_global * 32; base < end; base += stride) {
    uint32_t i = base + lane;
    bool valid = i < end;
    bool cont = false, nee = false, incoherent = false;
    uint32_t slot = slot_ahead;
    slot_ahead = i + stride < end ? queue[i + stride] : 0u;
    PathRegs p;
    p.nee = false;
    if (valid) {
      const float4 ro = ps.ray_o[slot], rd = ps.ray_d[slot], th = ps.thr[slot];
      float4 ra = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (kRadiance) ra = ps.rad[slot];
      uint32_t fi = slot / fp.n_pixels, pixel = slot - fi * fp.n_pixels;
      bool frame0 = fp.frame_ids[fi] == 0;
      se.frame0 = frame0;
      p.ray_o = f3(ro), p.ray_d = f3(rd), p.throughput = f3(th), p.radiance = f3(ra);
      p.seed = __float_as_uint(ro.w);
      p.bsdf_pdf = rd.w;
      uint32_t packed = __float_as_uint(th.w);
      p.depth = (int)(packed & 0xFFFFu);
      p.bsdf_flags = packed >> 16;
      p.stop = false;
      p.nee_L = f3(0.0f);
      if constexpr (KIND == kKindMiss) {
        shade_miss(se, p);
      } else {
        const uint4 hit = ps.hit[slot];
        const DInstance& in = sc.instances[hit.z];  // fields are fetched where they are used (L1-resident tables): a by-value
                                                    // copy of the 128-byte record lives in local memory
        Surface s;
        load_surface(sc, fp.pc, in, hit, p.ray_d, s);
        if constexpr (KIND == kKindLight) {
          shade_light_hit(se, p, in.light, s.pos);
        } else {
          const AsunaMaterial& m = sc.materials[in.material];
          constexpr uint32_t T = KIND - kKindMaterial0;
          if constexpr (T == ASUNA_MAT_LAMBERTIAN) shade_lambertian(se, p, s, m, pixel);
          else if constexpr (T == ASUNA_MAT_EMISSIVE) shade_emissive(se, p, s, m);
          else if constexpr (T == ASUNA_MAT_DIELECTRIC) shade_dielectric(se, p, s, m);
          else if constexpr (T == ASUNA_MAT_CONDUCTOR) shade_conductor(se, p, s, m);
          else if constexpr (T == ASUNA_MAT_PLASTIC) shade_plastic(se, p, s, m);
          else if constexpr (T == ASUNA_MAT_ROUGH_PLASTIC) shade_rough_plastic(se, p, s, m);
          else if constexpr (T == ASUNA_MAT_PBR_METALNESS_ROUGHNESS) shade_pbr(se, p, s, m, pixel);
          else if constexpr (T == ASUNA_MAT_KANG18) shade_kang18(se, p, s, m, in, pixel);
          else if constexpr (T == ASUNA_MAT_MIRROR) shade_mirror(se, p, s, m);
          else if constexpr (T == ASUNA_MAT_ROUGH_CONDUCTOR) shade_rough_conductor(se, p, s, m);
          else if constexpr (T == ASUNA_MAT_PHONG) shade_phong(se, p, s, m, pixel);
          else if constexpr (T == ASUNA_MAT_DISNEY) shade_disney(se, p, s, m);
          else p.stop = true;
        }
      }
      // rgen:112-131: shadow ray if the hit shader asked for one, then stop / depth++
      nee = p.nee && (p.nee_L.x != 0.0f || p.nee_L.y != 0.0f || p.nee_L.z != 0.0f);  // zero NEE adds nothing (A.3-4)
      int next_depth = p.depth + 1;
      cont = !p.stop && next_depth <= fp.pc.maxPathDepth;
      incoherent = cont && next_depth >= 2;
      if constexpr (kRadiance)
        if (p.radiance.x != ra.x || p.radiance.y != ra.y || p.radiance.z != ra.z)
          ps.rad[slot] = make_float4(p.radiance.x, p.radiance.y, p.radiance.z, ra.w);
      if (cont) {
        ps.ray_o[slot] = make_float4(p.ray_o.x, p.ray_o.y, p.ray_o.z, __uint_as_float(p.seed));
        ps.ray_d[slot] = make_float4(p.ray_d.x, p.ray_d.y, p.ray_d.z, p.bsdf_pdf);
        ps.thr[slot] = make_float4(p.throughput.x, p.throughput.y, p.throughput.z,
                                   pack_depth_flags(next_depth, p.bsdf_flags));
      }
    }
    // queue compaction: one atomic per warp per queue
    uint32_t lt = (1u << lane) - 1u;
    uint32_t m_cont = __ballot_sync(0xFFFFFFFFu, cont);
    uint32_t m_nee = __ballot_sync(0xFFFFFFFFu, nee);
    uint32_t m_inc = __ballot_sync(0xFFFFFFFFu, incoherent);
    uint32_t b_cont = 0, b_nee = 0;
    if (lane == 0) {
      if (m_cont) b_cont = atomicAdd(&cnt->queue[iter + 1], (uint32_t)__popc(m_cont));
      if (m_nee) b_nee = atomicAdd(&cnt->shadow[iter], (uint32_t)__popc(m_nee));
      if (m_inc) atomicAdd(&cnt->incoherent[iter + 1], (uint32_t)__popc(m_inc));
    }
    b_cont = __shfl_sync(0xFFFFFFFFu, b_cont, 0);
    b_nee = __shfl_sync(0xFFFFFFFFu, b_nee, 0);
    if (cont) next_queue[b_cont + __popc(m_cont & lt)] = slot;
    if (nee) {
      uint32_t k = b_nee + __popc(m_nee & lt);
      ps.sh_o[k] = make_float4(p.nee_o.x, p.nee_o.y, p.nee_o.z, p.nee_dist - 2 * kEps);
      ps.sh_d[k] = make_float4(p.nee_d.x, p.nee_d.y, p.nee_d.z, __uint_as_float(slot));
      ps.sh_l[k] = make_float4(p.nee_L.x, p.nee_L.y, p.nee_L.z, 0.f);
    }
  }
}

// rgen:134-178 for every frame of the batch, in frame order, one thread per pixel.
__global__ void __launch_bounds__(256) k_accumulate(const __grid_constant__ FrameParams fp, PathState ps, OutputImages out) {
  uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
  if (pixel >= fp.n_pixels) return;
  uint32_t x = pixel % fp.width, y = pixel / fp.width;
  float4 mean = out.img[0][pixel];
  float wsum = out.img[8][pixel].x;
  const float stddev = 0.5f, radius = 4 * stddev, alpha = -1.0f / (2.0f * stddev * stddev);
  const float exp_xy = expf(alpha * radius * radius);
  for (uint32_t fi = 0; fi < fp.n_frames; fi++) {
    int frame = fp.frame_ids[fi];
    float2 jitter = make_float2(0.5f, 0.5f);
    if (frame != 0) {
      uint32_t seed = xxhash32_seed(x, y, (uint32_t)frame);
      jitter = rnd2(seed);
    }
    float ox = jitter.x - 0.5f, oy = jitter.y - 0.5f;
    float w = fmaxf(0.0f, expf(alpha * ox * ox) - exp_xy) * fmaxf(0.0f, expf(alpha * oy * oy) - exp_xy);
    float4 r = ps.rad[fi * fp.n_pixels + pixel];
    float3 L = clamp3(f3(r), 0.0f, 10.0f);
    float3 wl = w * L;
    if (fi == 0 && fp.first_is_replace) {
      float3 m = wl / w;
      mean = make_float4(m.x, m.y, m.z, 1.f);
      wsum = w;
    } else {
      float3 old_sum = f3(mean) * wsum;
      float new_w = wsum + w;
      float3 m = (old_sum + wl) / new_w;
      mean = make_float4(m.x, m.y, m.z, 1.f);
      wsum = new_w;
    }
  }
  out.img[0][pixel] = mean;
  out.img[8][pixel] = make_float4(wsum, wsum, wsum, wsum);
}

__global__ void k_export_partial(OutputImages out, float4* partial, uint32_t n, int have_accum) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!have_accum) {
    partial[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float4 m = out.img[0][i];
  float w = out.img[8][i].x;
  partial[i] = make_float4(m.x * w, m.y * w, m.z * w, w);
}
__global__ void k_import_partial(OutputImages out, const float4* partial, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 s = partial[i];
  out.img[0][i] = make_float4(s.x / s.w, s.y / s.w, s.z / s.w, 1.f);
  out.img[8][i] = make_float4(s.w, s.w, s.w, s.w);
}

// ---- post-process: src/shaders/post.idle.frag:71-133 + utils/tonemapping.glsl, one thread per pixel -------------------
namespace post {
ADEV float3 pow3(float3 c, float e) { return f3(powf(c.x, e), powf(c.y, e), powf(c.z, e)); }
ADEV float3 linear_to_srgb(float3 c) { return pow3(c, 1.0f / 2.2f); }  // tonemapping.glsl:27-33 (INV_GAMMA = 1 / 2.2)
ADEV float3 srgb_to_linear(float3 c) { return pow3(c, 2.2f); }
ADEV float3 uncharted2(float3 c) {  // :43-54
  const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
  return ((c * (A * c + C * B) + D * E) / (c * (A * c + B) + D * F)) - E / F;
}
ADEV float3 tone_map_uncharted(float3 c) {  // :56-61
  c = uncharted2(c * 2.0f);
  const float3 white = f3(1.0f) / uncharted2(f3(11.2f));
  return linear_to_srgb(c * white);
}
ADEV float3 hejl_richard(float3 c) {  // :65-68
  c = f3(fmaxf(0.0f, c.x - 0.004f), fmaxf(0.0f, c.y - 0.004f), fmaxf(0.0f, c.z - 0.004f));
  return (c * (6.2f * c + 0.5f)) / (c * (6.2f * c + 1.7f) + 0.06f);
}
ADEV float3 aces(float3 c) {  // :73-81
  const float A = 2.51f, B = 0.03f, C = 2.43f, D = 0.59f, E = 0.14f;
  float3 v = (c * (A * c + B)) / (c * (C * c + D) + E);
  return linear_to_srgb(f3(fminf(fmaxf(v.x, 0.0f), 1.0f), fminf(fmaxf(v.y, 0.0f), 1.0f), fminf(fmaxf(v.z, 0.0f), 1.0f)));
}
ADEV float pbrt_channel(float x) { return x < 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f; }
ADEV float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
}  // namespace post

// mean of the radiance image (the 1x1 mip level textureLod(inImage, 0.5, 20) reads, post.idle.frag:104-105): fixed-shape
// two-level sum, so the result does not depend on scheduling
__global__ void __launch_bounds__(256) k_image_sum(const float4* __restrict__ img, uint32_t n, double* __restrict__ block_sums) {
  __shared__ double sh[3][8];
  double s[3] = {0.0, 0.0, 0.0};
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    float4 v = img[i];
    s[0] += v.x, s[1] += v.y, s[2] += v.z;
  }
  for (int c = 0; c < 3; c++) {
    for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xFFFFFFFFu, s[c], o);
    if ((threadIdx.x & 31) == 0) sh[c][threadIdx.x >> 5] = s[c];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += sh[threadIdx.x][w];
    block_sums[blockIdx.x * 3 + threadIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) k_post_process(const float4* __restrict__ hdr_img, float4* __restrict__ ldr, uint32_t w,
                                                      uint32_t h, AsunaPost tm, const double* __restrict__ block_sums,
                                                      uint32_t n_sum_blocks) {
  using namespace post;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const uint32_t x = i % w, y = i / w;
  const float4 in = hdr_img[i];
  float3 hdr = f3(in.x, in.y, in.z), out;
  switch (tm.tmType) {
    case ASUNA_TM_NONE: out = hdr; break;
    case ASUNA_TM_GAMMA: out = linear_to_srgb(hdr / (f3(1.0f) + hdr / 1.5f)); break;
    case ASUNA_TM_REINHARD: out = hejl_richard(hdr); break;  // post.idle.frag:80-81 maps "Reinhard" to Hejl-Richard
    case ASUNA_TM_ACES: out = aces(hdr); break;
    case ASUNA_TM_FILMIC: out = hejl_richard(hdr); break;     // :85-87, the same curve spelled out
    case ASUNA_TM_PBRT: out = f3(pbrt_channel(hdr.x), pbrt_channel(hdr.y), pbrt_channel(hdr.z)); break;
    default: {  // ASUNA_TM_CUSTOM, :101-132
      float3 c = hdr;
      if (tm.autoExposure & 1) {  // toneExposure, :39-44
        double sr = 0.0, sg = 0.0, sb = 0.0;
        for (uint32_t b = 0; b < n_sum_blocks; b++) sr += block_sums[3 * b], sg += block_sums[3 * b + 1], sb += block_sums[3 * b + 2];
        const double inv = 1.0 / ((double)w * (double)h);
        const float avg_lum = 0.2126f * (float)(sr * inv) + 0.7152f * (float)(sg * inv) + 0.0722f * (float)(sb * inv);
        const float Yxyz = 0.3575761f * c.x + 0.7151522f * c.y + 0.1191920f * c.z;  // row y of RGB2XYZ as GLSL builds it (column-major constructor)
        const float Y = (tm.key / avg_lum) * Yxyz;
        const float Yd = (Y * (1.0f + Y / (tm.Ywhite * tm.Ywhite))) / (1.0f + Y);
        c = c / Yxyz * Yd;
      }
      c = tone_map_uncharted(c * tm.avgLum);  // toneMap(), TONEMAP_UNCHARTED is defined at post.idle.frag:17
      // dithering, :22-27 with pcg3d noise (utils/random.glsl:74-84)
      uint32_t vx = x * 1664525u + 1013904223u, vy = y * 1664525u + 1013904223u, vz = 1013904223u;
      vx += vy * vz, vy += vz * vx, vz += vx * vy;
      vx ^= vx >> 16, vy ^= vy >> 16, vz ^= vz >> 16;
      vx += vy * vz, vy += vz * vx, vz += vx * vy;
      const float3 noise = f3(__uint_as_float(0x3f800000u | (vx >> 9)) - 1.0f, __uint_as_float(0x3f800000u | (vy >> 9)) - 1.0f,
                              __uint_as_float(0x3f800000u | (vz >> 9)) - 1.0f);
      const float quant = 1.0f / 255.0f;
      const float3 lin = srgb_to_linear(c), s = linear_to_srgb(lin);
      const float3 c0 = f3(floorf(s.x / quant) * quant, floorf(s.y / quant) * quant, floorf(s.z / quant) * quant);
      const float3 c1 = c0 + f3(quant);
      const float3 l0 = srgb_to_linear(c0), l1 = srgb_to_linear(c1);
      const float3 discr = f3(mixf(l0.x, l1.x, noise.x), mixf(l0.y, l1.y, noise.y), mixf(l0.z, l1.z, noise.z));
      c = f3(discr.x < lin.x ? c1.x : c0.x, discr.y < lin.y ? c1.y : c0.y, discr.z < lin.z ? c1.z : c0.z);
      // contrast, brightness, saturation, vignette
      c = f3(fminf(fmaxf(mixf(0.5f, c.x, tm.contrast), 0.0f), 1.0f), fminf(fmaxf(mixf(0.5f, c.y, tm.contrast), 0.0f), 1.0f),
             fminf(fmaxf(mixf(0.5f, c.z, tm.contrast), 0.0f), 1.0f));
      c = pow3(c, 1.0f / tm.brightness);
      const float g = 0.299f * c.x + 0.587f * c.y + 0.114f * c.z;
      c = f3(mixf(g, c.x, tm.saturation), mixf(g, c.y, tm.saturation), mixf(g, c.z, tm.saturation));
      const float ux = ((((float)x + 0.5f) / (float)w) * tm.renderingRatio[0] - 0.5f) * 2.0f;
      const float uy = ((((float)y + 0.5f) / (float)h) * tm.renderingRatio[1] - 0.5f) * 2.0f;
      c = c * (1.0f - (ux * ux + uy * uy) * tm.vignette);
      out = c;
    }
  }
  ldr[i] = make_float4(out.x, out.y, out.z, in.w);
}

__global__ void k_primary_rays(const __grid_constant__ FrameParams fp, float4* rays) {
  uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
  if (pixel >= fp.n_pixels) return;
  uint32_t x = pixel % fp.width, y = pixel / fp.width;
  uint32_t seed = xxhash32_seed(x, y, 0u);
  float3 o, d;
  camera_ray(fp.cam, x, y, make_float2(0.5f, 0.5f), seed, o, d);
  rays[2 * pixel] = make_float4(o.x, o.y, o.z, kMinimum);
  rays[2 * pixel + 1] = make_float4(d.x, d.y, d.z, kInfinity);
}

inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

}  // namespace

// ---- launchers ---------------------------------------------------------------------------------
void launch_raygen(cudaStream_t s, const FrameParams& fp, const PathState& ps, const OutputImages& out, Counters* cnt) {
  uint32_t total = fp.n_pixels * fp.n_frames;
  k_raygen<<<div_up(total, 256), 256, 0, s>>>(fp, ps, out, cnt);
}
void launch_trace_closest(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const PathState& ps, Counters* cnt,
                          int iter, int qsel, bool counting) {
  const bool single = sc.single_root != 0xFFFFFFFFu;  // every instance merged into the world BLAS: no instance level
  if (counting) {
    if (single) k_trace_closest<true, true><<<ld.trace_blocks_single, kTraceThreads, 0, s>>>(sc, ps, cnt, iter, qsel);
    else k_trace_closest<true, false><<<ld.trace_blocks, kTraceThreads, 0, s>>>(sc, ps, cnt, iter, qsel);
  } else {
    if (single) k_trace_closest<false, true><<<ld.trace_blocks_single, kTraceThreads, 0, s>>>(sc, ps, cnt, iter, qsel);
    else k_trace_closest<false, false><<<ld.trace_blocks, kTraceThreads, 0, s>>>(sc, ps, cnt, iter, qsel);
  }
}
void launch_sky_prepare(cudaStream_t s, const AsunaSunSky& ss, SkyPre* out) {
  k_sky_prepare<<<1, 1, 0, s>>>(ss, out);
}
void launch_fold_counters(cudaStream_t s, const Counters* cnt, Totals* tot, int iters) {
  k_fold_counters<<<1, 1, 0, s>>>(cnt, tot, iters);
}
template <uint32_t KIND>
static void launch_shade_kind(cudaStream_t s, uint32_t blocks, const SceneView& sc, const FrameParams& fp, const PathState& ps,
                              const OutputImages& out, Counters* cnt, int iter, int qsel, uint32_t nb) {
  k_shade<KIND><<<blocks, kShadeThreads, 0, s>>>(sc, fp, ps, out, cnt, iter, qsel, nb);
}
using ShadeLauncher = void (*)(cudaStream_t, uint32_t, const SceneView&, const FrameParams&, const PathState&,
                               const OutputImages&, Counters*, int, int, uint32_t);
template <uint32_t... K>
static void fill_launchers(ShadeLauncher* t, std::integer_sequence<uint32_t, K...>) {
  ((t[K] = launch_shade_kind<K>), ...);
}

// Regroups the hit queue of this bounce by kind, then runs one shade kernel per kind present in the scene.
// Returns the number of kernels launched.
int launch_shade(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const FrameParams& fp, const PathState& ps,
                 const OutputImages& out, Counters* cnt, int iter, int qsel, uint32_t kind_mask, uint32_t n_paths) {
  static ShadeLauncher table[kNumKinds];
  static bool init = false;
  if (!init) {
    fill_launchers(table, std::make_integer_sequence<uint32_t, kNumKinds>{});
    init = true;
  }
  const uint32_t nb = div_up(n_paths, kBinTile);
  k_bin_count<<<nb, 256, 0, s>>>(ps, cnt, iter, nb);
  k_scan_exclusive<<<1, 1024, 0, s>>>(ps.bin_hist, kNumKinds * nb);
  k_bin_scatter<<<nb, 256, 0, s>>>(ps, cnt, iter, qsel, nb);
  int launches = 3;
  for (uint32_t k = 0; k < kNumKinds; k++)
    if (kind_mask & (1u << k)) {
      table[k](s, ld.shade_blocks[k], sc, fp, ps, out, cnt, iter, qsel, nb);
      launches++;
    }
  return launches;
}
void launch_trace_shadow(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const PathState& ps, Counters* cnt,
                         int iter) {
  if (sc.single_root != 0xFFFFFFFFu) k_trace_shadow<true><<<ld.shadow_blocks_single, kTraceThreads, 0, s>>>(sc, ps, cnt, iter);
  else k_trace_shadow<false><<<ld.trace_blocks, kTraceThreads, 0, s>>>(sc, ps, cnt, iter);
}
void launch_accumulate(cudaStream_t s, const FrameParams& fp, const PathState& ps, const OutputImages& out) {
  k_accumulate<<<div_up(fp.n_pixels, 256), 256, 0, s>>>(fp, ps, out);
}
void launch_export_partial(cudaStream_t s, const OutputImages& out, float4* partial, uint32_t n, int have_accum) {
  k_export_partial<<<div_up(n, 256), 256, 0, s>>>(out, partial, n, have_accum);
}
void launch_import_partial(cudaStream_t s, const OutputImages& out, const float4* partial, uint32_t n) {
  k_import_partial<<<div_up(n, 256), 256, 0, s>>>(out, partial, n);
}
void launch_post_process(cudaStream_t s, const float4* hdr, float4* ldr, uint32_t w, uint32_t h, const AsunaPost& tm, double* block_sums) {
  const uint32_t n = w * h, sum_blocks = 296;
  if (tm.tmType == ASUNA_TM_CUSTOM && (tm.autoExposure & 1)) k_image_sum<<<sum_blocks, 256, 0, s>>>(hdr, n, block_sums);
  k_post_process<<<div_up(n, 256), 256, 0, s>>>(hdr, ldr, w, h, tm, block_sums, sum_blocks);
}
void launch_primary_rays(cudaStream_t s, const FrameParams& fp, float4* rays) {
  k_primary_rays<<<div_up(fp.n_pixels, 256), 256, 0, s>>>(fp, rays);
}
void launch_trace_user(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const float4* rays, uint32_t n, float* tuv,
                       uint32_t* inst_prim, uint8_t* occluded, Counters* cnt) {
  // the user-ray ticket lives in the last slot of the shadow tickets (never reached by a render: ASUNA_MAX_ITERS)
  uint32_t* ticket = &cnt->ticket_shadow[ASUNA_MAX_ITERS];
  cudaMemsetAsync(ticket, 0, sizeof(uint32_t), s);
  UserPolicy pol{rays, tuv, inst_prim, occluded};
  uint32_t blocks = std::min(ld.trace_blocks, std::max(1u, div_up(n, kTraceThreads)));
  const bool single = sc.single_root != 0xFFFFFFFFu;
  if (occluded) {
    if (single) k_trace_user<true, true><<<blocks, kTraceThreads, 0, s>>>(sc, pol, n, ticket, cnt);
    else k_trace_user<true, false><<<blocks, kTraceThreads, 0, s>>>(sc, pol, n, ticket, cnt);
  } else {
    if (single) k_trace_user<false, true><<<blocks, kTraceThreads, 0, s>>>(sc, pol, n, ticket, cnt);
    else k_trace_user<false, false><<<blocks, kTraceThreads, 0, s>>>(sc, pol, n, ticket, cnt);
  }
}

template <uint32_t... K>
static cudaError_t shade_occupancy(LaunchDims& ld, int sm_count, std::integer_sequence<uint32_t, K...>) {
  cudaError_t err = cudaSuccess;
  auto one = [&](auto kernel, uint32_t k) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kShadeThreads, 0);
    if (e != cudaSuccess) err = e;
    ld.shade_blocks[k] = (uint32_t)(sm_count * (per_sm > 0 ? per_sm : 1));
  };
  (one(k_shade<K>, K), ...);
  return err;
}

cudaError_t query_launch_dims(LaunchDims& ld, int sm_count) {
  int per_sm = 0;
  {
    // The single-level closest-hit kernel keeps 11 KB of prepared rays per block in shared memory, 79 KB per SM at 7
    // blocks.  Left alone the driver configures 132 KB of shared memory for it (ncu launch__shared_mem_config_size), which
    // takes 32 KB of L1 away from the node fetches; asking for the next smaller configuration (100 KB) gives them back.
    int carveout = 40;  // per cent of the 228 KB maximum
    if (const char* t = getenv("ASUNA_TRACE_CARVEOUT")) carveout = atoi(t);
    if (carveout >= 0) {
      cudaFuncSetAttribute(k_trace_closest<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
      cudaFuncSetAttribute(k_trace_closest<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
    }
  }
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_closest<false, false>, kTraceThreads, 0);
  if (e != cudaSuccess) return e;
  ld.trace_blocks = (uint32_t)(sm_count * (per_sm > 0 ? per_sm : 1));
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_closest<false, true>, kTraceThreads, 0);
  if (e != cudaSuccess) return e;
  ld.trace_blocks_single = (uint32_t)(sm_count * (per_sm > 0 ? per_sm : 1));
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_shadow<true>, kTraceThreads, 0);
  if (e != cudaSuccess) return e;
  ld.shadow_blocks_single = (uint32_t)(sm_count * (per_sm > 0 ? per_sm : 1));
  return shade_occupancy(ld, sm_count, std::make_integer_sequence<uint32_t, kNumKinds>{});
}

}  // namespace asuna
