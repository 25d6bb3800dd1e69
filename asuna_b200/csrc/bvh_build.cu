// Acceleration-structure builder, hand-written for sm_100a.
//
// Replaces the driver build behind vkCmdBuildAccelerationStructuresKHR that the reference reaches
// through nvvk::RaytracingBuilderKHR::buildBlas / buildTlas (reference
// ext/nvpro_core/nvvk/raytraceKHR_vk.cpp:77-183,302-376, called from
// src/pipeline/pipeline_raytrace.cpp:107-147).  Pipeline per BVH (all on one stream, no host sync):
//   prim boxes + bounds reduce -> 63-bit Morton keys -> LSD radix sort (8 x 8-bit, stable,
//   histogram / scan / warp-multisplit scatter) -> PLOC (Meister & Bittner 2018: parallel locally-
//   ordered agglomerative clustering along the Morton curve, one cooperative kernel) which also
//   fills, at every merge, the SAH dynamic-programming table of Ylitie et al. 2017 -> top-down
//   collapse into compressed 8-wide nodes (80 B, octant-ordered slots, 8-bit quantised child
//   boxes) with the primitive slots written in leaf order (second cooperative kernel).
// Everything is HBM-bound streaming work: loads are coalesced 16-byte where the input layout
// allows (the 44-byte vertex stride of the wire format does not), grids are sized from n.
#include <cooperative_groups.h>

#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "bvh_build.cuh"
#include "scan.cuh"

namespace cg = cooperative_groups;

namespace asuna {

namespace {

constexpr int kThreads = 256;
#ifndef ASUNA_SORT_ITEMS
#define ASUNA_SORT_ITEMS 16
#endif
#ifndef ASUNA_SORT_MIN_BLOCKS
#define ASUNA_SORT_MIN_BLOCKS 3
#endif
constexpr int kSortItems = ASUNA_SORT_ITEMS;
constexpr int kSortTile = kThreads * kSortItems;  // keys per block

// Developer instrumentation (-DASUNA_BUILD_PROFILE, tools/build_probe.py): block 0 stamps the global timer at the phase
// boundaries of the two cooperative kernels.
#ifdef ASUNA_BUILD_PROFILE
__device__ unsigned long long g_build_prof[2 * 2048];
__device__ uint32_t g_build_prof_n;
__device__ __forceinline__ void prof_stamp(uint32_t tag, uint32_t value) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    uint32_t k = g_build_prof_n++;
    if (k < 2048) g_build_prof[2 * k] = t, g_build_prof[2 * k + 1] = ((unsigned long long)tag << 32) | value;
  }
}
#define PROF(tag, value) prof_stamp(tag, value)
#else
#define PROF(tag, value)
#endif

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// ---- 1. primitive boxes + scene bounds ----------------------------------------------------
__global__ void k_bounds_init(int* bounds) {
  if (threadIdx.x < 3) bounds[threadIdx.x] = 0x7FFFFFFF;            // min = +max
  else if (threadIdx.x < 6) bounds[threadIdx.x] = (int)0x80000000;  // max = most negative ordered
}

// Centroid bounds of the primitives: warp shuffle reduce, then atomics only where they would still move the bound.  Six
// same-line atomics from every one of the million warps of a 32 M-triangle build serialise in one L2 slice (measured:
// 4.6 ms of a 28 ms build), and so do six volatile reads per warp (2.3 ms): each BLOCK reads the current bounds once
// (snapshot, at the top of the kernel so the latency hides behind the vertex loads) and its warps compare against that
// copy -- after the first few thousand blocks almost nothing improves the bounds any more.
__device__ __forceinline__ void bounds_snapshot(int* snap, const int* bounds) {
  if (threadIdx.x < 6) snap[threadIdx.x] = reinterpret_cast<const volatile int*>(bounds)[threadIdx.x];
}
__device__ __forceinline__ void reduce_bounds(float3 lo, float3 hi, bool valid, int* bounds, const int* snap) {
  float3 c = make_float3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
  float mnx = valid ? c.x : FLT_MAX, mny = valid ? c.y : FLT_MAX, mnz = valid ? c.z : FLT_MAX;
  float mxx = valid ? c.x : -FLT_MAX, mxy = valid ? c.y : -FLT_MAX, mxz = valid ? c.z : -FLT_MAX;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xFFFFFFFFu, mnx, o));
    mny = fminf(mny, __shfl_xor_sync(0xFFFFFFFFu, mny, o));
    mnz = fminf(mnz, __shfl_xor_sync(0xFFFFFFFFu, mnz, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xFFFFFFFFu, mxx, o));
    mxy = fmaxf(mxy, __shfl_xor_sync(0xFFFFFFFFu, mxy, o));
    mxz = fmaxf(mxz, __shfl_xor_sync(0xFFFFFFFFu, mxz, o));
  }
  __syncthreads();  // the snapshot is in shared memory
  if ((threadIdx.x & 31) == 0) {
    const int o0 = float_to_ordered(mnx), o1 = float_to_ordered(mny), o2 = float_to_ordered(mnz);
    const int o3 = float_to_ordered(mxx), o4 = float_to_ordered(mxy), o5 = float_to_ordered(mxz);
    if (o0 < snap[0]) atomicMin(&bounds[0], o0);
    if (o1 < snap[1]) atomicMin(&bounds[1], o1);
    if (o2 < snap[2]) atomicMin(&bounds[2], o2);
    if (o3 > snap[3]) atomicMax(&bounds[3], o3);
    if (o4 > snap[4]) atomicMax(&bounds[4], o4);
    if (o5 > snap[5]) atomicMax(&bounds[5], o5);
  }
}

__global__ void k_tri_boxes(const AsunaVertex* __restrict__ v, const uint32_t* __restrict__ idx, uint32_t n,
                            float4* __restrict__ blo, float4* __restrict__ bhi, int* bounds) {
  __shared__ int snap[6];
  bounds_snapshot(snap, bounds);
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = i < n;
  float3 lo = make_float3(0, 0, 0), hi = lo;
  if (valid) {
    const float* p0 = v[idx[3 * i + 0]].pos;
    const float* p1 = v[idx[3 * i + 1]].pos;
    const float* p2 = v[idx[3 * i + 2]].pos;
    lo = make_float3(fminf(p0[0], fminf(p1[0], p2[0])), fminf(p0[1], fminf(p1[1], p2[1])),
                     fminf(p0[2], fminf(p1[2], p2[2])));
    hi = make_float3(fmaxf(p0[0], fmaxf(p1[0], p2[0])), fmaxf(p0[1], fmaxf(p1[1], p2[1])),
                     fmaxf(p0[2], fmaxf(p1[2], p2[2])));
    blo[i] = make_float4(lo.x, lo.y, lo.z, 0.f);
    bhi[i] = make_float4(hi.x, hi.y, hi.z, 0.f);
  }
  reduce_bounds(lo, hi, valid, bounds, snap);
}

// Triangles of a single-use instance taken to world space once, at build time: the merged world-space BLAS
// (api.cu, "flattening") is built over this soup, so rays never pay a transform or a second tree for them.
// v' = o2w * v with a fixed fma order; the slot keeps (primitive id, instance id) in the two free w lanes.
__device__ __forceinline__ float4 world_vertex(const float* p, float4 r0, float4 r1, float4 r2, float w) {
  return make_float4(__fmaf_rn(r0.z, p[2], __fmaf_rn(r0.y, p[1], __fmaf_rn(r0.x, p[0], r0.w))),
                     __fmaf_rn(r1.z, p[2], __fmaf_rn(r1.y, p[1], __fmaf_rn(r1.x, p[0], r1.w))),
                     __fmaf_rn(r2.z, p[2], __fmaf_rn(r2.y, p[1], __fmaf_rn(r2.x, p[0], r2.w))), w);
}
// One launch for all instances (blockIdx.y = job): a scene of 100 instances used to pay 100 launches of ~1300 blocks
// each, every one with its own ramp-up and tail.
__global__ void __launch_bounds__(kThreads) k_world_triangles_batched(const WorldJob* __restrict__ jobs, TriSlot* __restrict__ soup,
                                                                      float4* __restrict__ blo, float4* __restrict__ bhi, int* bounds) {
  const WorldJob job = jobs[blockIdx.y];
  PROF(20, 0);
  if (blockIdx.x * kThreads >= job.n) return;  // whole block idle: no barrier or ballot is skipped by part of a warp
  __shared__ int snap[6];
  bounds_snapshot(snap, bounds);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < job.n;
  float3 lo = make_float3(0, 0, 0), hi = lo;
  if (valid) {
    const AsunaVertex* v = job.v;
    const uint32_t* idx = job.idx;
    TriSlot t;
    t.v0 = world_vertex(v[idx[3 * (size_t)i + 0]].pos, job.r0, job.r1, job.r2, __uint_as_float(i));
    t.v1 = world_vertex(v[idx[3 * (size_t)i + 1]].pos, job.r0, job.r1, job.r2, __uint_as_float(job.inst | (job.kind << 28)));
    t.v2 = world_vertex(v[idx[3 * (size_t)i + 2]].pos, job.r0, job.r1, job.r2, 0.f);
    const size_t o = (size_t)job.offset + i;
    soup[o] = t;
    lo = make_float3(fminf(t.v0.x, fminf(t.v1.x, t.v2.x)), fminf(t.v0.y, fminf(t.v1.y, t.v2.y)),
                     fminf(t.v0.z, fminf(t.v1.z, t.v2.z)));
    hi = make_float3(fmaxf(t.v0.x, fmaxf(t.v1.x, t.v2.x)), fmaxf(t.v0.y, fmaxf(t.v1.y, t.v2.y)),
                     fmaxf(t.v0.z, fmaxf(t.v1.z, t.v2.z)));
    blo[o] = make_float4(lo.x, lo.y, lo.z, 0.f);
    bhi[o] = make_float4(hi.x, hi.y, hi.z, 0.f);
  }
  reduce_bounds(lo, hi, valid, bounds, snap);
}

// World box of an instance = box of the 8 transformed corners of its mesh box (what a TLAS build sees).
__global__ void k_instance_boxes(const DInstance* __restrict__ inst, const uint32_t* __restrict__ ids,
                                 const float4* __restrict__ mesh_lo, const float4* __restrict__ mesh_hi, uint32_t n,
                                 float4* __restrict__ blo, float4* __restrict__ bhi, int* bounds) {
  __shared__ int snap[6];
  bounds_snapshot(snap, bounds);
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = i < n;
  float3 lo = make_float3(0, 0, 0), hi = lo;
  if (valid) {
    const DInstance& in = inst[ids ? ids[i] : i];
    float4 ml = mesh_lo[in.mesh], mh = mesh_hi[in.mesh];
    lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX);
    hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float px = (k & 1) ? mh.x : ml.x, py = (k & 2) ? mh.y : ml.y, pz = (k & 4) ? mh.z : ml.z;
      float x = in.o2w[0].x * px + in.o2w[0].y * py + in.o2w[0].z * pz + in.o2w[0].w;
      float y = in.o2w[1].x * px + in.o2w[1].y * py + in.o2w[1].z * pz + in.o2w[1].w;
      float z = in.o2w[2].x * px + in.o2w[2].y * py + in.o2w[2].z * pz + in.o2w[2].w;
      lo = make_float3(fminf(lo.x, x), fminf(lo.y, y), fminf(lo.z, z));
      hi = make_float3(fmaxf(hi.x, x), fmaxf(hi.y, y), fmaxf(hi.z, z));
    }
    blo[i] = make_float4(lo.x, lo.y, lo.z, 0.f);
    bhi[i] = make_float4(hi.x, hi.y, hi.z, 0.f);
  }
  reduce_bounds(lo, hi, valid, bounds, snap);
}

// ---- 2. Morton keys -----------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread21(uint32_t x) {
  uint64_t v = x & 0x1FFFFFull;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

__global__ void k_morton(const float4* __restrict__ blo, const float4* __restrict__ bhi, uint32_t n,
                         const int* __restrict__ bounds, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  PROF(21, 0);
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float3 mn = make_float3(ordered_to_float(bounds[0]), ordered_to_float(bounds[1]), ordered_to_float(bounds[2]));
  float3 mx = make_float3(ordered_to_float(bounds[3]), ordered_to_float(bounds[4]), ordered_to_float(bounds[5]));
  float ex = fmaxf(mx.x - mn.x, 1e-30f), ey = fmaxf(mx.y - mn.y, 1e-30f), ez = fmaxf(mx.z - mn.z, 1e-30f);
  float4 l = blo[i], h = bhi[i];
  const float scale = 2097151.0f;  // 2^21 - 1
  uint32_t qx = (uint32_t)fminf(fmaxf((0.5f * (l.x + h.x) - mn.x) / ex * scale, 0.f), scale);
  uint32_t qy = (uint32_t)fminf(fmaxf((0.5f * (l.y + h.y) - mn.y) / ey * scale, 0.f), scale);
  uint32_t qz = (uint32_t)fminf(fmaxf((0.5f * (l.z + h.z) - mn.z) / ez * scale, 0.f), scale);
  keys[i] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
  vals[i] = i;
}

// ---- 3. stable LSD radix sort of (key, value) pairs, 8 bits per pass ------------------------
__global__ void __launch_bounds__(kThreads) k_sort_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                                                         uint32_t* __restrict__ hist, uint32_t n_blocks) {
  PROF(22, 0);
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint32_t base = blockIdx.x * kSortTile;
  uint64_t key[kSortItems];  // one batch of loads, then the counting
#pragma unroll
  for (int k = 0; k < kSortItems; k++) {
    const uint32_t i = base + k * kThreads + threadIdx.x;
    key[k] = i < n ? keys[i] : 0ull;
  }
#pragma unroll
  for (int k = 0; k < kSortItems; k++)
    if (base + k * kThreads + threadIdx.x < n) atomicAdd(&h[(uint32_t)(key[k] >> shift) & 255u], 1u);
  __syncthreads();
  hist[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// One block per digit: exclusive scan of that digit's per-block counts in place (row d of hist), row total to
// totals[d].  The digit bases (exclusive scan of the 256 totals) are recomputed by every scatter block, so the pass
// needs no serial walk over the whole 256 x blocks table.
__global__ void __launch_bounds__(kThreads) k_sort_scan_rows(uint32_t* __restrict__ hist, uint32_t n_blocks,
                                                             uint32_t* __restrict__ totals) {
  PROF(23, 0);
  __shared__ uint32_t warp_sums[2][kThreads / 32];
  uint32_t* row = hist + (size_t)blockIdx.x * n_blocks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t carry = 0;
  int buf = 0;
  for (uint32_t base = 0; base < n_blocks; base += kThreads * 4, buf ^= 1) {
    const uint32_t i0 = base + threadIdx.x * 4;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = i0 + k < n_blocks ? row[i0 + k] : 0u, sum += v[k];
    uint32_t sc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, sc, o);
      if (lane >= o) sc += t;
    }
    if (lane == 31) warp_sums[buf][warp] = sc;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
      const uint32_t x = warp_sums[buf][w];
      before += w < warp ? x : 0u;
      all += x;
    }
    uint32_t run = carry + before + sc - sum;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < n_blocks) row[i0 + k] = run;
      run += v[k];
    }
    carry += all;
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// Scatter of one pass.  A block ranks its 4096 keys (per-warp digit counters, `match_any` inside a round), parks them in
// shared memory in digit order, and writes them out from there: consecutive threads then store consecutive addresses of a
// digit's run (16 keys on average) instead of 32 unrelated 8-byte words per instruction -- a 32.8 M-key pass was bound by
// the number of L2 write transactions, not by bytes.
constexpr size_t kScatterSmemBytes = (size_t)kSortTile * (sizeof(uint64_t) + sizeof(uint32_t));
// (3 blocks per SM: the 320 blocks of a 1.3 M-key pass are then co-resident instead of leaving 24 for a second wave)
__global__ void __launch_bounds__(kThreads, ASUNA_SORT_MIN_BLOCKS) k_sort_scatter(const uint64_t* __restrict__ keys_in,
                                                            const uint32_t* __restrict__ vals_in,
                                                            uint64_t* __restrict__ keys_out,
                                                            uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                            const uint32_t* __restrict__ hist, uint32_t n_blocks,
                                                            const uint32_t* __restrict__ totals) {
  PROF(24, 0);
  constexpr int kWarps = kThreads / 32;
  constexpr int kRounds = kSortTile / kWarps / 32;  // 16
  extern __shared__ __align__(16) unsigned char scatter_smem[];
  uint64_t* const st_key = reinterpret_cast<uint64_t*>(scatter_smem);
  uint32_t* const st_val = reinterpret_cast<uint32_t*>(st_key + kSortTile);
  __shared__ uint32_t wh[kWarps][256];
  __shared__ uint32_t goff[256];  // global position of a digit's run minus its position in the block's parked order
  __shared__ uint32_t sums[2][kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // base of digit `threadIdx.x` = number of keys with a smaller digit (exclusive scan of the 256 digit totals)
  uint32_t digit_base;
  {
    const uint32_t t = totals[threadIdx.x];
    uint32_t sc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, sc, o);
      if (lane >= o) sc += u;
    }
    if (lane == 31) sums[0][warp] = sc;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) before += w < warp ? sums[0][w] : 0u;
    digit_base = before + sc - t;
  }
  const uint32_t block_base = hist[threadIdx.x * n_blocks + blockIdx.x];  // keys of this digit in earlier blocks
  for (int b = threadIdx.x; b < kWarps * 256; b += kThreads) (&wh[0][0])[b] = 0;
  __syncthreads();
  const uint32_t tile0 = blockIdx.x * kSortTile, wbase = tile0 + warp * (kRounds * 32);
  // the warp's 16 x 32 keys and values are fetched in one batch and stay in registers; peers[r] = the lanes holding the
  // same digit in round r, kept for the second sweep
  uint64_t key[kRounds];
  uint32_t val[kRounds], peers[kRounds];
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const uint32_t i = wbase + r * 32 + lane;
    key[r] = i < n ? keys_in[i] : 0ull;
    val[r] = i < n ? vals_in[i] : 0u;
  }
  // sweep 1: per-warp digit counts (one leader lane per distinct digit per round)
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const bool valid = wbase + r * 32 + lane < n;
    const uint32_t digit = valid ? ((uint32_t)(key[r] >> shift) & 255u) : (256u + lane);
    peers[r] = __match_any_sync(0xFFFFFFFFu, digit);
    if (valid && lane == __ffs(peers[r]) - 1) wh[warp][digit] += __popc(peers[r]);
    __syncwarp();
  }
  __syncthreads();
  // per digit: where its run starts in the parked order (exclusive scan of the block's digit counts), the same for
  // every warp of the block, and the offset that takes a parked index to its global position
  {
    const uint32_t bin = threadIdx.x;
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) cnt += wh[w][bin];
    uint32_t sc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, sc, o);
      if (lane >= o) sc += u;
    }
    if (lane == 31) sums[1][warp] = sc;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) before += w < warp ? sums[1][w] : 0u;
    uint32_t run = before + sc - cnt;  // parked position of the digit's first key
    goff[bin] = digit_base + block_base - run;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      const uint32_t c = wh[w][bin];
      wh[w][bin] = run;
      run += c;
    }
  }
  __syncthreads();
  // sweep 2: rank within the warp round, park
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const bool valid = wbase + r * 32 + lane < n;
    const uint32_t digit = (uint32_t)(key[r] >> shift) & 255u;
    if (valid) {
      const uint32_t pos = wh[warp][digit] + __popc(peers[r] & ((1u << lane) - 1u));
      st_key[pos] = key[r];
      st_val[pos] = val[r];
    }
    __syncwarp();
    if (valid && lane == __ffs(peers[r]) - 1) wh[warp][digit] += __popc(peers[r]);
    __syncwarp();
  }
  __syncthreads();
  const uint32_t tile_n = min((uint32_t)kSortTile, n - tile0);
#pragma unroll 4
  for (uint32_t i = threadIdx.x; i < tile_n; i += kThreads) {
    const uint64_t k = st_key[i];
    const uint32_t pos = goff[(uint32_t)(k >> shift) & 255u] + i;
    keys_out[pos] = k;
    vals_out[pos] = st_val[i];
  }
}

// ---- 4. PLOC: binary hierarchy by locally-ordered clustering + SAH collapse table ------------
// Binary-node numbering during the build: leaf j (sorted position) is n-1+j; inner nodes are handed
// out from n-2 downwards in merge order, so the last merge -- the root -- is node 0.
#ifndef ASUNA_PLOC_RADIUS
#define ASUNA_PLOC_RADIUS 10
#endif
constexpr int kPlocRadius = ASUNA_PLOC_RADIUS;
#ifndef ASUNA_PLOC_TAIL
#define ASUNA_PLOC_TAIL 2048  // measured 1024 / 2048 / 4096: 1.34 / 1.325 / 1.36 ms at 1.31 M triangles
#endif
constexpr int kPlocTail = ASUNA_PLOC_TAIL;
#ifndef ASUNA_PLOC_MIN_BLOCKS
#define ASUNA_PLOC_MIN_BLOCKS 4  // 64 registers: 32 resident warps per SM instead of 24 at the natural 75
#endif  // clusters left when one block takes over (see k_ploc)
constexpr uint64_t kDecLeaf = 1ull;

struct PlocParams {
  int n;
  float cost_node, cost_prim;
  const float4* blo;   // primitive boxes
  const float4* bhi;
  const uint32_t* order;  // sorted position -> primitive id
  float4* nlo;         // [2n-1] binary-node boxes, lo.w = primitive count (bits)
  float4* nhi;
  int2* children;      // [n-1]
  float* cost;         // [2n-1][7]  C(node, i): cheapest forest of <= i wide-BVH roots
  uint64_t* dec;       // [2n-1]    argmin bookkeeping of the table
  int* cid[2];         // cluster arrays (ping-pong): binary node id ...
  float4* clo[2];      // ... and its box, kept in cluster order so the neighbour search streams
  float4* chi[2];
  int* nn;             // nearest neighbour of cluster i within the radius
  uint2* pre;          // per-cluster exclusive prefix inside its block chunk {survivors, merges}
  uint2* block_sums;   // per block {survivors, merges}
  int* state;          // {clusters left, current cluster array, next inner node}: handed from k_ploc to k_ploc_tail
};

__device__ __forceinline__ float half_area(float lx, float hx, float ly, float hy, float lz, float hz) {
  float ex = hx - lx, ey = hy - ly, ez = hz - lz;
  return ex * ey + ey * ez + ez * ex;
}
__device__ __forceinline__ float union_half_area(float4 alo, float4 ahi, float4 blo, float4 bhi) {
  return half_area(fminf(alo.x, blo.x), fmaxf(ahi.x, bhi.x), fminf(alo.y, blo.y), fmaxf(ahi.y, bhi.y),
                   fminf(alo.z, blo.z), fmaxf(ahi.z, bhi.z));
}

// Table of Ylitie et al. 2017 section 4.1 for the merged node `idx` = (l, r):
//   C(n,1) = min( leaf: A P c_prim  if P <= 3 ,  inner: A c_node + min_k C(l,k) + C(r,8-k) )
//   C(n,i) = min( min_k C(l,k) + C(r,i-k) , C(n,i-1) )          i = 2..7
// dec: bit 0 = "C(n,1) is a leaf"; 6 bits per i = 2..8 at 4 + 6 (i-2): (k_left, k_right), 0 = the node itself.
// The row of a binary leaf is seven times (its area x c_prim); it is never stored, the merge that consumes it gets the
// value from the box it holds anyway (a 32.8 M-triangle build saved 1.2 GB of writes and as many scattered reads).
__device__ void dp_merge(const PlocParams& a, int idx, int l, int r, float area, uint32_t count, float leaf_l, float leaf_r) {
  float cl[7], cr[7], c[7];
  const bool is_leaf_l = l >= a.n - 1, is_leaf_r = r >= a.n - 1;
#pragma unroll
  for (int i = 0; i < 7; i++) {
    cl[i] = is_leaf_l ? leaf_l : a.cost[(size_t)l * 7 + i];
    cr[i] = is_leaf_r ? leaf_r : a.cost[(size_t)r * 7 + i];
  }
  float best8 = FLT_MAX;
  uint32_t k8 = 1;
#pragma unroll
  for (int k = 1; k <= 7; k++) {
    float v = cl[k - 1] + cr[7 - k];
    if (v < best8) best8 = v, k8 = k;
  }
  float c_inner = best8 + area * a.cost_node;
  float c_leaf = count <= (uint32_t)kMaxLeafPrims ? area * (float)count * a.cost_prim : FLT_MAX;
  uint64_t dec = c_leaf <= c_inner ? kDecLeaf : 0ull;
  c[0] = fminf(c_leaf, c_inner);
  dec |= (uint64_t)(k8 | ((8 - k8) << 3)) << (4 + 6 * 6);
  uint32_t prev = 0;
#pragma unroll
  for (int i = 2; i <= 7; i++) {
    float best = FLT_MAX;
    uint32_t kb = 1;
    for (int k = 1; k < i; k++) {
      float v = cl[k - 1] + cr[i - k - 1];
      if (v < best) best = v, kb = k;
    }
    if (best < c[i - 2]) c[i - 1] = best, prev = kb | ((i - kb) << 3);
    else c[i - 1] = c[i - 2];
    dec |= (uint64_t)prev << (4 + 6 * (i - 2));
  }
#pragma unroll
  for (int i = 0; i < 7; i++) a.cost[(size_t)idx * 7 + i] = c[i];
  a.dec[idx] = dec;
}

// exclusive scan of one packed counter pair (two 16-bit fields) over a 256-thread block
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t& total, uint32_t* warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
    if (lane >= o) s += t;
  }
  __syncthreads();  // previous use of warp_sums is over
  if (lane == 31) warp_sums[warp] = s;
  __syncthreads();
  uint32_t before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; w++) {
    uint32_t x = warp_sums[w];
    if (w < warp) before += x;
    tot += x;
  }
  total = tot;
  return before + s - v;
}

// The leaf clusters of the build: binary node n-1+j for sorted position j.  Only the box is stored (its collapse-table row
// is synthesised by dp_merge, its decision word is never read); a one-primitive BVH has no merge, so its row is written.
__device__ __forceinline__ void ploc_leaf(const PlocParams& a, int j, float4& lo, float4& hi, int& node) {
  const uint32_t prim = a.order[j];
  lo = a.blo[prim], hi = a.bhi[prim];
  lo.w = __uint_as_float(1u);
  hi.w = 0.f;
  node = a.n - 1 + j;
  a.nlo[node] = lo;
  a.nhi[node] = hi;
  if (a.n == 1) {
    const float c = half_area(lo.x, hi.x, lo.y, hi.y, lo.z, hi.z) * a.cost_prim;
#pragma unroll
    for (int i = 0; i < 7; i++) a.cost[(size_t)node * 7 + i] = c;
    a.dec[node] = kDecLeaf;
  }
}

// The new inner node of a merge (numbered `idx`), its box and collapse-table row; returns the merged cluster.
__device__ __forceinline__ void ploc_merge(const PlocParams& a, int idx, float4& lo, float4& hi, int& id, float4 lo2, float4 hi2,
                                           int id2) {
  const uint32_t count = __float_as_uint(lo.w) + __float_as_uint(lo2.w);
  const float leaf_l = half_area(lo.x, hi.x, lo.y, hi.y, lo.z, hi.z) * a.cost_prim;
  const float leaf_r = half_area(lo2.x, hi2.x, lo2.y, hi2.y, lo2.z, hi2.z) * a.cost_prim;
  const int id1 = id;
  lo = make_float4(fminf(lo.x, lo2.x), fminf(lo.y, lo2.y), fminf(lo.z, lo2.z), __uint_as_float(count));
  hi = make_float4(fmaxf(hi.x, hi2.x), fmaxf(hi.y, hi2.y), fmaxf(hi.z, hi2.z), 0.f);
  a.children[idx] = make_int2(id, id2);
  a.nlo[idx] = lo;
  a.nhi[idx] = hi;
  dp_merge(a, idx, id1, id2, half_area(lo.x, hi.x, lo.y, hi.y, lo.z, hi.z), count, leaf_l, leaf_r);
  id = idx;
}

// Grid rounds.  One round = two grid-wide barriers:
//   A. every block walks its contiguous chunk of the cluster array in tiles of kThreads clusters.  The boxes of a tile and
//      of 2 r neighbours on either side go to shared memory, each pair distance inside the radius is evaluated once
//      (d(i, i+k), k = 1..r; the backward distances of a cluster are its predecessors' forward ones), the nearest
//      neighbour of the tile and of r clusters on either side follows, and with it the tile's merge flags and their
//      prefix inside the chunk -- no global round trip between the search and the flags;
//   B. global offsets from the block sums, then merge / copy into the other cluster array (Morton order is kept).
// The search visits the neighbours in ascending index order with a strict `<`, so ties go to the lowest index and a
// mutual pair always exists.  The rounds below kPlocTail clusters belong to k_ploc_tail.
constexpr int kPlocTileBoxes = kThreads + 4 * kPlocRadius;  // t0 - 2r .. t1 + 2r
constexpr int kPlocTileDist = kThreads + 3 * kPlocRadius;   // forward distances of t0 - 2r .. t1 + r
constexpr int kPlocTileNn = kThreads + 2 * kPlocRadius;     // nearest neighbours of t0 - r .. t1 + r

__global__ void __launch_bounds__(kThreads, ASUNA_PLOC_MIN_BLOCKS) k_ploc(const PlocParams a) {
  cg::grid_group grid = cg::this_grid();
  PROF(25, 0);
  __shared__ uint32_t warp_sums[kThreads / 32];
  __shared__ uint32_t red[4];
  __shared__ float4 s_lo[kPlocTileBoxes], s_hi[kPlocTileBoxes];
  __shared__ float s_fd[kPlocRadius][kPlocTileDist];
  __shared__ int s_nn[kPlocTileNn];
  const int n = a.n;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  for (uint32_t j = gtid; j < (uint32_t)n; j += gsize) {
    float4 lo, hi;
    int node;
    ploc_leaf(a, (int)j, lo, hi, node);
  }
  grid.sync();
  PROF(1, n);
  int m = n, cur = 0, next_inner = n - 2;
  const uint32_t nb = gridDim.x, bid = blockIdx.x;
  bool first = true;  // the clusters of the first round are the leaves themselves: boxes nlo / nhi[n-1 ..), ids n-1+i
  while (m > kPlocTail) {
    const int* cid = first ? nullptr : a.cid[cur];
    const float4* clo = first ? a.nlo + (n - 1) : a.clo[cur];
    const float4* chi = first ? a.nhi + (n - 1) : a.chi[cur];
    const int chunk = (m + (int)nb - 1) / (int)nb;
    const int c0 = min(m, (int)bid * chunk), c1 = min(m, c0 + chunk);
    // ---- A: nearest neighbours, merge flags, prefix inside the chunk
    uint32_t run_valid = 0, run_lead = 0;
    for (int t0 = c0; t0 < c1; t0 += kThreads) {
      const int t1 = min(c1, t0 + kThreads), base = t0 - 2 * kPlocRadius;
      const int n_box = t1 + 2 * kPlocRadius - base, n_dist = t1 + kPlocRadius - base;
      for (int sidx = (int)threadIdx.x; sidx < n_box; sidx += kThreads) {
        const int g = base + sidx;
        if (g >= 0 && g < m) s_lo[sidx] = clo[g], s_hi[sidx] = chi[g];
      }
      __syncthreads();
      for (int sidx = (int)threadIdx.x; sidx < n_dist; sidx += kThreads) {
        const int g = base + sidx;
        if (g < 0 || g >= m) continue;
        const float4 lo = s_lo[sidx], hi = s_hi[sidx];
#pragma unroll
        for (int k = 1; k <= kPlocRadius; k++)
          s_fd[k - 1][sidx] = g + k < m ? union_half_area(lo, hi, s_lo[sidx + k], s_hi[sidx + k]) : FLT_MAX;
      }
      __syncthreads();
      for (int sidx = kPlocRadius + (int)threadIdx.x; sidx < n_dist; sidx += kThreads) {
        const int g = base + sidx;
        if (g < 0 || g >= m) continue;
        float best = FLT_MAX;
        int bj = g == 0 ? 1 : max(0, g - kPlocRadius);
#pragma unroll
        for (int k = kPlocRadius; k >= 1; k--) {
          if (g - k < 0) continue;
          const float d = s_fd[k - 1][sidx - k];
          if (d < best) best = d, bj = g - k;
        }
#pragma unroll
        for (int k = 1; k <= kPlocRadius; k++) {
          const float d = s_fd[k - 1][sidx];  // FLT_MAX past the end of the array
          if (d < best) best = d, bj = g + k;
        }
        s_nn[sidx - kPlocRadius] = bj;
      }
      __syncthreads();
      const int i = t0 + (int)threadIdx.x;
      uint32_t packed = 0;
      int j = 0;
      bool mutual = false;
      if (i < t1) {
        j = s_nn[i - base - kPlocRadius];
        mutual = s_nn[j - base - kPlocRadius] == i;
        packed = ((!mutual || i < j) ? 1u : 0u) | ((mutual && i < j) ? 0x10000u : 0u);
      }
      uint32_t total;
      const uint32_t excl = block_scan_excl(packed, total, warp_sums);
      if (i < t1) {
        a.pre[i] = make_uint2(run_valid + (excl & 0xFFFFu), run_lead + (excl >> 16));
        a.nn[i] = mutual ? (j | (int)0x80000000) : j;
      }
      run_valid += total & 0xFFFFu;
      run_lead += total >> 16;
    }
    if (threadIdx.x == 0) a.block_sums[bid] = make_uint2(run_valid, run_lead);
    grid.sync();
    PROF(3, m);
    // ---- B: global offsets from the block sums, then merge / copy into the next cluster array
    {
      uint32_t bv = 0, bl = 0, tv = 0, tl = 0;
      for (uint32_t b = threadIdx.x; b < nb; b += kThreads) {
        uint2 sm = a.block_sums[b];
        tv += sm.x, tl += sm.y;
        if (b < bid) bv += sm.x, bl += sm.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        bv += __shfl_xor_sync(0xFFFFFFFFu, bv, o);
        bl += __shfl_xor_sync(0xFFFFFFFFu, bl, o);
        tv += __shfl_xor_sync(0xFFFFFFFFu, tv, o);
        tl += __shfl_xor_sync(0xFFFFFFFFu, tl, o);
      }
      __syncthreads();
      if (threadIdx.x < 4) red[threadIdx.x] = 0;
      __syncthreads();
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&red[0], bv);
        atomicAdd(&red[1], bl);
        atomicAdd(&red[2], tv);
        atomicAdd(&red[3], tl);
      }
      __syncthreads();
    }
    const uint32_t base_valid = red[0], base_lead = red[1], tot_valid = red[2], tot_lead = red[3];
    int* ocid = a.cid[cur ^ 1];
    float4* oclo = a.clo[cur ^ 1];
    float4* ochi = a.chi[cur ^ 1];
    for (int i = c0 + (int)threadIdx.x; i < c1; i += kThreads) {
      const int jn = a.nn[i];
      const bool mutual = jn < 0;
      const int j = jn & 0x7FFFFFFF;
      if (mutual && i > j) continue;
      const uint2 pr = a.pre[i];
      float4 lo = clo[i], hi = chi[i];
      int id = cid ? cid[i] : n - 1 + i;
      if (mutual) ploc_merge(a, next_inner - (int)(base_lead + pr.y), lo, hi, id, clo[j], chi[j], cid ? cid[j] : n - 1 + j);
      const uint32_t pos = base_valid + pr.x;
      ocid[pos] = id;
      oclo[pos] = lo;
      ochi[pos] = hi;
    }
    grid.sync();
    PROF(4, m);
    m = (int)tot_valid;
    next_inner -= (int)tot_lead;
    cur ^= 1;
    first = false;
  }
  if (gtid == 0) a.state[0] = m, a.state[1] = cur, a.state[2] = next_inner;
}

// The last rounds (half of all rounds happen below a few thousand clusters) are pure synchronisation latency on a grid,
// so ONE block of 1024 threads finishes them with the cluster array -- boxes and node ids -- in shared memory, in place:
// a round is the neighbour search, the flags and their block-wide prefix over a blocked assignment (thread t owns the
// clusters t E .. t E + E - 1, which keeps Morton order), the merges into registers, a barrier, and the compacted write.
// A BVH of at most kPlocTail primitives is built here from its leaves without the grid kernel.
constexpr int kTailThreads = 1024;
constexpr int kTailPer = kPlocTail / kTailThreads;
static_assert(kPlocTail % kTailThreads == 0 && kTailPer >= 1 && kTailPer <= 8, "tail clusters per thread");
constexpr size_t kTailSmemBytes = (size_t)kPlocTail * (2 * sizeof(float4) + 2 * sizeof(int));

__global__ void __launch_bounds__(kTailThreads, 1) k_ploc_tail(const PlocParams a, int from_leaves) {
  extern __shared__ __align__(16) unsigned char tail_smem[];
  float4* const s_lo = reinterpret_cast<float4*>(tail_smem);
  float4* const s_hi = s_lo + kPlocTail;
  int* const s_id = reinterpret_cast<int*>(s_hi + kPlocTail);
  int* const s_nn = s_id + kPlocTail;
  __shared__ uint32_t warp_sums[32];
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  PROF(26, 0);
  int m, next_inner;
  if (from_leaves) {
    m = a.n, next_inner = a.n - 2;
    for (int j = tid; j < m; j += kTailThreads) ploc_leaf(a, j, s_lo[j], s_hi[j], s_id[j]);
  } else {
    m = a.state[0], next_inner = a.state[2];
    const int cur = a.state[1];
    for (int j = tid; j < m; j += kTailThreads) s_lo[j] = a.clo[cur][j], s_hi[j] = a.chi[cur][j], s_id[j] = a.cid[cur][j];
  }
  __syncthreads();
  PROF(5, m);
  while (m > 1) {
    for (int i = tid; i < m; i += kTailThreads) {
      const float4 lo = s_lo[i], hi = s_hi[i];
      float best = FLT_MAX;
      int bj = i == 0 ? 1 : max(0, i - kPlocRadius);
#pragma unroll
      for (int k = -kPlocRadius; k <= kPlocRadius; k++) {  // ascending index, strict <: ties go to the lowest index
        if (k == 0) continue;
        const int j = i + k;
        if (j < 0 || j >= m) continue;
        const float d = union_half_area(lo, hi, s_lo[j], s_hi[j]);
        if (d < best) best = d, bj = j;
      }
      s_nn[i] = bj;
    }
    __syncthreads();
    PROF(2, m);
    const int per = (m + kTailThreads - 1) / kTailThreads;
    const int i0 = min(m, tid * per);
    int partner[kTailPer];  // >= 0: merge with it, -1: survive alone, -2: nothing (absorbed or out of range)
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < kTailPer; k++) {
      const int i = i0 + k;
      partner[k] = -2;
      if (k < per && i < m) {
        const int j = s_nn[i];
        const bool mutual = s_nn[j] == i;
        if (!mutual) partner[k] = -1, packed += 1u;
        else if (i < j) partner[k] = j, packed += 0x10001u;
      }
    }
    // exclusive prefix of the packed {survivors, merges} pair over the block
    uint32_t incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t x = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += t;
      }
      warp_sums[lane] = x;
    }
    __syncthreads();
    const uint32_t total = warp_sums[31];
    const uint32_t excl = (warp ? warp_sums[warp - 1] : 0u) + incl - packed;
    uint32_t run_valid = excl & 0xFFFFu, run_lead = excl >> 16;
    float4 rlo[kTailPer], rhi[kTailPer];
    int rid[kTailPer];
#pragma unroll
    for (int k = 0; k < kTailPer; k++) {
      if (partner[k] == -2) continue;
      const int i = i0 + k;
      rlo[k] = s_lo[i], rhi[k] = s_hi[i], rid[k] = s_id[i];
      if (partner[k] >= 0) {
        const int j = partner[k];
        ploc_merge(a, next_inner - (int)run_lead, rlo[k], rhi[k], rid[k], s_lo[j], s_hi[j], s_id[j]);
        run_lead++;
      }
    }
    __syncthreads();  // everybody has read its clusters: the array is compacted in place
#pragma unroll
    for (int k = 0; k < kTailPer; k++) {
      if (partner[k] == -2) continue;
      s_lo[run_valid] = rlo[k], s_hi[run_valid] = rhi[k], s_id[run_valid] = rid[k];
      run_valid++;
    }
    __syncthreads();
    PROF(4, m);
    m = (int)(total & 0xFFFFu);
    next_inner -= (int)(total >> 16);
  }
}

// ---- 5. collapse into compressed 8-wide nodes, level by level ---------------------------------
struct EmitParams {
  int n;
  const int2* children;
  const float4* nlo;
  const float4* nhi;
  const uint64_t* dec;
  const float* cost;
  const uint32_t* order;   // sorted position -> primitive id
  WideNode* nodes;         // absolute array
  uint32_t node_base;      // this BVH's first wide node (its root)
  uint32_t prim_base;      // this BVH's first primitive slot
  int* root_of;            // [n] binary subtree root of wide node (node_base + i)
  uint32_t* counters;      // [0] wide nodes handed out, [1] primitive slots handed out, [2..4] level ends (see k_emit_wide)
  uint32_t* slot_pos;      // [n] primitive slot -> sorted position
  // primitive payload: triangles of a mesh, or instance ids of the top level
  const AsunaVertex* v;
  const uint32_t* idx;
  const TriSlot* soup;      // world-space triangles of the merged BLAS (then v / idx are unused)
  const uint32_t* prim_ids; // top level: primitive -> instance index (nullptr = identity)
  TriSlot* tris;
  uint32_t* leaf_inst;
  // results for the host / the top-level build
  float4* out_lo;
  float4* out_hi;
  BuildResult* result;
};

__device__ void emit_prim(const EmitParams& a, uint32_t slot, uint32_t prim) {
  if (a.soup) {
    a.tris[slot] = a.soup[prim];
  } else if (a.tris) {
    const float* p0 = a.v[a.idx[3 * (size_t)prim + 0]].pos;
    const float* p1 = a.v[a.idx[3 * (size_t)prim + 1]].pos;
    const float* p2 = a.v[a.idx[3 * (size_t)prim + 2]].pos;
    TriSlot t;
    t.v0 = make_float4(p0[0], p0[1], p0[2], __uint_as_float(prim));
    t.v1 = make_float4(p1[0], p1[1], p1[2], 0.f);
    t.v2 = make_float4(p2[0], p2[1], p2[2], 0.f);
    a.tris[slot] = t;
  } else {
    a.leaf_inst[slot] = a.prim_ids ? a.prim_ids[prim] : prim;
  }
}

// One wide node.  Called by whole warps (`valid` = this lane has a node): node and primitive-slot ranges are handed out
// with ONE atomic per warp per counter -- a 1.3 M-triangle build emits 200 k wide nodes, and 400 k same-address atomics
// cost more than all the arithmetic of this kernel.
//
// A level of the wide tree takes as long as its slowest thread, and a thread's time is the LENGTH of its chain of
// dependent scattered loads (~0.7 us each), not its bytes: the depth-first form of this function (walk the collapse
// subtree node by node, then child boxes, then every leaf child's primitives one after another) chained 40-60 of them.
// Here every stage issues the loads of all <= 8 children together -- the collapse walk expands breadth-first (one
// latency per level of the collapse subtree, typically 3), the boxes are one latency, the <= 3 primitives of all leaf
// children two -- and the primitive payload itself is moved by a separate streaming pass (k_emit_wide's last loop):
// about ten latencies per node.  Every array indexed by the child number is indexed by compile-time constants (fully
// unrolled), so the loaded values stay in registers and no load is followed by a local-memory store that would wait for it.
__device__ void emit_wide_node(const EmitParams& a, uint32_t w, bool valid, uint32_t* next_level_end) {
  static_assert(kMaxLeafPrims == 3, "the leaf walk below enumerates subtrees of at most 3 primitives");
  const int n = a.n;
  const int root = valid ? a.root_of[w] : 0;
  // ---- 1. the <= 8 children: (node, budget) entries in left-to-right order; budget -1 = settled inner child, -2 = settled leaf
  int ln[8], lb[8];
  int nc = 0;
  float4 rlo = make_float4(0.f, 0.f, 0.f, 0.f), rhi = rlo;
  if (valid) {
    const uint64_t droot = a.dec[root];
    const int2 c = root < n - 1 ? a.children[root] : make_int2(0, 0);
    rlo = a.nlo[root], rhi = a.nhi[root];
    if (root >= n - 1 || (droot & kDecLeaf)) {
      ln[0] = root, lb[0] = -2, nc = 1;  // a BVH of <= 3 primitives: one leaf child
    } else {
      const uint32_t k = (uint32_t)(droot >> (4 + 6 * 6)) & 63u;
      ln[0] = c.x, lb[0] = c.x >= n - 1 ? -2 : (int)(k & 7u);
      ln[1] = c.y, lb[1] = c.y >= n - 1 ? -2 : (int)(k >> 3);
      nc = 2;
    }
  }
  for (bool again = valid; again;) {
    int x[8], bud[8];
    uint64_t d[8];
    int2 cc[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      x[e] = e < nc ? ln[e] : 0, bud[e] = e < nc ? lb[e] : -1;
      d[e] = 0ull, cc[e] = make_int2(0, 0);
    }
#pragma unroll
    for (int e = 0; e < 8; e++)
      if (bud[e] >= 0) d[e] = a.dec[x[e]], cc[e] = a.children[x[e]];  // an open entry is never a binary leaf
    int c2 = 0;
    again = false;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      if (e >= nc) continue;
      if (bud[e] < 0) {
        ln[c2] = x[e], lb[c2] = bud[e], c2++;
        continue;
      }
      const uint32_t code = bud[e] >= 2 ? (uint32_t)(d[e] >> (4 + 6 * (bud[e] - 2))) & 63u : 0u;
      if (code == 0u) {
        ln[c2] = x[e], lb[c2] = (d[e] & kDecLeaf) ? -2 : -1, c2++;
      } else {
        const bool open_l = cc[e].x < n - 1, open_r = cc[e].y < n - 1;
        ln[c2] = cc[e].x, lb[c2] = open_l ? (int)(code & 7u) : -2, c2++;
        ln[c2] = cc[e].y, lb[c2] = open_r ? (int)(code >> 3) : -2, c2++;
        again = again || open_l || open_r;
      }
    }
    nc = c2;
  }
  int node[8];
  bool leaf[8];
#pragma unroll
  for (int i = 0; i < 8; i++) node[i] = i < nc ? ln[i] : 0, leaf[i] = i < nc && lb[i] == -2;
  // ---- 2. child boxes (lo.w = primitive count of the subtree), all in flight together
  float4 clo[8], chi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) clo[i] = chi[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (i < nc) clo[i] = a.nlo[node[i]], chi[i] = a.nhi[node[i]];

  // ---- 3. octant-ordered slots: greedy assignment maximising sum dot(child centre - node centre, dir(slot)),
  // dir(slot) = +1 on the axes whose slot bit is set (x = bit 0), so slot ^ ray-octant is front to back
  int slot_of[8];
  {
    const float cx = 0.5f * (rlo.x + rhi.x), cy = 0.5f * (rlo.y + rhi.y), cz = 0.5f * (rlo.z + rhi.z);
    float dx[8], dy[8], dz[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      dx[i] = 0.5f * (clo[i].x + chi[i].x) - cx, dy[i] = 0.5f * (clo[i].y + chi[i].y) - cy;
      dz[i] = 0.5f * (clo[i].z + chi[i].z) - cz;
      slot_of[i] = 0;
    }
    // The best free pair overall is the best over the free children of each child's best free slot, so every child
    // keeps its row maximum and only the rows that pointed at the slot just taken are searched again: the picks, ties
    // included (lowest child, then lowest slot), are those of the exhaustive 8 x 64 search at a third of its instructions.
    auto row_best = [&](int i, uint32_t used, float& best, int& bs) {
      best = -FLT_MAX, bs = 0;
#pragma unroll
      for (int sl = 0; sl < 8; sl++) {
        const float score = ((sl & 1) ? dx[i] : -dx[i]) + ((sl & 2) ? dy[i] : -dy[i]) + ((sl & 4) ? dz[i] : -dz[i]);
        if (!((used >> sl) & 1u) && score > best) best = score, bs = sl;
      }
    };
    float row_score[8];
    int row_slot[8];
#pragma unroll
    for (int i = 0; i < 8; i++) row_best(i, 0u, row_score[i], row_slot[i]);
    uint32_t slot_used = 0, child_done = nc >= 8 ? 0u : (0xFFu << nc) & 0xFFu;
#pragma unroll
    for (int round = 0; round < 8; round++) {
      if (round < nc) {
        float best = -FLT_MAX;
        int bi = 0, bs = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (!((child_done >> i) & 1u) && row_score[i] > best) best = row_score[i], bi = i, bs = row_slot[i];
        slot_used |= 1u << bs;
        child_done |= 1u << bi;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (i == bi) slot_of[i] = bs;
          if (!((child_done >> i) & 1u) && row_slot[i] == bs) row_best(i, slot_used, row_score[i], row_slot[i]);
        }
      }
    }
  }
  uint32_t imask = 0, n_prims = 0;
  uint32_t cnt[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    cnt[i] = leaf[i] ? __float_as_uint(clo[i].w) : 0u;
    n_prims += cnt[i];
    if (i < nc && !leaf[i]) imask |= 1u << slot_of[i];
  }
  const uint32_t n_inner = __popc(imask);
  uint32_t cb, pb;
  {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t xi = n_inner, xp = n_prims;  // inclusive warp scans (both are 0 on lanes without a node)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t ti = __shfl_up_sync(0xFFFFFFFFu, xi, o), tp = __shfl_up_sync(0xFFFFFFFFu, xp, o);
      if (lane >= (uint32_t)o) xi += ti, xp += tp;
    }
    const uint32_t tot_i = __shfl_sync(0xFFFFFFFFu, xi, 31), tot_p = __shfl_sync(0xFFFFFFFFu, xp, 31);
    uint32_t bi = 0, bp = 0;
    if (lane == 0) {
      if (tot_i) {
        bi = atomicAdd(&a.counters[0], tot_i);
        atomicMax(next_level_end, bi + tot_i);  // the counter only grows: the maximum is where the next level ends
      }
      if (tot_p) bp = atomicAdd(&a.counters[1], tot_p);
    }
    cb = __shfl_sync(0xFFFFFFFFu, bi, 0) + xi - n_inner;
    pb = __shfl_sync(0xFFFFFFFFu, bp, 0) + xp - n_prims;
  }
  if (!valid) return;

  // ---- 4. sorted positions of the primitives of the leaf children, subtree order (left first): a subtree of 2 is
  // (leaf, leaf), one of 3 is (leaf, (leaf, leaf)) or ((leaf, leaf), leaf)
  int2 c1[8], c2[8];
#pragma unroll
  for (int i = 0; i < 8; i++) c1[i] = c2[i] = make_int2(0, 0);
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (cnt[i] >= 2u) c1[i] = a.children[node[i]];
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (cnt[i] == 3u) c2[i] = a.children[c1[i].x < n - 1 ? c1[i].x : c1[i].y];

  // ---- 5. quantisation grid: origin = padded lower corner, per-axis power-of-two scale covering the padded extent
  float pad[3], p[3], inv_scale[3];
  uint32_t e[3];
  {
    const float lo3[3] = {rlo.x, rlo.y, rlo.z}, hi3[3] = {rhi.x, rhi.y, rhi.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      pad[k] = fmaxf(fabsf(lo3[k]), fabsf(hi3[k])) * 2.4e-7f + 1e-30f;
      p[k] = lo3[k] - pad[k];
      float f = ((hi3[k] + pad[k]) - p[k]) * (1.0f / 254.0f);
      uint32_t bits = __float_as_uint(f);
      uint32_t eb = (bits >> 23) + ((bits & 0x7FFFFFu) ? 1u : 0u);
      eb = min(max(eb, 1u), 253u);
      e[k] = eb;
      inv_scale[k] = __uint_as_float((254u - eb) << 23);
    }
  }
  // one byte per slot in each 64-bit plane; empty slots keep lo = 255, hi = 0, meta = 0
  uint64_t qlo[3] = {~0ull, ~0ull, ~0ull}, qhi[3] = {0ull, 0ull, 0ull}, meta = 0ull;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i >= nc) continue;
    const uint32_t sh = 8u * (uint32_t)slot_of[i];
    const float lo3[3] = {clo[i].x, clo[i].y, clo[i].z}, hi3[3] = {chi[i].x, chi[i].y, chi[i].z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float fl = floorf((lo3[k] - pad[k] - p[k]) * inv_scale[k]);
      const float fh = ceilf((hi3[k] + pad[k] - p[k]) * inv_scale[k]);
      const uint64_t ql = (uint8_t)fminf(fmaxf(fl, 0.f), 255.f), qh = (uint8_t)fminf(fmaxf(fh, 0.f), 255.f);
      qlo[k] = (qlo[k] & ~(0xFFull << sh)) | (ql << sh);
      qhi[k] |= qh << sh;
    }
    uint32_t m;
    if (!leaf[i]) {
      // inner children are numbered in slot order (the traversal ranks a hit by the inner slots below it)
      m = 0x20u | (24u + (uint32_t)slot_of[i]);
      a.root_of[cb + __popc(imask & ((1u << slot_of[i]) - 1u))] = node[i];
    } else {
      uint32_t prim_off = 0;  // primitives of the leaf children in lower slots
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (k != i && slot_of[k] < slot_of[i]) prim_off += cnt[k];
      m = (((1u << cnt[i]) - 1u) << 5) | prim_off;
      uint32_t* sp = a.slot_pos + pb + prim_off;
      const int first = n - 1;
      if (cnt[i] == 1u) {
        sp[0] = (uint32_t)(node[i] - first);
      } else if (cnt[i] == 2u) {
        sp[0] = (uint32_t)(c1[i].x - first), sp[1] = (uint32_t)(c1[i].y - first);
      } else if (c1[i].x < first) {
        sp[0] = (uint32_t)(c2[i].x - first), sp[1] = (uint32_t)(c2[i].y - first), sp[2] = (uint32_t)(c1[i].y - first);
      } else {
        sp[0] = (uint32_t)(c1[i].x - first), sp[1] = (uint32_t)(c2[i].x - first), sp[2] = (uint32_t)(c2[i].y - first);
      }
    }
    meta |= (uint64_t)m << sh;
  }
  // five 16-byte stores
  uint4* out = reinterpret_cast<uint4*>(a.nodes + a.node_base + w);
  out[0] = make_uint4(__float_as_uint(p[0]), __float_as_uint(p[1]), __float_as_uint(p[2]), e[0] | (e[1] << 8) | (e[2] << 16) | (imask << 24));
  out[1] = make_uint4(a.node_base + cb, a.prim_base + pb, (uint32_t)meta, (uint32_t)(meta >> 32));
  out[2] = make_uint4((uint32_t)qlo[0], (uint32_t)(qlo[0] >> 32), (uint32_t)qlo[1], (uint32_t)(qlo[1] >> 32));
  out[3] = make_uint4((uint32_t)qlo[2], (uint32_t)(qlo[2] >> 32), (uint32_t)qhi[0], (uint32_t)(qhi[0] >> 32));
  out[4] = make_uint4((uint32_t)qhi[1], (uint32_t)(qhi[1] >> 32), (uint32_t)qhi[2], (uint32_t)(qhi[2] >> 32));
}

// Two builds of the same kernel: MIN_BLOCKS = 2 keeps the unrolled per-child state in registers and is faster up to a
// few million primitives, where a level's barrier waits for the slowest thread; MIN_BLOCKS = 4 (64 registers, twice the
// resident warps) is for the builds that are throughput-bound.
#ifndef ASUNA_EMIT_BIG_BLOCKS
#define ASUNA_EMIT_BIG_BLOCKS 4
#endif
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kThreads, MIN_BLOCKS) k_emit_wide(const EmitParams a) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  PROF(27, 0);
  if (gtid == 0) {
    a.root_of[0] = 0;  // binary root (for n == 1 the single leaf is node n-1 = 0 as well)
    a.counters[0] = 1;
    a.counters[1] = 0;
    a.counters[2] = a.counters[3] = a.counters[4] = 0;  // where the level after this one ends, three levels in rotation
  }
  grid.sync();
  PROF(8, 0);
  uint32_t begin = 0, end = 1, level = 0;
  while (begin < end) {
    // one barrier per level: the nodes of this level publish the end of the next one through an atomicMax of their own
    // slot, which nobody resets before everyone has read it (the slot two levels ahead is cleared here)
    uint32_t* next_end = a.counters + 2 + (level + 1) % 3;
    if (gtid == 0) a.counters[2 + (level + 2) % 3] = 0;
    for (uint32_t w0 = begin + (gtid & ~31u); w0 < end; w0 += gsize) {  // whole warps enter together
      const uint32_t w = w0 + (threadIdx.x & 31u);
      emit_wide_node(a, w, w < end, next_end);
    }
    grid.sync();
    begin = end;
    end = max(end, *(volatile uint32_t*)next_end);
    level++;
    PROF(9, begin);
  }
  // primitive payload: one streaming pass over the slots (sorted position -> primitive id -> triangle / instance)
  const uint32_t n_slots = *(volatile uint32_t*)&a.counters[1];
  for (uint32_t sl = gtid; sl < n_slots; sl += gsize) emit_prim(a, a.prim_base + sl, a.order[a.slot_pos[sl]]);
  PROF(10, n_slots);
  if (gtid == 0) {
    float4 lo = a.nlo[0], hi = a.nhi[0];
    if (a.out_lo) *a.out_lo = lo, *a.out_hi = hi;
    if (a.result) {
      float area = half_area(lo.x, hi.x, lo.y, hi.y, lo.z, hi.z);
      a.result->wide_nodes = end;
      a.result->prim_slots = n_slots;
      a.result->sah_cost = area > 0.f ? a.cost[0] / area : 0.f;
      a.result->pad = 0;
    }
  }
  PROF(28, 0);
}

inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

}  // namespace

// ------------------------------------------------------------------------------------------------
BuildScratch::~BuildScratch() { release(); }
void BuildScratch::release() {
  if (arena) cudaFree(arena);
  arena = nullptr;
  blo = bhi = nlo = nhi = nullptr;
  keys[0] = keys[1] = nullptr;
  vals[0] = vals[1] = nullptr;
  hist = counters = nullptr;
  children = nullptr;
  cost = nullptr;
  dec = nullptr;
  cid[0] = cid[1] = nn = root_of = nullptr;
  clo[0] = clo[1] = chi[0] = chi[1] = nullptr;
  pre = block_sums = nullptr;
  bounds = nullptr;
  capacity = 0;
}
cudaError_t BuildScratch::reserve(uint32_t n) {
  cudaError_t e;
  if (coop_blocks == 0) {
    int dev = 0, sms = 0, per_sm_a = 0, per_sm_b = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, k_ploc, kThreads, 0)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_ploc_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmemBytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScatterSmemBytes)) != cudaSuccess) return e;
    int per_sm_c = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, k_emit_wide<ASUNA_EMIT_BIG_BLOCKS>, kThreads, 0)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, k_emit_wide<2>, kThreads, 0)) != cudaSuccess) return e;
    // each cooperative kernel gets the largest co-resident grid IT fits: both walk trees through dependent, scattered
    // loads, so resident warps are what hides their latency (ASUNA_BUILD_BLOCKS_PER_SM caps both, for experiments)
    int cap_blocks = 4;  // measured on B200: 3 / 4 / 6 / 8 blocks per SM -> 26.9 / 25.5 / 25.9 / 25.8 ms at 32.8 M triangles, 2.15 / 2.11 / 2.18 / 2.18 ms at 1.31 M
    if (const char* t = getenv("ASUNA_BUILD_BLOCKS_PER_SM")) cap_blocks = std::max(1, atoi(t));
    coop_blocks = (uint32_t)(sms * std::max(1, std::min(per_sm_a, cap_blocks)));
    coop_blocks_emit = (uint32_t)(sms * std::max(1, std::min(per_sm_b, cap_blocks)));
    coop_blocks_emit_small = (uint32_t)(sms * std::max(1, std::min(per_sm_c, cap_blocks)));
  }
  if (n <= capacity) return cudaSuccess;
  release();
  uint32_t cap = std::max<uint32_t>(n, 1024);
  // two passes over the same list: sizes first, then one cudaMalloc and the pointers into it
  char* base = nullptr;
  for (int pass = 0; pass < 2; pass++) {
    size_t off = 0;
#define A(ptr, bytes)                                                                   \
  {                                                                                     \
    if (pass) ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(base + off);                       \
    off += (((size_t)(bytes)) + 255) & ~(size_t)255;                                    \
  }
  A(blo, sizeof(float4) * cap);
  A(bhi, sizeof(float4) * cap);
  A(keys[0], sizeof(uint64_t) * cap);
  A(keys[1], sizeof(uint64_t) * cap);
  A(vals[0], sizeof(uint32_t) * cap);
  A(vals[1], sizeof(uint32_t) * cap);
  A(hist, sizeof(uint32_t) * 256 * ((size_t)div_up(cap, kSortTile) + 1));  // [256][blocks] + the 256 digit totals
  A(children, sizeof(int2) * cap);
  A(nlo, sizeof(float4) * 2 * (size_t)cap);
  A(nhi, sizeof(float4) * 2 * (size_t)cap);
  A(cost, sizeof(float) * 7 * 2 * (size_t)cap);
  A(dec, sizeof(uint64_t) * 2 * (size_t)cap);
  for (int k = 0; k < 2; k++) {
    A(cid[k], sizeof(int) * cap);
    A(clo[k], sizeof(float4) * cap);
    A(chi[k], sizeof(float4) * cap);
  }
  A(nn, sizeof(int) * cap);
  A(pre, sizeof(uint2) * cap);
  A(block_sums, sizeof(uint2) * 4096);
  A(root_of, sizeof(int) * cap);
  A(counters, sizeof(uint32_t) * 8);
  A(bounds, sizeof(int) * 8);
    if (!pass) {
      if ((e = cudaMalloc((void**)&base, off)) != cudaSuccess) return e;
      arena = base;
    }
  }
#undef A
  capacity = cap;
  return cudaSuccess;
}

void launch_tri_boxes(cudaStream_t s, const AsunaVertex* v, const uint32_t* idx, uint32_t n, BuildScratch& sc) {
  k_bounds_init<<<1, 32, 0, s>>>(sc.bounds);
  k_tri_boxes<<<div_up(n, kThreads), kThreads, 0, s>>>(v, idx, n, sc.blo, sc.bhi, sc.bounds);
}

void launch_world_triangles_batched(cudaStream_t s, const WorldJob* d_jobs, uint32_t n_jobs, uint32_t max_n, TriSlot* soup,
                                    BuildScratch& sc) {
  k_bounds_init<<<1, 32, 0, s>>>(sc.bounds);
  for (uint32_t j0 = 0; j0 < n_jobs; j0 += 65535u) {  // gridDim.y limit
    const uint32_t nj = std::min(65535u, n_jobs - j0);
    k_world_triangles_batched<<<dim3(div_up(max_n, kThreads), nj), kThreads, 0, s>>>(d_jobs + j0, soup, sc.blo, sc.bhi, sc.bounds);
  }
}

void launch_instance_boxes(cudaStream_t s, const DInstance* inst, const uint32_t* ids, const float4* mesh_lo,
                           const float4* mesh_hi, uint32_t n, BuildScratch& sc) {
  k_bounds_init<<<1, 32, 0, s>>>(sc.bounds);
  k_instance_boxes<<<div_up(n, kThreads), kThreads, 0, s>>>(inst, ids, mesh_lo, mesh_hi, n, sc.blo, sc.bhi, sc.bounds);
}

// LSD radix sort of the 64-bit keys from bit `first_shift` (a multiple of 8) up; returns the buffer the result is in.
// The builder sorts Morton keys of fewer than 16 M primitives on their top 47 bits only (15.7 bits per axis: a 1 / 52 000
// grid, two orders of magnitude finer than the primitive spacing such a scene can have; primitives sharing a cell keep
// their input order and PLOC's neighbour search is insensitive to it -- the SAH-quality test guards this): 6 passes
// instead of 8; larger scenes on their top 55 bits (18.3 per axis): 7 passes.
static int sort_passes(cudaStream_t s, uint32_t n, BuildScratch& sc, int first_shift = 0) {
  uint32_t sort_blocks = div_up(n, kSortTile);
  int cur = 0;
  for (int shift = first_shift; shift < 64; shift += 8) {
    k_sort_hist<<<sort_blocks, kThreads, 0, s>>>(sc.keys[cur], n, shift, sc.hist, sort_blocks);
    uint32_t* totals = sc.hist + 256u * (size_t)sort_blocks;
    k_sort_scan_rows<<<256, kThreads, 0, s>>>(sc.hist, sort_blocks, totals);
    k_sort_scatter<<<sort_blocks, kThreads, kScatterSmemBytes, s>>>(sc.keys[cur], sc.vals[cur], sc.keys[cur ^ 1], sc.vals[cur ^ 1],
                                                                    n, shift, sc.hist, sort_blocks, totals);
    cur ^= 1;
  }
  return cur;  // the buffer that holds the sorted pairs
}

// Builds the wide BVH over the n boxes already in sc.blo/bhi (bounds in sc.bounds): nodes[node_base ..) receive the
// wide nodes (root first), primitive slots prim_base.. are filled through `payload` (triangles or instance ids).
cudaError_t launch_build_wide(cudaStream_t s, uint32_t n, WideNode* nodes, uint32_t node_base, uint32_t prim_base,
                              BuildScratch& sc, const PrimPayload& payload, float cost_prim, float4* root_lo,
                              float4* root_hi, BuildResult* result) {
  int sorted = 0;
  if (n >= 2) {
    k_morton<<<div_up(n, kThreads), kThreads, 0, s>>>(sc.blo, sc.bhi, n, sc.bounds, sc.keys[0], sc.vals[0]);
    sorted = sort_passes(s, n, sc, n < (1u << 24) ? 16 : 8);
  } else {
    cudaMemsetAsync(sc.vals[0], 0, sizeof(uint32_t), s);
  }
  uint32_t grid = std::max(1u, std::min(sc.coop_blocks, div_up(n, kThreads)));
  PlocParams pp;
  pp.n = (int)n;
  pp.cost_node = 1.0f;
  pp.cost_prim = cost_prim;
  pp.blo = sc.blo, pp.bhi = sc.bhi, pp.order = sc.vals[sorted];
  pp.nlo = sc.nlo, pp.nhi = sc.nhi, pp.children = sc.children, pp.cost = sc.cost, pp.dec = sc.dec;
  for (int k = 0; k < 2; k++) pp.cid[k] = sc.cid[k], pp.clo[k] = sc.clo[k], pp.chi[k] = sc.chi[k];
  pp.nn = sc.nn, pp.pre = sc.pre, pp.block_sums = sc.block_sums;
  pp.state = reinterpret_cast<int*>(sc.counters + 5);
  cudaError_t e;
  if (n > (uint32_t)kPlocTail) {
    void* pargs[] = {&pp};
    if ((e = cudaLaunchCooperativeKernel((const void*)k_ploc, dim3(grid), dim3(kThreads), pargs, 0, s)) != cudaSuccess) return e;
  }
  k_ploc_tail<<<1, kTailThreads, kTailSmemBytes, s>>>(pp, n <= (uint32_t)kPlocTail ? 1 : 0);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  EmitParams ep;
  ep.n = (int)n;
  ep.children = sc.children, ep.nlo = sc.nlo, ep.nhi = sc.nhi, ep.dec = sc.dec, ep.cost = sc.cost;
  ep.order = sc.vals[sorted];
  ep.nodes = nodes, ep.node_base = node_base, ep.prim_base = prim_base;
  ep.root_of = sc.root_of, ep.counters = sc.counters;
  ep.slot_pos = reinterpret_cast<uint32_t*>(sc.nn);  // PLOC's scratch is free by now
  ep.v = payload.vertices, ep.idx = payload.indices, ep.tris = payload.tris, ep.leaf_inst = payload.leaf_inst;
  ep.soup = payload.soup, ep.prim_ids = payload.prim_ids;
  ep.out_lo = root_lo, ep.out_hi = root_hi, ep.result = result;
  void* eargs[] = {&ep};
  const bool small = n <= (4u << 20);
  const uint32_t grid_emit = std::max(1u, std::min(small ? sc.coop_blocks_emit_small : sc.coop_blocks_emit, div_up(n, kThreads)));
  return cudaLaunchCooperativeKernel(small ? (const void*)k_emit_wide<2> : (const void*)k_emit_wide<ASUNA_EMIT_BIG_BLOCKS>, dim3(grid_emit),
                                     dim3(kThreads), eargs, 0, s);
}

// Host-side debug/test hook: sorts (key,value) pairs with the builder's radix sort.
cudaError_t radix_sort_pairs(cudaStream_t s, uint64_t* keys_io, uint32_t* vals_io, uint32_t n, BuildScratch& sc) {
  cudaError_t e = sc.reserve(n);
  if (e != cudaSuccess) return e;
  cudaMemcpyAsync(sc.keys[0], keys_io, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(sc.vals[0], vals_io, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s);
  const int sorted = sort_passes(s, n, sc);
  cudaMemcpyAsync(keys_io, sc.keys[sorted], sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(vals_io, sc.vals[sorted], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s);
  return cudaGetLastError();
}

}  // namespace asuna

#ifdef ASUNA_BUILD_PROFILE
extern "C" int asuna_debug_build_profile(unsigned long long* out, uint32_t cap) {
  uint32_t n = 0, zero = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, asuna::g_build_prof_n, sizeof(n));
  n = n < 2048u ? n : 2048u;
  n = n < cap ? n : cap;
  cudaMemcpyFromSymbol(out, asuna::g_build_prof, (size_t)n * 16);
  cudaMemcpyToSymbol(asuna::g_build_prof_n, &zero, sizeof(zero));
  return (int)n;
}
#endif
