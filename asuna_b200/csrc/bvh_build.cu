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
constexpr int kSortItems = 16;
constexpr int kSortTile = kThreads * kSortItems;  // 4096 keys per block

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

__device__ __forceinline__ void reduce_bounds(float3 lo, float3 hi, bool valid, int* bounds) {
  // centroid bounds: warp shuffle reduce, one atomic per warp
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
  // One atomic per warp only where it would still move the bound: six same-line atomics from every one of the million
  // warps of a 32 M-triangle build serialise in one L2 slice (measured: 4.6 ms of a 28 ms build); the plain read is cheap
  // and after the first few thousand warps almost nothing improves the bounds any more.
  if ((threadIdx.x & 31) == 0) {
    const volatile int* b = bounds;
    const int o0 = float_to_ordered(mnx), o1 = float_to_ordered(mny), o2 = float_to_ordered(mnz);
    const int o3 = float_to_ordered(mxx), o4 = float_to_ordered(mxy), o5 = float_to_ordered(mxz);
    if (o0 < b[0]) atomicMin(&bounds[0], o0);
    if (o1 < b[1]) atomicMin(&bounds[1], o1);
    if (o2 < b[2]) atomicMin(&bounds[2], o2);
    if (o3 > b[3]) atomicMax(&bounds[3], o3);
    if (o4 > b[4]) atomicMax(&bounds[4], o4);
    if (o5 > b[5]) atomicMax(&bounds[5], o5);
  }
}

__global__ void k_tri_boxes(const AsunaVertex* __restrict__ v, const uint32_t* __restrict__ idx, uint32_t n,
                            float4* __restrict__ blo, float4* __restrict__ bhi, int* bounds) {
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
  reduce_bounds(lo, hi, valid, bounds);
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
  if (blockIdx.x * kThreads >= job.n) return;  // whole block idle: no barrier or ballot is skipped by part of a warp
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
  reduce_bounds(lo, hi, valid, bounds);
}

// World box of an instance = box of the 8 transformed corners of its mesh box (what a TLAS build sees).
__global__ void k_instance_boxes(const DInstance* __restrict__ inst, const uint32_t* __restrict__ ids,
                                 const float4* __restrict__ mesh_lo, const float4* __restrict__ mesh_hi, uint32_t n,
                                 float4* __restrict__ blo, float4* __restrict__ bhi, int* bounds) {
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
  reduce_bounds(lo, hi, valid, bounds);
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
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint32_t base = blockIdx.x * kSortTile;
#pragma unroll 4
  for (int k = 0; k < kSortItems; k++) {
    uint32_t i = base + k * kThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// One block per digit: exclusive scan of that digit's per-block counts in place (row d of hist), row total to
// totals[d].  The digit bases (exclusive scan of the 256 totals) are recomputed by every scatter block, so the pass
// needs no serial walk over the whole 256 x blocks table.
__global__ void __launch_bounds__(kThreads) k_sort_scan_rows(uint32_t* __restrict__ hist, uint32_t n_blocks,
                                                             uint32_t* __restrict__ totals) {
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

__global__ void __launch_bounds__(kThreads) k_sort_scatter(const uint64_t* __restrict__ keys_in,
                                                            const uint32_t* __restrict__ vals_in,
                                                            uint64_t* __restrict__ keys_out,
                                                            uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                            const uint32_t* __restrict__ hist, uint32_t n_blocks,
                                                            const uint32_t* __restrict__ totals) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kRounds = kSortTile / kWarps / 32;  // 16
  __shared__ uint32_t wh[kWarps][256];
  __shared__ uint32_t digit_sums[kWarps];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    if (lane == 31) digit_sums[warp] = sc;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) before += w < warp ? digit_sums[w] : 0u;
    digit_base = before + sc - t;
  }
  for (int b = threadIdx.x; b < kWarps * 256; b += kThreads) (&wh[0][0])[b] = 0;
  __syncthreads();
  uint32_t wbase = blockIdx.x * kSortTile + warp * (kRounds * 32);
  // pass 1: per-warp digit counts (one leader lane per distinct digit per round)
  for (int r = 0; r < kRounds; r++) {
    uint32_t i = wbase + r * 32 + lane;
    bool valid = i < n;
    uint32_t digit = valid ? ((uint32_t)(keys_in[i] >> shift) & 255u) : (256u + lane);
    uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
    if (valid && lane == __ffs(peers) - 1) wh[warp][digit] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // per digit: global base of this block, then exclusive over the warps of the block
  {
    uint32_t bin = threadIdx.x;
    uint32_t run = digit_base + hist[bin * n_blocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      uint32_t c = wh[w][bin];
      wh[w][bin] = run;
      run += c;
    }
  }
  __syncthreads();
  // pass 2: rank within the warp round and scatter
  for (int r = 0; r < kRounds; r++) {
    uint32_t i = wbase + r * 32 + lane;
    bool valid = i < n;
    uint64_t key = valid ? keys_in[i] : 0ull;
    uint32_t digit = valid ? ((uint32_t)(key >> shift) & 255u) : (256u + lane);
    uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (valid) {
      pos = wh[warp][digit] + rank;
      keys_out[pos] = key;
      vals_out[pos] = vals_in[i];
    }
    __syncwarp();
    if (valid && lane == __ffs(peers) - 1) wh[warp][digit] += __popc(peers);
    __syncwarp();
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
#define ASUNA_PLOC_TAIL 4096
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
__device__ void dp_merge(const PlocParams& a, int idx, int l, int r, float area, uint32_t count) {
  float cl[7], cr[7], c[7];
#pragma unroll
  for (int i = 0; i < 7; i++) cl[i] = a.cost[(size_t)l * 7 + i], cr[i] = a.cost[(size_t)r * 7 + i];
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

__global__ void __launch_bounds__(kThreads, ASUNA_PLOC_MIN_BLOCKS) k_ploc(const PlocParams a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ uint32_t warp_sums[kThreads / 32];
  __shared__ uint32_t red[4];
  const int n = a.n;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  for (uint32_t j = gtid; j < (uint32_t)n; j += gsize) {
    uint32_t prim = a.order[j];
    float4 lo = a.blo[prim], hi = a.bhi[prim];
    lo.w = __uint_as_float(1u);
    hi.w = 0.f;
    int node = n - 1 + (int)j;
    a.nlo[node] = lo;
    a.nhi[node] = hi;
    float c = half_area(lo.x, hi.x, lo.y, hi.y, lo.z, hi.z) * a.cost_prim;
#pragma unroll
    for (int i = 0; i < 7; i++) a.cost[(size_t)node * 7 + i] = c;
    a.dec[node] = kDecLeaf;
    a.cid[0][j] = node;
    a.clo[0][j] = lo;
    a.chi[0][j] = hi;
  }
  grid.sync();
  int m = n, cur = 0, next_inner = n - 2;
  // Tail: once few clusters are left an iteration is pure synchronisation latency (3 grid-wide barriers for a handful of
  // merges, and half of all iterations happen below a few thousand clusters), so block 0 finishes alone behind
  // __syncthreads and the other blocks leave.  nb / bid / the strides describe whichever grid is still running.
  bool tail = false;
  uint32_t nb = gridDim.x, bid = blockIdx.x, tid = gtid, tsize = gsize;
  while (m > 1) {
    if (!tail && m <= kPlocTail) {
      if (blockIdx.x != 0) return;
      tail = true;
      nb = 1, bid = 0, tid = threadIdx.x, tsize = kThreads;
    }
    const int* cid = a.cid[cur];
    const float4* clo = a.clo[cur];
    const float4* chi = a.chi[cur];
    // phase 1: nearest neighbour inside the radius (ties -> lowest index, so a mutual pair always exists)
    for (uint32_t i = tid; i < (uint32_t)m; i += tsize) {
      float4 lo = clo[i], hi = chi[i];
      int j0 = max(0, (int)i - kPlocRadius), j1 = min(m - 1, (int)i + kPlocRadius);
      float best = FLT_MAX;
      int bj = (int)i == j0 ? j0 + 1 : j0;
      for (int j = j0; j <= j1; j++) {
        if (j == (int)i) continue;
        float d = union_half_area(lo, hi, clo[j], chi[j]);
        if (d < best) best = d, bj = j;
      }
      a.nn[i] = bj;
    }
    if (tail) __syncthreads(); else grid.sync();
    // phase 2: survivor / merge flags, prefix inside this block's contiguous chunk (keeps Morton order)
    const int chunk = (m + (int)nb - 1) / (int)nb;
    const int c0 = min(m, (int)bid * chunk), c1 = min(m, c0 + chunk);
    uint32_t run_valid = 0, run_lead = 0;
    for (int base = c0; base < c1; base += kThreads) {
      int i = base + (int)threadIdx.x;
      uint32_t packed = 0;
      if (i < c1) {
        int j = a.nn[i];
        bool mutual = a.nn[j] == i;
        packed = ((!mutual || i < j) ? 1u : 0u) | ((mutual && i < j) ? 0x10000u : 0u);
      }
      uint32_t total;
      uint32_t excl = block_scan_excl(packed, total, warp_sums);
      if (i < c1) a.pre[i] = make_uint2(run_valid + (excl & 0xFFFFu), run_lead + (excl >> 16));
      run_valid += total & 0xFFFFu;
      run_lead += total >> 16;
    }
    if (threadIdx.x == 0) a.block_sums[bid] = make_uint2(run_valid, run_lead);
    if (tail) __syncthreads(); else grid.sync();
    // phase 3: global offsets from the block sums, then merge / copy into the next cluster array
    {
      uint32_t bv = 0, bl = 0, tv = 0, tl = 0;
      for (uint32_t b = threadIdx.x; b < nb; b += kThreads) {
        uint2 s = a.block_sums[b];
        tv += s.x, tl += s.y;
        if (b < bid) bv += s.x, bl += s.y;
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
      int j = a.nn[i];
      bool mutual = a.nn[j] == i;
      if (mutual && i > j) continue;
      uint2 p = a.pre[i];
      uint32_t pos = base_valid + p.x;
      float4 lo = clo[i], hi = chi[i];
      int id = cid[i];
      if (mutual) {
        float4 lo2 = clo[j], hi2 = chi[j];
        int id2 = cid[j];
        uint32_t count = __float_as_uint(lo.w) + __float_as_uint(lo2.w);
        lo = make_float4(fminf(lo.x, lo2.x), fminf(lo.y, lo2.y), fminf(lo.z, lo2.z), __uint_as_float(count));
        hi = make_float4(fmaxf(hi.x, hi2.x), fmaxf(hi.y, hi2.y), fmaxf(hi.z, hi2.z), 0.f);
        int idx = next_inner - (int)(base_lead + p.y);
        a.children[idx] = make_int2(id, id2);
        a.nlo[idx] = lo;
        a.nhi[idx] = hi;
        dp_merge(a, idx, id, id2, half_area(lo.x, hi.x, lo.y, hi.y, lo.z, hi.z), count);
        id = idx;
      }
      ocid[pos] = id;
      oclo[pos] = lo;
      ochi[pos] = hi;
    }
    if (tail) __syncthreads(); else grid.sync();
    m = (int)tot_valid;
    next_inner -= (int)tot_lead;
    cur ^= 1;
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
  uint32_t* counters;      // [0] wide nodes handed out, [1] primitive slots handed out
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

__device__ __forceinline__ uint32_t pack4(const uint8_t* b) {
  return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
}

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

// Called by whole warps (`valid` = this lane has a node): node and primitive-slot ranges are handed out with ONE atomic
// per warp per counter -- a 1.3 M-triangle build emits 200 k wide nodes, and 400 k same-address atomics cost more than
// all the arithmetic of this kernel.
__device__ void emit_wide_node(const EmitParams& a, uint32_t w, bool valid) {
  const int n = a.n;
  const int root = valid ? a.root_of[w] : 0;
  int ch_node[8];
  bool ch_leaf[8];
  int nc = 0;
  if (valid) {
    uint64_t droot = a.dec[root];
    if (root >= n - 1 || (droot & kDecLeaf)) {
      ch_node[0] = root, ch_leaf[0] = true, nc = 1;  // a BVH of <= 3 primitives: one leaf child
    } else {
      int st_node[8], st_bud[8], sp = 0;
      int2 c = a.children[root];
      uint32_t k = (uint32_t)(droot >> (4 + 6 * 6)) & 63u;
      st_node[sp] = c.y, st_bud[sp++] = (int)(k >> 3);
      st_node[sp] = c.x, st_bud[sp++] = (int)(k & 7u);
      while (sp > 0) {
        int x = st_node[--sp], bud = st_bud[sp];
        bool leaf2 = x >= n - 1;
        uint64_t d = a.dec[x];
        uint32_t code = (bud >= 2 && !leaf2) ? (uint32_t)(d >> (4 + 6 * (bud - 2))) & 63u : 0u;
        if (code == 0u) {
          ch_node[nc] = x, ch_leaf[nc] = leaf2 || (d & kDecLeaf), nc++;
        } else {
          int2 cc = a.children[x];
          st_node[sp] = cc.y, st_bud[sp++] = (int)(code >> 3);
          st_node[sp] = cc.x, st_bud[sp++] = (int)(code & 7u);
        }
      }
    }
  }
  float4 rlo = make_float4(0.f, 0.f, 0.f, 0.f), rhi = rlo;
  if (valid) rlo = a.nlo[root], rhi = a.nhi[root];
  float4 clo[8], chi[8];
  for (int i = 0; i < nc; i++) clo[i] = a.nlo[ch_node[i]], chi[i] = a.nhi[ch_node[i]];

  // octant-ordered slots: greedy assignment maximising sum dot(child centre - node centre, dir(slot)),
  // dir(slot) = +1 on the axes whose slot bit is set (x = bit 0), so slot ^ ray-octant is front to back
  int slot_of[8];
  {
    float cx = 0.5f * (rlo.x + rhi.x), cy = 0.5f * (rlo.y + rhi.y), cz = 0.5f * (rlo.z + rhi.z);
    float dx[8], dy[8], dz[8];
    for (int i = 0; i < nc; i++) {
      dx[i] = 0.5f * (clo[i].x + chi[i].x) - cx, dy[i] = 0.5f * (clo[i].y + chi[i].y) - cy;
      dz[i] = 0.5f * (clo[i].z + chi[i].z) - cz;
      slot_of[i] = -1;
    }
    // all loops have constant bounds and are fully unrolled, so the 8 x 8 score table and the bookkeeping live in
    // registers: the dynamically indexed form kept everything in local memory, and this assignment is the longest
    // serial stretch of a thread that a whole level's barrier waits for
    float score[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int sl = 0; sl < 8; sl++)
        score[i][sl] = i < nc ? ((sl & 1) ? dx[i] : -dx[i]) + ((sl & 2) ? dy[i] : -dy[i]) + ((sl & 4) ? dz[i] : -dz[i]) : -FLT_MAX;
    uint32_t slot_used = 0, child_done = 0;
#pragma unroll
    for (int round = 0; round < 8; round++) {
      if (round < nc) {
        float best = -FLT_MAX;
        int bi = 0, bs = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int sl = 0; sl < 8; sl++) {
            const bool free_pair = i < nc && !((child_done >> i) & 1u) && !((slot_used >> sl) & 1u);
            if (free_pair && score[i][sl] > best) best = score[i][sl], bi = i, bs = sl;
          }
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (i == bi) slot_of[i] = bs;
        slot_used |= 1u << bs;
        child_done |= 1u << bi;
      }
    }
  }
  int child_in_slot[8];
  for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
  uint32_t n_inner = 0, n_prims = 0;
  for (int i = 0; i < nc; i++) {
    child_in_slot[slot_of[i]] = i;
    if (ch_leaf[i]) n_prims += __float_as_uint(clo[i].w);
    else n_inner++;
  }
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
      if (tot_i) bi = atomicAdd(&a.counters[0], tot_i);
      if (tot_p) bp = atomicAdd(&a.counters[1], tot_p);
    }
    cb = __shfl_sync(0xFFFFFFFFu, bi, 0) + xi - n_inner;
    pb = __shfl_sync(0xFFFFFFFFu, bp, 0) + xp - n_prims;
  }
  if (!valid) return;

  // quantisation grid: origin = padded lower corner, per-axis power-of-two scale covering the padded extent
  float pad[3], p[3], inv_scale[3];
  uint8_t e[3];
  {
    const float lo3[3] = {rlo.x, rlo.y, rlo.z}, hi3[3] = {rhi.x, rhi.y, rhi.z};
    for (int k = 0; k < 3; k++) {
      pad[k] = fmaxf(fabsf(lo3[k]), fabsf(hi3[k])) * 2.4e-7f + 1e-30f;
      p[k] = lo3[k] - pad[k];
      float f = ((hi3[k] + pad[k]) - p[k]) * (1.0f / 254.0f);
      uint32_t bits = __float_as_uint(f);
      uint32_t eb = (bits >> 23) + ((bits & 0x7FFFFFu) ? 1u : 0u);
      eb = min(max(eb, 1u), 253u);
      e[k] = (uint8_t)eb;
      inv_scale[k] = __uint_as_float((254u - eb) << 23);
    }
  }
  WideNode nd;
  nd.px = p[0], nd.py = p[1], nd.pz = p[2];
  nd.ex = e[0], nd.ey = e[1], nd.ez = e[2];
  nd.child_base = a.node_base + cb;
  nd.prim_base = a.prim_base + pb;
  uint32_t imask = 0, inner_rank = 0, prim_off = 0;
  for (int s = 0; s < 8; s++) {
    int i = child_in_slot[s];
    if (i < 0) {
      nd.meta[s] = 0;
      nd.qlox[s] = nd.qloy[s] = nd.qloz[s] = 255;
      nd.qhix[s] = nd.qhiy[s] = nd.qhiz[s] = 0;
      continue;
    }
    const float lo3[3] = {clo[i].x, clo[i].y, clo[i].z}, hi3[3] = {chi[i].x, chi[i].y, chi[i].z};
    uint8_t ql[3], qh[3];
    for (int k = 0; k < 3; k++) {
      float fl = floorf((lo3[k] - pad[k] - p[k]) * inv_scale[k]);
      float fh = ceilf((hi3[k] + pad[k] - p[k]) * inv_scale[k]);
      ql[k] = (uint8_t)fminf(fmaxf(fl, 0.f), 255.f);
      qh[k] = (uint8_t)fminf(fmaxf(fh, 0.f), 255.f);
    }
    nd.qlox[s] = ql[0], nd.qloy[s] = ql[1], nd.qloz[s] = ql[2];
    nd.qhix[s] = qh[0], nd.qhiy[s] = qh[1], nd.qhiz[s] = qh[2];
    if (!ch_leaf[i]) {
      imask |= 1u << s;
      nd.meta[s] = (uint8_t)(0x20u | (24u + (uint32_t)s));
      a.root_of[cb + inner_rank] = ch_node[i];
      inner_rank++;
    } else {
      uint32_t cnt = __float_as_uint(clo[i].w);
      nd.meta[s] = (uint8_t)((((1u << cnt) - 1u) << 5) | prim_off);
      // the <= 3 primitives of the subtree, in sorted order
      int st[4], sp = 0;
      st[sp++] = ch_node[i];
      uint32_t k = 0;
      while (sp > 0) {
        int x = st[--sp];
        if (x >= n - 1) {
          emit_prim(a, a.prim_base + pb + prim_off + k, a.order[x - (n - 1)]);
          k++;
        } else {
          int2 cc = a.children[x];
          st[sp++] = cc.y;
          st[sp++] = cc.x;
        }
      }
      prim_off += cnt;
    }
  }
  nd.imask = (uint8_t)imask;
  // five 16-byte stores
  uint4* out = reinterpret_cast<uint4*>(a.nodes + a.node_base + w);
  const uint8_t eb[4] = {nd.ex, nd.ey, nd.ez, nd.imask};
  out[0] = make_uint4(__float_as_uint(nd.px), __float_as_uint(nd.py), __float_as_uint(nd.pz), pack4(eb));
  out[1] = make_uint4(nd.child_base, nd.prim_base, pack4(nd.meta), pack4(nd.meta + 4));
  out[2] = make_uint4(pack4(nd.qlox), pack4(nd.qlox + 4), pack4(nd.qloy), pack4(nd.qloy + 4));
  out[3] = make_uint4(pack4(nd.qloz), pack4(nd.qloz + 4), pack4(nd.qhix), pack4(nd.qhix + 4));
  out[4] = make_uint4(pack4(nd.qhiy), pack4(nd.qhiy + 4), pack4(nd.qhiz), pack4(nd.qhiz + 4));
}

// Two builds of the same kernel: MIN_BLOCKS = 2 keeps the unrolled slot assignment in registers (104 of them) and is
// 6-9 % faster up to a few million primitives, where a level's barrier waits for the slowest thread; MIN_BLOCKS = 4
// (64 registers, twice the resident warps) wins on the 32.8 M-triangle build, which is throughput-bound (25.4 vs 26.2 ms).
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kThreads, MIN_BLOCKS) k_emit_wide(const EmitParams a) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  if (gtid == 0) {
    a.root_of[0] = 0;  // binary root (for n == 1 the single leaf is node n-1 = 0 as well)
    a.counters[0] = 1;
    a.counters[1] = 0;
  }
  grid.sync();
  uint32_t begin = 0, end = 1;
  while (begin < end) {
    for (uint32_t w0 = begin + (gtid & ~31u); w0 < end; w0 += gsize) {  // whole warps enter together
      const uint32_t w = w0 + (threadIdx.x & 31u);
      emit_wide_node(a, w, w < end);
    }
    grid.sync();
    begin = end;
    end = *(volatile uint32_t*)&a.counters[0];
    grid.sync();  // everyone has read the level boundary before the next level allocates
  }
  if (gtid == 0) {
    float4 lo = a.nlo[0], hi = a.nhi[0];
    if (a.out_lo) *a.out_lo = lo, *a.out_hi = hi;
    if (a.result) {
      float area = half_area(lo.x, hi.x, lo.y, hi.y, lo.z, hi.z);
      a.result->wide_nodes = end;
      a.result->prim_slots = a.counters[1];
      a.result->sah_cost = area > 0.f ? a.cost[0] / area : 0.f;
      a.result->pad = 0;
    }
  }
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
    int per_sm_c = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, k_emit_wide<4>, kThreads, 0)) != cudaSuccess) return e;
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
  A(counters, sizeof(uint32_t) * 4);
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

// LSD radix sort of the 64-bit keys from bit `first_shift` up (a multiple of 16, so the pass count stays even and the
// result lands in buffer 0).  The builder sorts Morton keys of fewer than 16 M primitives on their top 47 bits only
// (15.7 bits per axis: a 1 / 52 000 grid, two orders of magnitude finer than the primitive spacing such a scene can have;
// primitives sharing a cell keep their input order and PLOC's neighbour search is insensitive to it -- the SAH-quality
// test guards this): 6 passes instead of 8.
static void sort_passes(cudaStream_t s, uint32_t n, BuildScratch& sc, int first_shift = 0) {
  uint32_t sort_blocks = div_up(n, kSortTile);
  int cur = 0;
  for (int shift = first_shift; shift < 64; shift += 8) {
    k_sort_hist<<<sort_blocks, kThreads, 0, s>>>(sc.keys[cur], n, shift, sc.hist, sort_blocks);
    uint32_t* totals = sc.hist + 256u * (size_t)sort_blocks;
    k_sort_scan_rows<<<256, kThreads, 0, s>>>(sc.hist, sort_blocks, totals);
    k_sort_scatter<<<sort_blocks, kThreads, 0, s>>>(sc.keys[cur], sc.vals[cur], sc.keys[cur ^ 1], sc.vals[cur ^ 1],
                                                    n, shift, sc.hist, sort_blocks, totals);
    cur ^= 1;
  }
  // an even number of passes: the result is back in buffer 0
}

// Builds the wide BVH over the n boxes already in sc.blo/bhi (bounds in sc.bounds): nodes[node_base ..) receive the
// wide nodes (root first), primitive slots prim_base.. are filled through `payload` (triangles or instance ids).
cudaError_t launch_build_wide(cudaStream_t s, uint32_t n, WideNode* nodes, uint32_t node_base, uint32_t prim_base,
                              BuildScratch& sc, const PrimPayload& payload, float cost_prim, float4* root_lo,
                              float4* root_hi, BuildResult* result) {
  if (n >= 2) {
    k_morton<<<div_up(n, kThreads), kThreads, 0, s>>>(sc.blo, sc.bhi, n, sc.bounds, sc.keys[0], sc.vals[0]);
    sort_passes(s, n, sc, n < (1u << 24) ? 16 : 0);
  } else {
    cudaMemsetAsync(sc.vals[0], 0, sizeof(uint32_t), s);
  }
  uint32_t grid = std::max(1u, std::min(sc.coop_blocks, div_up(n, kThreads)));
  PlocParams pp;
  pp.n = (int)n;
  pp.cost_node = 1.0f;
  pp.cost_prim = cost_prim;
  pp.blo = sc.blo, pp.bhi = sc.bhi, pp.order = sc.vals[0];
  pp.nlo = sc.nlo, pp.nhi = sc.nhi, pp.children = sc.children, pp.cost = sc.cost, pp.dec = sc.dec;
  for (int k = 0; k < 2; k++) pp.cid[k] = sc.cid[k], pp.clo[k] = sc.clo[k], pp.chi[k] = sc.chi[k];
  pp.nn = sc.nn, pp.pre = sc.pre, pp.block_sums = sc.block_sums;
  void* pargs[] = {&pp};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_ploc, dim3(grid), dim3(kThreads), pargs, 0, s);
  if (e != cudaSuccess) return e;
  EmitParams ep;
  ep.n = (int)n;
  ep.children = sc.children, ep.nlo = sc.nlo, ep.nhi = sc.nhi, ep.dec = sc.dec, ep.cost = sc.cost;
  ep.order = sc.vals[0];
  ep.nodes = nodes, ep.node_base = node_base, ep.prim_base = prim_base;
  ep.root_of = sc.root_of, ep.counters = sc.counters;
  ep.v = payload.vertices, ep.idx = payload.indices, ep.tris = payload.tris, ep.leaf_inst = payload.leaf_inst;
  ep.soup = payload.soup, ep.prim_ids = payload.prim_ids;
  ep.out_lo = root_lo, ep.out_hi = root_hi, ep.result = result;
  void* eargs[] = {&ep};
  const bool small = n <= (4u << 20);
  const uint32_t grid_emit = std::max(1u, std::min(small ? sc.coop_blocks_emit_small : sc.coop_blocks_emit, div_up(n, kThreads)));
  return cudaLaunchCooperativeKernel(small ? (const void*)k_emit_wide<2> : (const void*)k_emit_wide<4>, dim3(grid_emit),
                                     dim3(kThreads), eargs, 0, s);
}

// Host-side debug/test hook: sorts (key,value) pairs with the builder's radix sort.
cudaError_t radix_sort_pairs(cudaStream_t s, uint64_t* keys_io, uint32_t* vals_io, uint32_t n, BuildScratch& sc) {
  cudaError_t e = sc.reserve(n);
  if (e != cudaSuccess) return e;
  cudaMemcpyAsync(sc.keys[0], keys_io, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(sc.vals[0], vals_io, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s);
  sort_passes(s, n, sc);
  cudaMemcpyAsync(keys_io, sc.keys[0], sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(vals_io, sc.vals[0], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s);
  return cudaGetLastError();
}

}  // namespace asuna
