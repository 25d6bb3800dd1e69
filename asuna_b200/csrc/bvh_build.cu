// Acceleration-structure builder, hand-written for sm_100a.
//
// Replaces the driver build behind vkCmdBuildAccelerationStructuresKHR that the reference reaches
// through nvvk::RaytracingBuilderKHR::buildBlas / buildTlas (reference
// ext/nvpro_core/nvvk/raytraceKHR_vk.cpp:77-183,302-376, called from
// src/pipeline/pipeline_raytrace.cpp:107-147).  Pipeline per BVH (all on one stream, no host sync):
//   prim boxes + bounds reduce -> 63-bit Morton keys -> LSD radix sort (8 x 8-bit, stable,
//   histogram / scan / warp-multisplit scatter) -> Karras 2012 LBVH hierarchy -> bottom-up box
//   fit (atomic arrival flags) -> emit 64-byte child-box nodes + leaf-ordered triangle slots.
// Everything is HBM-bound streaming work: loads are coalesced 16-byte where the input layout
// allows (the 44-byte vertex stride of the wire format does not), grids are sized from n.
#include <algorithm>
#include <cfloat>

#include "bvh_build.cuh"

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
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&bounds[0], float_to_ordered(mnx));
    atomicMin(&bounds[1], float_to_ordered(mny));
    atomicMin(&bounds[2], float_to_ordered(mnz));
    atomicMax(&bounds[3], float_to_ordered(mxx));
    atomicMax(&bounds[4], float_to_ordered(mxy));
    atomicMax(&bounds[5], float_to_ordered(mxz));
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

// World box of an instance = box of the 8 transformed corners of its mesh box (what a TLAS build sees).
__global__ void k_instance_boxes(const DInstance* __restrict__ inst, const float4* __restrict__ mesh_lo,
                                 const float4* __restrict__ mesh_hi, uint32_t n, float4* __restrict__ blo,
                                 float4* __restrict__ bhi, int* bounds) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = i < n;
  float3 lo = make_float3(0, 0, 0), hi = lo;
  if (valid) {
    const DInstance& in = inst[i];
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

// Exclusive scan of `total` counters in place, one 1024-thread block walking the array.
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t* __restrict__ hist, uint32_t total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < total; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < total ? hist[i] : 0u;
    uint32_t s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane], ws = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, ws, o);
        if (lane >= o) ws += t;
      }
      warp_sums[lane] = ws - w;  // exclusive
    }
    __syncthreads();
    uint32_t c = carry;
    uint32_t excl = c + warp_sums[warp] + s - v;
    if (i < total) hist[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) k_sort_scatter(const uint64_t* __restrict__ keys_in,
                                                            const uint32_t* __restrict__ vals_in,
                                                            uint64_t* __restrict__ keys_out,
                                                            uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                            const uint32_t* __restrict__ hist, uint32_t n_blocks) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kRounds = kSortTile / kWarps / 32;  // 16
  __shared__ uint32_t wh[kWarps][256];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    uint32_t run = hist[bin * n_blocks + blockIdx.x];
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

// ---- 4. LBVH hierarchy (Karras 2012) ---------------------------------------------------------
__device__ __forceinline__ int key_delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}

// Node numbering during the build: inner nodes 0..n-2 (root 0), leaf j is n-1+j.
__global__ void k_lbvh_hierarchy(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children,
                                 int* __restrict__ parent) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = (key_delta(keys, n, i, i + 1) - key_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = key_delta(keys, n, i, i - d);
  int lmax = 2;
  while (key_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (key_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = key_delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (key_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  int gamma = i + s * d + min(d, 0);
  int left = (min(i, j) == gamma) ? (n - 1 + gamma) : gamma;
  int right = (max(i, j) == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
  children[i] = make_int2(left, right);
  parent[left] = i;
  parent[right] = i;
  if (i == 0) parent[0] = -1;
}

// ---- 5. bottom-up fit: second arrival at a node merges its children ---------------------------
__global__ void k_lbvh_fit(const float4* __restrict__ blo, const float4* __restrict__ bhi,
                           const uint32_t* __restrict__ order, int n, const int2* __restrict__ children,
                           const int* __restrict__ parent, float4* nlo, float4* nhi, uint32_t* flags) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t prim = order[j];
  float4 lo = blo[prim], hi = bhi[prim];
  int node = n - 1 + j;
  nlo[node] = lo;
  nhi[node] = hi;
  __threadfence();
  int p = parent[node];
  while (p >= 0) {
    if (atomicAdd(&flags[p], 1u) == 0u) return;  // first arrival: the sibling will carry on
    __threadfence();
    int2 c = children[p];
    volatile float4* vlo = nlo;
    volatile float4* vhi = nhi;
    float ax = vlo[c.x].x, ay = vlo[c.x].y, az = vlo[c.x].z, bx = vlo[c.y].x, by = vlo[c.y].y, bz = vlo[c.y].z;
    float Ax = vhi[c.x].x, Ay = vhi[c.x].y, Az = vhi[c.x].z, Bx = vhi[c.y].x, By = vhi[c.y].y, Bz = vhi[c.y].z;
    nlo[p] = make_float4(fminf(ax, bx), fminf(ay, by), fminf(az, bz), 0.f);
    nhi[p] = make_float4(fmaxf(Ax, Bx), fmaxf(Ay, By), fmaxf(Az, Bz), 0.f);
    __threadfence();
    p = parent[p];
  }
}

// ---- 6. emit traversal nodes ---------------------------------------------------------------
__global__ void k_emit_nodes(int n, const int2* __restrict__ children, const float4* __restrict__ nlo,
                             const float4* __restrict__ nhi, BvhNode* __restrict__ nodes, int node_base,
                             int leaf_base) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int2 c = children[i];
  float4 l0 = nlo[c.x], h0 = nhi[c.x], l1 = nlo[c.y], h1 = nhi[c.y];
  BvhNode nd;
  nd.c0xy = make_float4(l0.x, h0.x, l0.y, h0.y);
  nd.c1xy = make_float4(l1.x, h1.x, l1.y, h1.y);
  nd.cz = make_float4(l0.z, h0.z, l1.z, h1.z);
  bool leaf0 = c.x >= n - 1, leaf1 = c.y >= n - 1;
  nd.link.x = leaf0 ? ~((leaf_base + (c.x - (n - 1))) << 3) : node_base + c.x;
  nd.link.y = leaf1 ? ~((leaf_base + (c.y - (n - 1))) << 3) : node_base + c.y;
  nd.link.z = leaf0 ? 1 : 0;
  nd.link.w = leaf1 ? 1 : 0;
  nodes[node_base + i] = nd;
}

// A BVH over a single primitive: one node whose second child is an empty (inverted) box.
__global__ void k_emit_single(const float4* __restrict__ blo, const float4* __restrict__ bhi,
                              BvhNode* __restrict__ nodes, int node_base, int leaf_base, uint32_t* order) {
  float4 l = blo[0], h = bhi[0];
  BvhNode nd;
  nd.c0xy = make_float4(l.x, h.x, l.y, h.y);
  nd.c1xy = make_float4(FLT_MAX, -FLT_MAX, FLT_MAX, -FLT_MAX);
  nd.cz = make_float4(l.z, h.z, FLT_MAX, -FLT_MAX);
  nd.link = make_int4(~(leaf_base << 3), ~(leaf_base << 3), 1, 0);
  nodes[node_base] = nd;
  order[0] = 0;
}

__global__ void k_emit_tris(const AsunaVertex* __restrict__ v, const uint32_t* __restrict__ idx,
                            const uint32_t* __restrict__ order, uint32_t n, TriSlot* __restrict__ tris) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t prim = order[j];
  const float* p0 = v[idx[3 * prim + 0]].pos;
  const float* p1 = v[idx[3 * prim + 1]].pos;
  const float* p2 = v[idx[3 * prim + 2]].pos;
  TriSlot t;
  t.v0 = make_float4(p0[0], p0[1], p0[2], __uint_as_float(prim));
  t.v1 = make_float4(p1[0], p1[1], p1[2], 0.f);
  t.v2 = make_float4(p2[0], p2[1], p2[2], 0.f);
  tris[j] = t;
}

__global__ void k_root_box(int n, const float4* __restrict__ nlo, const float4* __restrict__ nhi,
                           float4* out_lo, float4* out_hi) {
  // root is inner node 0 for n >= 2, the single leaf (index 0) for n == 1
  *out_lo = nlo[0];
  *out_hi = nhi[0];
}

// ---- statistics: SAH cost and depth of an emitted BVH (single-thread-per-node walk up is avoided;
// cost is summed per node from child boxes) ------------------------------------------------------
__device__ __forceinline__ float half_area(float lx, float hx, float ly, float hy, float lz, float hz) {
  float ex = hx - lx, ey = hy - ly, ez = hz - lz;
  return ex * ey + ey * ez + ez * ex;
}
__global__ void k_sah_cost(const BvhNode* __restrict__ nodes, int node_base, int n_nodes, double* cost_sum) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double c = 0.0;
  if (i < n_nodes) {
    BvhNode nd = nodes[node_base + i];
    float a0 = half_area(nd.c0xy.x, nd.c0xy.y, nd.c0xy.z, nd.c0xy.w, nd.cz.x, nd.cz.y);
    float a1 = half_area(nd.c1xy.x, nd.c1xy.y, nd.c1xy.z, nd.c1xy.w, nd.cz.z, nd.cz.w);
    // inner child: traversal step cost 1; leaf child: intersection cost 1 per primitive
    if (nd.c0xy.x <= nd.c0xy.y) c += (double)a0 * (nd.link.x >= 0 ? 1.0 : (double)nd.link.z);
    if (nd.c1xy.x <= nd.c1xy.y) c += (double)a1 * (nd.link.y >= 0 ? 1.0 : (double)nd.link.w);
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31) == 0 && c != 0.0) atomicAdd(cost_sum, c);
}

inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

}  // namespace

// ------------------------------------------------------------------------------------------------
BuildScratch::~BuildScratch() { release(); }
void BuildScratch::release() {
  void* ptrs[] = {blo, bhi, keys[0], keys[1], vals[0], vals[1], hist, children, parent, nlo, nhi, flags, bounds};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  blo = bhi = nlo = nhi = nullptr;
  keys[0] = keys[1] = nullptr;
  vals[0] = vals[1] = nullptr;
  hist = flags = nullptr;
  children = nullptr;
  parent = nullptr;
  bounds = nullptr;
  capacity = 0;
}
cudaError_t BuildScratch::reserve(uint32_t n) {
  if (n <= capacity) return cudaSuccess;
  release();
  uint32_t cap = std::max<uint32_t>(n, 1024);
  cudaError_t e;
#define A(ptr, bytes)                                   \
  if ((e = cudaMalloc((void**)&ptr, (bytes))) != cudaSuccess) return e;
  A(blo, sizeof(float4) * cap);
  A(bhi, sizeof(float4) * cap);
  A(keys[0], sizeof(uint64_t) * cap);
  A(keys[1], sizeof(uint64_t) * cap);
  A(vals[0], sizeof(uint32_t) * cap);
  A(vals[1], sizeof(uint32_t) * cap);
  A(hist, sizeof(uint32_t) * 256 * (size_t)div_up(cap, kSortTile));
  A(children, sizeof(int2) * cap);
  A(parent, sizeof(int) * 2 * (size_t)cap);
  A(nlo, sizeof(float4) * 2 * (size_t)cap);
  A(nhi, sizeof(float4) * 2 * (size_t)cap);
  A(flags, sizeof(uint32_t) * cap);
  A(bounds, sizeof(int) * 8);
#undef A
  capacity = cap;
  return cudaSuccess;
}

void launch_tri_boxes(cudaStream_t s, const AsunaVertex* v, const uint32_t* idx, uint32_t n, BuildScratch& sc) {
  k_bounds_init<<<1, 32, 0, s>>>(sc.bounds);
  k_tri_boxes<<<div_up(n, kThreads), kThreads, 0, s>>>(v, idx, n, sc.blo, sc.bhi, sc.bounds);
}

void launch_instance_boxes(cudaStream_t s, const DInstance* inst, const float4* mesh_lo, const float4* mesh_hi,
                           uint32_t n, BuildScratch& sc) {
  k_bounds_init<<<1, 32, 0, s>>>(sc.bounds);
  k_instance_boxes<<<div_up(n, kThreads), kThreads, 0, s>>>(inst, mesh_lo, mesh_hi, n, sc.blo, sc.bhi, sc.bounds);
}

// Builds the BVH over the n boxes already in sc.blo/bhi (bounds in sc.bounds).  On return (stream
// order) nodes[node_base .. node_base+max(n-1,1)) are written and sc.vals[0] holds the leaf order.
void launch_lbvh(cudaStream_t s, uint32_t n, BvhNode* nodes, int node_base, int leaf_base, BuildScratch& sc,
                 float4* root_lo, float4* root_hi) {
  if (n == 1) {
    k_emit_single<<<1, 1, 0, s>>>(sc.blo, sc.bhi, nodes, node_base, leaf_base, sc.vals[0]);
    if (root_lo) k_root_box<<<1, 1, 0, s>>>(1, sc.blo, sc.bhi, root_lo, root_hi);
    return;
  }
  uint32_t nb = div_up(n, kThreads);
  k_morton<<<nb, kThreads, 0, s>>>(sc.blo, sc.bhi, n, sc.bounds, sc.keys[0], sc.vals[0]);
  uint32_t sort_blocks = div_up(n, kSortTile);
  int cur = 0;
  for (int shift = 0; shift < 64; shift += 8) {
    k_sort_hist<<<sort_blocks, kThreads, 0, s>>>(sc.keys[cur], n, shift, sc.hist, sort_blocks);
    k_sort_scan<<<1, 1024, 0, s>>>(sc.hist, 256u * sort_blocks);
    k_sort_scatter<<<sort_blocks, kThreads, 0, s>>>(sc.keys[cur], sc.vals[cur], sc.keys[cur ^ 1], sc.vals[cur ^ 1],
                                                    n, shift, sc.hist, sort_blocks);
    cur ^= 1;
  }
  // 8 passes: result is back in buffer 0
  k_lbvh_hierarchy<<<div_up(n - 1, kThreads), kThreads, 0, s>>>(sc.keys[0], (int)n, sc.children, sc.parent);
  cudaMemsetAsync(sc.flags, 0, sizeof(uint32_t) * n, s);
  k_lbvh_fit<<<nb, kThreads, 0, s>>>(sc.blo, sc.bhi, sc.vals[0], (int)n, sc.children, sc.parent, sc.nlo, sc.nhi,
                                      sc.flags);
  k_emit_nodes<<<div_up(n - 1, kThreads), kThreads, 0, s>>>((int)n, sc.children, sc.nlo, sc.nhi, nodes, node_base,
                                                            leaf_base);
  if (root_lo) k_root_box<<<1, 1, 0, s>>>((int)n, sc.nlo, sc.nhi, root_lo, root_hi);
}

void launch_emit_tris(cudaStream_t s, const AsunaVertex* v, const uint32_t* idx, uint32_t n, const uint32_t* order,
                      TriSlot* tris) {
  k_emit_tris<<<div_up(n, kThreads), kThreads, 0, s>>>(v, idx, order, n, tris);
}

void launch_sah_cost(cudaStream_t s, const BvhNode* nodes, int node_base, int n_nodes, double* cost_sum) {
  k_sah_cost<<<div_up((uint32_t)n_nodes, kThreads), kThreads, 0, s>>>(nodes, node_base, n_nodes, cost_sum);
}

// Host-side debug/test hook: sorts (key,value) pairs with the builder's radix sort.
cudaError_t radix_sort_pairs(cudaStream_t s, uint64_t* keys_io, uint32_t* vals_io, uint32_t n, BuildScratch& sc) {
  cudaError_t e = sc.reserve(n);
  if (e != cudaSuccess) return e;
  cudaMemcpyAsync(sc.keys[0], keys_io, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(sc.vals[0], vals_io, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s);
  uint32_t sort_blocks = div_up(n, kSortTile);
  int cur = 0;
  for (int shift = 0; shift < 64; shift += 8) {
    k_sort_hist<<<sort_blocks, kThreads, 0, s>>>(sc.keys[cur], n, shift, sc.hist, sort_blocks);
    k_sort_scan<<<1, 1024, 0, s>>>(sc.hist, 256u * sort_blocks);
    k_sort_scatter<<<sort_blocks, kThreads, 0, s>>>(sc.keys[cur], sc.vals[cur], sc.keys[cur ^ 1], sc.vals[cur ^ 1],
                                                    n, shift, sc.hist, sort_blocks);
    cur ^= 1;
  }
  cudaMemcpyAsync(keys_io, sc.keys[0], sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(vals_io, sc.vals[0], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s);
  return cudaGetLastError();
}

}  // namespace asuna
