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
#include "integrator.cuh"
#include "shading.cuh"

namespace asuna {

namespace {

constexpr int kTraceThreads = 128;
constexpr int kShadeThreads = 128;
constexpr int kStackSize = 40;  // uint2 entries: wide-BVH depth of the instance level + one mesh level

struct HitRec {
  float t, b1, b2;
  uint32_t inst, prim;
};

ADEV float safe_rcp(float d) { return 1.0f / (fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); }

struct RaySpace {  // ray constants in the space being traversed (world or one instance's object space)
  float3 o, idir;
  int kx, ky, kz;       // watertight test: axis permutation
  float Sx, Sy, Sz;     // and shear
  float3 d;
  uint32_t octinv4;     // (7 ^ octant) replicated in the four bytes; octant bit k = direction negative on axis k
};

ADEV void setup_shear(RaySpace& r) {
  float ax = fabsf(r.d.x), ay = fabsf(r.d.y), az = fabsf(r.d.z);
  r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
  r.kx = r.kz == 2 ? 0 : r.kz + 1;
  r.ky = r.kx == 2 ? 0 : r.kx + 1;
  float dz = comp(r.d, r.kz);
  if (dz < 0.0f) {
    int t = r.kx;
    r.kx = r.ky;
    r.ky = t;
  }
  r.Sx = comp(r.d, r.kx) / dz;
  r.Sy = comp(r.d, r.ky) / dz;
  r.Sz = 1.0f / dz;
}
ADEV void setup_space(RaySpace& r, float3 o, float3 d) {
  r.o = o;
  r.d = d;
  r.idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
  uint32_t oct = (r.idir.x < 0.f ? 1u : 0u) | (r.idir.y < 0.f ? 2u : 0u) | (r.idir.z < 0.f ? 4u : 0u);
  r.octinv4 = (7u ^ oct) * 0x01010101u;
}

ADEV float byte_f(uint32_t v, int j) { return (float)((v >> (8 * j)) & 0xFFu); }

// Slab test of the eight quantised child boxes of one compressed wide node (Ylitie et al. 2017, section 3):
// plane t = q * (2^e / d) + (p - o) / d.  Returns the hit mask: inner children set bit 24 + (slot ^ octinv)
// (so the highest set bit is the nearest octant), leaf children set their primitive bits [offset, offset+count).
// The far plane is widened by 2 ulp so the box never rejects what the watertight triangle test accepts.
ADEV uint32_t intersect_wide_node(uint4 n0, uint4 n1, uint4 n2, uint4 n3, uint4 n4, const RaySpace& r, float tmin,
                                  float tmax) {
  const float ax = __uint_as_float((n0.w & 0xFFu) << 23) * r.idir.x;
  const float ay = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * r.idir.y;
  const float az = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * r.idir.z;
  const float bx = (__uint_as_float(n0.x) - r.o.x) * r.idir.x;
  const float by = (__uint_as_float(n0.y) - r.o.y) * r.idir.y;
  const float bz = (__uint_as_float(n0.z) - r.o.z) * r.idir.z;
  const bool nx = r.idir.x < 0.f, ny = r.idir.y < 0.f, nz = r.idir.z < 0.f;
  uint32_t hitmask = 0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t meta4 = h ? n1.w : n1.z;
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xFFu;
    const uint32_t bit_index4 = (meta4 ^ (r.octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t qlx = h ? n2.y : n2.x, qly = h ? n2.w : n2.z, qlz = h ? n3.y : n3.x;
    const uint32_t qhx = h ? n3.w : n3.z, qhy = h ? n4.y : n4.x, qhz = h ? n4.w : n4.z;
    const uint32_t nearx = nx ? qhx : qlx, farx = nx ? qlx : qhx;
    const uint32_t neary = ny ? qhy : qly, fary = ny ? qly : qhy;
    const uint32_t nearz = nz ? qhz : qlz, farz = nz ? qlz : qhz;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float tnx = fmaf(byte_f(nearx, j), ax, bx), tfx = fmaf(byte_f(farx, j), ax, bx);
      float tny = fmaf(byte_f(neary, j), ay, by), tfy = fmaf(byte_f(fary, j), ay, by);
      float tnz = fmaf(byte_f(nearz, j), az, bz), tfz = fmaf(byte_f(farz, j), az, bz);
      float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
      float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
      if (cmin <= cmax * 1.0000004f)
        hitmask |= ((child_bits4 >> (8 * j)) & 0xFFu) << ((bit_index4 >> (8 * j)) & 0xFFu);
    }
  }
  return hitmask;
}

// Watertight ray/triangle test, no culling.  Barycentrics in the Vulkan convention.
ADEV bool hit_triangle(const RaySpace& r, float3 v0, float3 v1, float3 v2, float& t, float& b1, float& b2) {
  float3 A = v0 - r.o, B = v1 - r.o, C = v2 - r.o;
  float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
  float Ax = comp(A, r.kx) - r.Sx * Akz, Ay = comp(A, r.ky) - r.Sy * Akz;
  float Bx = comp(B, r.kx) - r.Sx * Bkz, By = comp(B, r.ky) - r.Sy * Bkz;
  float Cx = comp(C, r.kx) - r.Sx * Ckz, Cy = comp(C, r.ky) - r.Sy * Ckz;
  // Edge functions with individually rounded products (no FMA contraction): the neighbour across a
  // shared edge then computes the exact negative, which is what makes the test watertight.
  float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
  float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
  float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {  // edge case: redo the edge functions in double
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  float det = U + V + W;
  if (det == 0.0f) return false;
  float T = U * (r.Sz * Akz) + V * (r.Sz * Bkz) + W * (r.Sz * Ckz);
  float inv = 1.0f / det;
  t = T * inv;
  b1 = V * inv;
  b2 = W * inv;
  return true;
}

// Two-level traversal with traceRayEXT semantics: per instance the ray is taken into object space
// (origin and unnormalised direction through world->object), t is shared between spaces.
// Ties are broken toward the lower (instance, primitive) pair, as the oracle defines.
// Stack entries are (base index, mask) groups: mask > 0x00FFFFFF = a node group (hit bits of inner
// children in the top byte, imask in the low byte), otherwise a primitive group (<= 24 hit bits).
template <bool ANY, bool COUNT = false>
__device__ bool traverse(const SceneView& sc, float3 wo, float3 wd, float tmin, float tmax, HitRec& best,
                         uint32_t* overflow, uint32_t* n_nodes = nullptr, uint32_t* n_tris = nullptr) {
  if (sc.n_instances == 0) return false;
  uint2 stack[kStackSize];
  int sp = 0, blas_sp = 0;
  RaySpace rs;
  setup_space(rs, wo, wd);
  bool in_blas = false, found = false;
  uint32_t cur_inst = 0;
  const WideNode* nodes = sc.tlas_nodes;
  uint2 ng = make_uint2(0u, 0x80000000u), tg = make_uint2(0u, 0u);
  for (;;) {
    if (ng.y > 0x00FFFFFFu) {
      const uint32_t hits = ng.y;
      const uint32_t bit = 31u - (uint32_t)__clz(hits);
      ng.y &= ~(1u << bit);
      if (ng.y > 0x00FFFFFFu) {
        if (sp < kStackSize) stack[sp++] = ng;
        else atomicAdd(overflow, 1u);
      }
      const uint32_t slot = (bit - 24u) ^ (rs.octinv4 & 7u);
      const uint32_t rel = __popc(hits & 0xFFu & ~(0xFFFFFFFFu << slot));
      const uint4* np = reinterpret_cast<const uint4*>(nodes + ng.x + rel);
      if (COUNT) (*n_nodes)++;
      const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
      const uint32_t hm = intersect_wide_node(n0, n1, n2, n3, n4, rs, tmin, tmax);
      ng = make_uint2(n1.x, (hm & 0xFF000000u) | (n0.w >> 24));
      tg = make_uint2(n1.y, hm & 0x00FFFFFFu);
    } else {
      tg = ng;
      ng = make_uint2(0u, 0u);
    }
    if (in_blas) {
      while (tg.y) {
        const uint32_t k = (uint32_t)__ffs((int)tg.y) - 1u;
        tg.y &= tg.y - 1u;
        const TriSlot* tp = sc.tris + tg.x + k;
        float4 v0 = __ldg(&tp->v0), v1 = __ldg(&tp->v1), v2 = __ldg(&tp->v2);
        float t, b1, b2;
        if (COUNT) (*n_tris)++;
        if (!hit_triangle(rs, f3(v0), f3(v1), f3(v2), t, b1, b2)) continue;
        if (!(t > tmin)) continue;
        uint32_t prim = __float_as_uint(v0.w);
        bool closer = t < tmax || (t == tmax && found &&
                                   (cur_inst < best.inst || (cur_inst == best.inst && prim < best.prim)));
        if (!closer) continue;
        best.t = t, best.b1 = b1, best.b2 = b2, best.inst = cur_inst, best.prim = prim;
        tmax = t;
        found = true;
        if (ANY) return true;
      }
    } else if (tg.y) {
      // instance group: enter the first instance, keep the rest (and the pending node group) for later
      const uint32_t k = (uint32_t)__ffs((int)tg.y) - 1u;
      tg.y &= tg.y - 1u;
      if (sp + 2 > kStackSize) {
        atomicAdd(overflow, 1u);
      } else {
        if (tg.y) stack[sp++] = tg;
        if (ng.y > 0x00FFFFFFu) stack[sp++] = ng;
        cur_inst = __ldg(&sc.tlas_leaf_inst[tg.x + k]);
        const DInstance* in = sc.instances + cur_inst;
        float4 r0 = __ldg(&in->w2o[0]), r1 = __ldg(&in->w2o[1]), r2 = __ldg(&in->w2o[2]);
        float4 m[3] = {r0, r1, r2};
        setup_space(rs, xf_point(m, wo), xf_vector(m, wd));
        setup_shear(rs);
        in_blas = true;
        blas_sp = sp;
        nodes = sc.blas_nodes;
        ng = make_uint2((uint32_t)__ldg(&in->blas_root), 0x80000000u);
        tg = make_uint2(0u, 0u);
        continue;
      }
    }
    if (ng.y <= 0x00FFFFFFu) {
      if (in_blas && sp == blas_sp) {  // this instance is exhausted: back to world space
        in_blas = false;
        nodes = sc.tlas_nodes;
        setup_space(rs, wo, wd);
      }
      if (sp == 0) break;
      ng = stack[--sp];
    }
  }
  return found;
}

// ---- trace kernels: persistent, warp-granular dynamic fetch -----------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(kTraceThreads)
k_trace_closest(const __grid_constant__ SceneView sc, PathState ps, Counters* cnt, int iter, int qsel) {
  const uint32_t count = cnt->queue[iter];
  const uint32_t* queue = ps.queue[qsel];
  const int lane = threadIdx.x & 31;
  uint32_t n_nodes = 0, n_tris = 0;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&cnt->ticket_closest[iter], 32u);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= count) break;
    uint32_t i = base + lane;
    if (i < count) {
      uint32_t slot = queue[i];
      float4 o = ps.ray_o[slot], d = ps.ray_d[slot];
      HitRec h;
      h.inst = 0xFFFFFFFFu, h.prim = 0xFFFFFFFFu, h.b1 = h.b2 = h.t = 0.f;
      traverse<false, COUNT>(sc, f3(o), f3(d), kMinimum, kInfinity, h, &cnt->stack_overflow, &n_nodes, &n_tris);
      ps.hit[slot] = make_uint4(__float_as_uint(h.b1), __float_as_uint(h.b2), h.inst, h.prim);
    }
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      n_nodes += __shfl_xor_sync(0xFFFFFFFFu, n_nodes, o);
      n_tris += __shfl_xor_sync(0xFFFFFFFFu, n_tris, o);
    }
    if (lane == 0) {
      atomicAdd(&cnt->node_visits, (unsigned long long)n_nodes);
      atomicAdd(&cnt->tri_tests, (unsigned long long)n_tris);
    }
  }
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
}

__global__ void __launch_bounds__(kTraceThreads) k_trace_shadow(const __grid_constant__ SceneView sc, PathState ps, Counters* cnt, int iter) {
  const uint32_t count = cnt->shadow[iter];
  const int lane = threadIdx.x & 31;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&cnt->ticket_shadow[iter], 32u);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= count) break;
    uint32_t i = base + lane;
    if (i < count) {
      float4 o = ps.sh_o[i], d = ps.sh_d[i];
      HitRec h;
      // rgen:117-122: tmin 0, tmax = dist - 2 EPS, terminate on first hit
      bool occluded = traverse<true>(sc, f3(o), f3(d), 0.0f, o.w, h, &cnt->stack_overflow);
      if (!occluded) {
        uint32_t slot = __float_as_uint(d.w);
        float4 L = ps.sh_l[i], r = ps.rad[slot];
        ps.rad[slot] = make_float4(r.x + L.x, r.y + L.y, r.z + L.z, r.w);
      }
    }
  }
}

// Generic ray queries for the parity tests (asuna_trace_rays / asuna_occlusion_rays / asuna_trace_primary).
__global__ void k_trace_user(const __grid_constant__ SceneView sc, const float4* rays, uint32_t n, float* tuv, uint32_t* inst_prim,
                             uint8_t* occluded, Counters* cnt) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = rays[2 * i], b = rays[2 * i + 1];
  HitRec h;
  h.inst = 0xFFFFFFFFu, h.prim = 0xFFFFFFFFu, h.b1 = h.b2 = h.t = 0.f;
  if (occluded) {
    occluded[i] = traverse<true>(sc, f3(a), f3(b), a.w, b.w, h, &cnt->stack_overflow) ? 1 : 0;
  } else {
    bool f = traverse<false>(sc, f3(a), f3(b), a.w, b.w, h, &cnt->stack_overflow);
    if (tuv) tuv[3 * i] = f ? h.t : 0.f, tuv[3 * i + 1] = h.b1, tuv[3 * i + 2] = h.b2;
    inst_prim[2 * i] = h.inst, inst_prim[2 * i + 1] = h.prim;
  }
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

__global__ void __launch_bounds__(kShadeThreads)
k_shade(const __grid_constant__ SceneView sc, const __grid_constant__ FrameParams fp, PathState ps, OutputImages out, Counters* cnt, int iter,
        int qsel) {
  const uint32_t count = cnt->queue[iter];
  const uint32_t* queue = ps.queue[qsel];
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

  for (uint32_t base = warp_global * 32; base < count; base += n_warps * 32) {
    uint32_t i = base + lane;
    bool valid = i < count;
    bool cont = false, nee = false, incoherent = false;
    uint32_t slot = 0;
    PathRegs p;
    p.nee = false;
    if (valid) {
      slot = queue[i];
      float4 ro = ps.ray_o[slot], rd = ps.ray_d[slot], th = ps.thr[slot], ra = ps.rad[slot];
      uint4 hit = ps.hit[slot];
      uint32_t fi = slot / fp.n_pixels, pixel = slot - fi * fp.n_pixels;
      bool frame0 = fp.frame_ids[fi] == 0;
      for (int c = 0; c < ASUNA_NUM_OUTPUT_IMAGES - 1; c++) se.aov[c] = frame0 ? out.img[c + 1] : nullptr;
      p.ray_o = f3(ro), p.ray_d = f3(rd), p.throughput = f3(th), p.radiance = f3(ra);
      p.seed = __float_as_uint(ro.w);
      p.bsdf_pdf = rd.w;
      uint32_t packed = __float_as_uint(th.w);
      p.depth = (int)(packed & 0xFFFFu);
      p.bsdf_flags = packed >> 16;
      p.stop = false;
      p.nee_L = f3(0.0f);
      if (hit.z == 0xFFFFFFFFu) {
        shade_miss(se, p);
      } else {
        const DInstance in = sc.instances[hit.z];
        Surface s;
        load_surface(sc, fp.pc, in, hit, p.ray_d, s);
        if (in.light >= 0) {
          shade_light_hit(se, p, in.light, s.pos);
        } else {
          const AsunaMaterial m = sc.materials[in.material];
          switch (m.type) {
            case ASUNA_MAT_LAMBERTIAN: shade_lambertian(se, p, s, m, pixel); break;
            case ASUNA_MAT_EMISSIVE: shade_emissive(se, p, s, m); break;
            case ASUNA_MAT_DIELECTRIC: shade_dielectric(se, p, s, m); break;
            case ASUNA_MAT_CONDUCTOR: shade_conductor(se, p, s, m); break;
            case ASUNA_MAT_PLASTIC: shade_plastic(se, p, s, m); break;
            case ASUNA_MAT_ROUGH_PLASTIC: shade_rough_plastic(se, p, s, m); break;
            case ASUNA_MAT_PBR_METALNESS_ROUGHNESS: shade_pbr(se, p, s, m, pixel); break;
            case ASUNA_MAT_KANG18: shade_kang18(se, p, s, m, in, pixel); break;
            default: p.stop = true; break;
          }
        }
      }
      // rgen:112-131: shadow ray if the hit shader asked for one, then stop / depth++
      nee = p.nee && (p.nee_L.x != 0.0f || p.nee_L.y != 0.0f || p.nee_L.z != 0.0f);  // zero NEE adds nothing (A.3-4)
      int next_depth = p.depth + 1;
      cont = !p.stop && next_depth <= fp.pc.maxPathDepth;
      incoherent = cont && next_depth >= 2;
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
  if (counting) k_trace_closest<true><<<ld.trace_blocks, kTraceThreads, 0, s>>>(sc, ps, cnt, iter, qsel);
  else k_trace_closest<false><<<ld.trace_blocks, kTraceThreads, 0, s>>>(sc, ps, cnt, iter, qsel);
}
void launch_fold_counters(cudaStream_t s, const Counters* cnt, Totals* tot, int iters) {
  k_fold_counters<<<1, 1, 0, s>>>(cnt, tot, iters);
}
void launch_shade(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const FrameParams& fp, const PathState& ps,
                  const OutputImages& out, Counters* cnt, int iter, int qsel) {
  k_shade<<<ld.shade_blocks, kShadeThreads, 0, s>>>(sc, fp, ps, out, cnt, iter, qsel);
}
void launch_trace_shadow(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const PathState& ps, Counters* cnt,
                         int iter) {
  k_trace_shadow<<<ld.trace_blocks, kTraceThreads, 0, s>>>(sc, ps, cnt, iter);
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
void launch_primary_rays(cudaStream_t s, const FrameParams& fp, float4* rays) {
  k_primary_rays<<<div_up(fp.n_pixels, 256), 256, 0, s>>>(fp, rays);
}
void launch_trace_user(cudaStream_t s, const SceneView& sc, const float4* rays, uint32_t n, float* tuv,
                       uint32_t* inst_prim, uint8_t* occluded, Counters* cnt) {
  k_trace_user<<<div_up(n, 128), 128, 0, s>>>(sc, rays, n, tuv, inst_prim, occluded, cnt);
}

cudaError_t query_launch_dims(LaunchDims& ld, int sm_count) {
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_closest<false>, kTraceThreads, 0);
  if (e != cudaSuccess) return e;
  ld.trace_blocks = (uint32_t)(sm_count * (per_sm > 0 ? per_sm : 1));
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_shade, kShadeThreads, 0);
  if (e != cudaSuccess) return e;
  ld.shade_blocks = (uint32_t)(sm_count * (per_sm > 0 ? per_sm : 1));
  return cudaSuccess;
}

}  // namespace asuna
