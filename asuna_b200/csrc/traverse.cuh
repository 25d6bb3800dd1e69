// Ray traversal of the two-level compressed wide BVH, hand-written for sm_100a (no RT cores on B200).
//
// Replaces traceRayEXT of the reference (src/shaders/raytrace.projective.rgen:108-109 closest hit,
// :117-122 shadow ray).  Semantics kept: per instance the ray is taken into object space (origin and
// unnormalised direction through world->object) and t is shared between spaces; no culling; all
// geometry opaque.  Ties (unspecified by Vulkan) go to the lower (instance, primitive) pair, as the
// oracle defines.  The triangle test is the watertight test of Woop, Benthin, Wald 2013 written
// without the axis permutation (three constant vectors per ray instead), bit-identical to
// oracle/bvh.h.
//
// Execution model: persistent warps.  Every lane owns one ray at a time; the warp loops over
//   pop / leave instance / terminate  ->  one wide-node step  ->  instance entry  ->  one triangle step
// where the triangle step only runs when an eighth of the live lanes want it (SceneView::tri_vote_shift) (or nothing else is left):
// lanes with a pending triangle group swap it under the next node group of their stack and keep
// traversing ("triangle postponing", Ylitie et al. 2017 section 5).  Finished lanes are refilled from the
// ray queue with one atomic per warp as soon as `refill_lanes` of them are idle.
#pragma once
#include <cuda_fp16.h>

#include "device_types.cuh"
#include "vec.cuh"

namespace asuna {

#ifndef ASUNA_TRACE_THREADS
#define ASUNA_TRACE_THREADS 128
#endif
constexpr int kTraceThreads = ASUNA_TRACE_THREADS;
#ifndef ASUNA_TRACE_MIN_BLOCKS
#define ASUNA_TRACE_MIN_BLOCKS 6  // two-level kernels, 80 registers: 24 resident warps per SM (measured best; 7 spills)
#endif
#ifndef ASUNA_TRACE_MIN_BLOCKS_SINGLE
#define ASUNA_TRACE_MIN_BLOCKS_SINGLE 7  // single-level kernels fit 72 registers without spilling: 28 warps per SM (+5 %)
#endif
#ifndef ASUNA_SHADOW_MIN_BLOCKS_SINGLE
#define ASUNA_SHADOW_MIN_BLOCKS_SINGLE 8  // the any-hit kernel keeps less state: 64 registers, 32 warps per SM (+5-7 % shadow rays/s over 7)
#endif
constexpr int kStackSize = 40;        // uint2 entries: wide-BVH depth of the instance level + one mesh level

struct HitRec {
  float t, b1, b2;
  uint32_t inst, prim;
};
constexpr uint32_t kInstMask = 0x0FFFFFFFu;  // instance id in a world-space triangle slot; the 4 bits above hold the hit kind
constexpr uint32_t kKindUnknown = 0xFFu;

// Reciprocal for the slab tests only (MUFU.RCP, ~1 ulp): box tests are conservative by half a quantisation step
// plus 2 ulp, results never depend on it.  The triangle test uses IEEE divisions (bit-exact with the oracle).
ADEV float safe_rcp(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)));
  return r;
}

struct RaySpace {  // ray constants in the space being traversed (world or one instance's object space)
  float3 o, idir;
  float3 sx, sy, sz;  // watertight test: rows of the shear (e_kx - Sx e_kz, e_ky - Sy e_kz, Sz e_kz)
  uint32_t oct;       // bit k set = direction negative on axis k
};

ADEV void setup_space(RaySpace& r, float3 o, float3 d) {
  r.o = o;
  r.idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
  // sign bits of the direction (rcp keeps them): three shifts instead of three compare/select pairs
  r.oct = (__float_as_uint(r.idir.x) >> 31) | ((__float_as_uint(r.idir.y) >> 31) << 1) |
          ((__float_as_uint(r.idir.z) >> 31) << 2);
}
// kz = dominant axis of d, kx = kz+1, ky = kz+2 (cyclic).  The kx/ky swap of the paper only flips the sign of
// all three edge functions together (exactly, in floating point), which changes no result without culling.
ADEV void setup_shear(RaySpace& r, float3 d) {
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  int kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
  float dz = comp(d, kz);
  float dx = kz == 0 ? d.y : (kz == 1 ? d.z : d.x);
  float dy = kz == 0 ? d.z : (kz == 1 ? d.x : d.y);
  float Sx = dx / dz, Sy = dy / dz, Sz = 1.0f / dz;
  if (kz == 0) r.sx = f3(-Sx, 1.f, 0.f), r.sy = f3(-Sy, 0.f, 1.f), r.sz = f3(Sz, 0.f, 0.f);
  else if (kz == 1) r.sx = f3(0.f, -Sx, 1.f), r.sy = f3(1.f, -Sy, 0.f), r.sz = f3(0.f, Sz, 0.f);
  else r.sx = f3(1.f, 0.f, -Sx), r.sy = f3(0.f, 1.f, -Sy), r.sz = f3(0.f, 0.f, Sz);
}

ADEV float dot_chain(float3 s, float3 a) { return __fmaf_rn(s.z, a.z, __fmaf_rn(s.y, a.y, __fmul_rn(s.x, a.x))); }

// Watertight ray/triangle test, no culling.  Barycentrics in the Vulkan convention.
ADEV bool hit_triangle(const RaySpace& r, float3 v0, float3 v1, float3 v2, float& t, float& b1, float& b2) {
  float3 A = v0 - r.o, B = v1 - r.o, C = v2 - r.o;
  float Ax = dot_chain(r.sx, A), Ay = dot_chain(r.sy, A);
  float Bx = dot_chain(r.sx, B), By = dot_chain(r.sy, B);
  float Cx = dot_chain(r.sx, C), Cy = dot_chain(r.sy, C);
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
  float det = __fadd_rn(__fadd_rn(U, V), W);
  if (det == 0.0f) return false;
  float Az = dot_chain(r.sz, A), Bz = dot_chain(r.sz, B), Cz = dot_chain(r.sz, C);
  float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));
  float inv = 1.0f / det;
  t = T * inv;
  b1 = V * inv;
  b2 = W * inv;
  return true;
}

// 0x4B000000 | byte = 2^23 + byte exactly: one PRMT turns a quantised plane into a float without the XU pipe.
// `magic` must come from a register / the constant bank so that the selector stays an immediate.
template <int J>
ADEV float magic_byte(uint32_t v, uint32_t magic) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(magic), "n"(0x7440 + J));
  return __uint_as_float(d);
}

// Slab test of the eight quantised child boxes of one compressed wide node (Ylitie et al. 2017, section 3).
// Plane t = q * a + b with a = 2^e / d, b = (p - o) / d; q arrives as 2^23 + q, so b is pre-biased by
// -2^23 a, whose rounding error (<= |a| / 2, half a quantisation step) is covered by moving the near planes
// half a step down and the far planes half a step up.  The exit distance is also widened by ~3 ulp so that a box
// never rejects what the watertight triangle test accepts.
//
// Returns the hit mask of the node in stack-entry form: bit 24 + (slot ^ octinv) for a hit inner child (so that the
// highest set bit is the child nearest along the ray's octant), bits [offset, offset + count) of the primitive
// slots of a hit leaf child.  Both come from the meta byte (count/"1" in its top 3 bits, offset/"24 + slot" in
// its low 5) handled four children at a time, as in the paper's listing.
template <int J>
ADEV uint32_t byte_of(uint32_t x) { return __byte_perm(x, 0u, 0x4440 + J); }

// Two quantised planes per PRMT: bytes (2K, 2K+1) of v become the half2 (1024 + q, 1024 + q') -- exact in fp16 --
// and two HADD2.F32 on the FMA pipe widen them.  Trades one ALU-pipe PRMT for two FMA-pipe conversions per pair;
// ASUNA_HALF_UNPACK selects which planes go this way (0 none, 1 the far planes, 2 all).
#ifndef ASUNA_HALF_UNPACK
#define ASUNA_HALF_UNPACK 0
#endif
template <int K>
ADEV float2 half_pair(uint32_t v, uint32_t magic_h) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(magic_h), "n"(K == 0 ? 0x4140 : 0x4342));
  return __half22float2(*reinterpret_cast<__half2*>(&d));
}
template <bool HALF>
ADEV void unpack4(uint32_t v, uint32_t magic, float f[4]) {
  if (HALF) {
    const float2 a = half_pair<0>(v, 0x64646464u), b = half_pair<1>(v, 0x64646464u);
    f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y;
  } else {
    f[0] = magic_byte<0>(v, magic), f[1] = magic_byte<1>(v, magic), f[2] = magic_byte<2>(v, magic), f[3] = magic_byte<3>(v, magic);
  }
}

// Two plane distances per instruction: Blackwell's packed fp32 FMA (fma.rn.f32x2 -> FFMA2; the scale and the bias are
// broadcast from single registers).  Each half is an ordinary IEEE fp32 fma, so results equal the scalar form bit for
// bit, and it halves the FMA share of the slab test.
// Measured on B200 (profiles/README.md, round 2): the closest-hit kernel gets 7 % SLOWER (2.58 -> 2.41 G rays/s: the
// 64-bit-aligned register pairs push it over its 72-register budget into spills, and FFMA2 saves no ALU-pipe work, which
// is what binds), the shadow kernel 3.5 % faster.  Off by default.
#ifndef ASUNA_FFMA2
#define ASUNA_FFMA2 0
#endif
ADEV void fma_pair(float q0, float q1, float a, float b, float& r0, float& r1) {
#if ASUNA_FFMA2
  unsigned long long pq, pa, pb, pr;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pq) : "f"(q0), "f"(q1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pr) : "l"(pq), "l"(pa), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(pr));
#else
  r0 = fmaf(q0, a, b), r1 = fmaf(q1, a, b);
#endif
}

ADEV uint32_t intersect_wide_node(uint4 n0, uint4 n1, uint4 n2, uint4 n3, uint4 n4, const RaySpace& r, float tmin,
                                  float tmax, uint32_t magic) {
  constexpr bool kHalfNear = ASUNA_HALF_UNPACK >= 2, kHalfFar = ASUNA_HALF_UNPACK >= 1;
  const float kWiden = 1.0000004f;
  const float ax = __uint_as_float((n0.w & 0xFFu) << 23) * r.idir.x;
  const float ay = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * r.idir.y;
  const float az = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * r.idir.z;
  const float cx = (__uint_as_float(n0.x) - r.o.x) * r.idir.x, cy = (__uint_as_float(n0.y) - r.o.y) * r.idir.y,
              cz = (__uint_as_float(n0.z) - r.o.z) * r.idir.z;
  const float hx = 0.50001f * fabsf(ax), hy = 0.50001f * fabsf(ay), hz = 0.50001f * fabsf(az);
  const float kBn = kHalfNear ? 1024.0f : 8388608.0f, kBf = kHalfFar ? 1024.0f : 8388608.0f;
  const float bnx = fmaf(-kBn, ax, cx) - hx, bny = fmaf(-kBn, ay, cy) - hy, bnz = fmaf(-kBn, az, cz) - hz;
  // far planes: the ~3 ulp widening is folded into scale and bias (6 multiplies per node instead of 8 per node on
  // the results; still conservative: each product is within half an ulp of the widened value)
#ifdef ASUNA_FOLD_WIDEN
  const float afx = ax * kWiden, afy = ay * kWiden, afz = az * kWiden;
  const float bfx = (fmaf(-kBf, ax, cx) + hx) * kWiden, bfy = (fmaf(-kBf, ay, cy) + hy) * kWiden,
              bfz = (fmaf(-kBf, az, cz) + hz) * kWiden;
  const float kW = 1.0f;
#else
  const float afx = ax, afy = ay, afz = az, kW = kWiden;
  const float bfx = fmaf(-kBf, ax, cx) + hx, bfy = fmaf(-kBf, ay, cy) + hy, bfz = fmaf(-kBf, az, cz) + hz;
#endif
  const bool nx = (r.oct & 1u) != 0u, ny = (r.oct & 2u) != 0u, nz = (r.oct & 4u) != 0u;
  const uint32_t octinv4 = (7u ^ r.oct) * 0x01010101u;
  uint32_t mask = 0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t meta4 = h ? n1.w : n1.z;
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;  // offsets >= 24 mark inner children
    uint32_t inner_mask4;                                             // 0xFF in the bytes of inner children
    asm("prmt.b32 %0, %1, %2, 0xba98;" : "=r"(inner_mask4) : "r"(is_inner4 << 3), "r"(0u));
    const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t qlx = h ? n2.y : n2.x, qly = h ? n2.w : n2.z, qlz = h ? n3.y : n3.x;
    const uint32_t qhx = h ? n3.w : n3.z, qhy = h ? n4.y : n4.x, qhz = h ? n4.w : n4.z;
    float qnx[4], qny[4], qnz[4], qfx[4], qfy[4], qfz[4];
    unpack4<kHalfNear>(nx ? qhx : qlx, magic, qnx), unpack4<kHalfFar>(nx ? qlx : qhx, magic, qfx);
    unpack4<kHalfNear>(ny ? qhy : qly, magic, qny), unpack4<kHalfFar>(ny ? qly : qhy, magic, qfy);
    unpack4<kHalfNear>(nz ? qhz : qlz, magic, qnz), unpack4<kHalfFar>(nz ? qlz : qhz, magic, qfz);
#define ASUNA_CHILD(J, TNX, TNY, TNZ, TFX, TFY, TFZ)                                                          \
    {                                                                                                         \
      float cmin = fmaxf(fmaxf(TNX, TNY), fmaxf(TNZ, tmin));                                                  \
      float cmax = kW == 1.0f ? fminf(fminf(TFX, TFY), fminf(TFZ, tmax)) : fminf(fminf(fminf(TFX, TFY), TFZ) * kW, tmax); \
      if (cmin <= cmax) mask |= byte_of<J>(child_bits4) << byte_of<J>(bit_index4);                            \
    }
#define ASUNA_CHILD_PAIR(J)                                                                                  \
    {                                                                                                         \
      float nx0, nx1, ny0, ny1, nz0, nz1, fx0, fx1, fy0, fy1, fz0, fz1;                                       \
      fma_pair(qnx[J], qnx[J + 1], ax, bnx, nx0, nx1), fma_pair(qfx[J], qfx[J + 1], afx, bfx, fx0, fx1);      \
      fma_pair(qny[J], qny[J + 1], ay, bny, ny0, ny1), fma_pair(qfy[J], qfy[J + 1], afy, bfy, fy0, fy1);      \
      fma_pair(qnz[J], qnz[J + 1], az, bnz, nz0, nz1), fma_pair(qfz[J], qfz[J + 1], afz, bfz, fz0, fz1);      \
      ASUNA_CHILD(J, nx0, ny0, nz0, fx0, fy0, fz0) ASUNA_CHILD(J + 1, nx1, ny1, nz1, fx1, fy1, fz1)           \
    }
    ASUNA_CHILD_PAIR(0) ASUNA_CHILD_PAIR(2)
#undef ASUNA_CHILD_PAIR
#undef ASUNA_CHILD
  }
  return mask;
}

// Per-lane traversal state.  Stack entries are (base index, mask) groups: mask > 0x00FFFFFF = a node group
// (hit bits of inner children in the top byte, already in visiting order: highest bit = nearest octant; imask in
// the low byte), otherwise a primitive group (<= 24 hit bits of consecutive primitive slots).  The newest stack
// entry lives in registers (`top`); local memory holds the ones below it, so the common pop / peek / swap of the
// postponing logic never waits on a load.
struct Lane {
  int sp, blas_sp;
  uint2 ng, tg, top;
  RaySpace rs;
  float3 wo, wd;
  float tmin, tmax;
  HitRec best;  // best.inst == ~0: nothing found yet
  uint32_t cur_inst;
  bool in_blas;
  bool shear_ok;  // the watertight-test constants of the current instance are computed at its first triangle
};

// SINGLE = every instance lives in the merged world-space BLAS (SceneView::single_root): there is no instance
// level, the ray stays in world space and the instance id of a hit comes from its triangle slot.
template <bool SINGLE>
ADEV void lane_begin(Lane& L, const SceneView& sc, float3 o, float3 d, float tmin, float tmax) {
  L.sp = 0, L.blas_sp = 0;
  L.ng = make_uint2(SINGLE ? sc.single_root : 0u, 0x80000000u);  // the root (instance level unless SINGLE)
  L.tg = make_uint2(0u, 0u);
  L.wo = o, L.wd = d, L.tmin = tmin, L.tmax = tmax;
  setup_space(L.rs, o, d);
  L.in_blas = SINGLE, L.shear_ok = SINGLE;
  if (SINGLE) setup_shear(L.rs, d);
  L.cur_inst = 0;
  L.best.inst = 0xFFFFFFFFu, L.best.prim = 0xFFFFFFFFu, L.best.b1 = L.best.b2 = L.best.t = 0.f;
}

// Where the stack entries below the register-resident top live.  Default: a per-thread local-memory array (L1-cached,
// 64-bit STL / LDL).  ASUNA_SMEM_STACK = N > 0 is the "short shared-memory stack" variant BASELINE.json's north star
// names: the N entries nearest the top of an empty stack sit in shared memory ([entry][thread] layout, conflict-free),
// deeper ones overflow to the local array.  Measured on B200 (profiles/README.md, round 2) -- see there for why the
// local-memory form stays the default.
#ifndef ASUNA_SMEM_STACK
#define ASUNA_SMEM_STACK 0
#endif
struct StackMem {
  uint2* local;
#if ASUNA_SMEM_STACK > 0
  uint2* shared;  // &smem[threadIdx.x], stride kTraceThreads
#endif
  ADEV void store(int i, uint2 e) const {
#if ASUNA_SMEM_STACK > 0
    if (i < ASUNA_SMEM_STACK) shared[i * kTraceThreads] = e;
    else
#endif
      local[i] = e;
  }
  ADEV uint2 load(int i) const {
#if ASUNA_SMEM_STACK > 0
    if (i < ASUNA_SMEM_STACK) return shared[i * kTraceThreads];
#endif
    return local[i];
  }
};
// entry k of the storage holds stack entry k - 1 (slot 0 is a dummy), so push and pop need no "is there an entry below" branch
ADEV void lane_push(Lane& L, const StackMem& stack, uint2 e, uint32_t* overflow) {
  if (L.sp < kStackSize) {
    stack.store(L.sp++, L.top);
    L.top = e;
  } else {
    atomicAdd(overflow, 1u);
  }
}
ADEV uint2 lane_pop(Lane& L, const StackMem& stack) {
  const uint2 e = L.top;
  L.top = stack.load(--L.sp);
  return e;
}

// One wide-node step: take the nearest pending child of the current node group, test its eight children.
template <bool COUNT, bool SINGLE>
ADEV void lane_node_step(Lane& L, const StackMem& stack, const SceneView& sc, uint32_t* overflow, uint32_t& n_nodes) {
  const uint32_t hits = L.ng.y;
  const uint32_t bit = 31u - (uint32_t)__clz(hits);
  uint2 rest = make_uint2(L.ng.x, hits & ~(1u << bit));
  const uint32_t slot = (bit - 24u) ^ 7u ^ L.rs.oct;
  const uint32_t rel = __popc(hits & 0xFFu & ~(0xFFFFFFFFu << slot));
  const WideNode* nodes = (SINGLE || L.in_blas) ? sc.blas_nodes : sc.tlas_nodes;
  const uint4* np = reinterpret_cast<const uint4*>(nodes + L.ng.x + rel);
  const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
  if (COUNT) n_nodes++;
  if (rest.y > 0x00FFFFFFu) lane_push(L, stack, rest, overflow);
  const uint32_t mask = intersect_wide_node(n0, n1, n2, n3, n4, L.rs, L.tmin, L.tmax, sc.magic);
  L.ng = make_uint2(n1.x, (mask & 0xFF000000u) | (n0.w >> 24));
  if (mask & 0x00FFFFFFu) {
    if (L.tg.y) lane_push(L, stack, L.tg, overflow);  // an older postponed group goes under the newer, nearer one
    L.tg = make_uint2(n1.y, mask & 0x00FFFFFFu);
  }
}

// Instance group at the top level: enter its first instance, keep the rest (and the pending node group) for later.
ADEV void lane_enter_instance(Lane& L, const StackMem& stack, const SceneView& sc, uint32_t* overflow) {
  const uint32_t k = (uint32_t)__ffs((int)L.tg.y) - 1u;
  L.tg.y &= L.tg.y - 1u;
  if (L.sp + 2 > kStackSize) {
    atomicAdd(overflow, 1u);
    L.tg.y = 0;
    return;
  }
  if (L.tg.y) lane_push(L, stack, L.tg, overflow);
  if (L.ng.y > 0x00FFFFFFu) lane_push(L, stack, L.ng, overflow);
  L.cur_inst = __ldg(&sc.tlas_leaf_inst[L.tg.x + k]);
  const DInstance* in = sc.instances + L.cur_inst;
  float4 r0 = __ldg(&in->w2o[0]), r1 = __ldg(&in->w2o[1]), r2 = __ldg(&in->w2o[2]);
  float4 m[3] = {r0, r1, r2};
  setup_space(L.rs, xf_point(m, L.wo), xf_vector(m, L.wd));
  L.shear_ok = false;
  L.in_blas = true;
  L.blas_sp = L.sp;
  L.ng = make_uint2((uint32_t)__ldg(&in->blas_root), 0x80000000u);
  L.tg = make_uint2(0u, 0u);
}

// One triangle of the current primitive group.  Returns true when the hit ends an any-hit query.
template <bool ANY, bool COUNT, bool SINGLE>
ADEV bool lane_triangle_step(Lane& L, const SceneView& sc, uint32_t& n_tris) {
  const uint32_t k = (uint32_t)__ffs((int)L.tg.y) - 1u;
  L.tg.y &= L.tg.y - 1u;
  const TriSlot* tp = sc.tris + L.tg.x + k;
  float4 v0 = __ldg(&tp->v0), v1 = __ldg(&tp->v1), v2 = __ldg(&tp->v2);
  float t, b1, b2;
  if (COUNT) n_tris++;
  if (!SINGLE && !L.shear_ok) {  // many instance visits never reach a triangle: the three IEEE divisions are paid only here
    const DInstance* in = sc.instances + L.cur_inst;
    float4 m[3] = {__ldg(&in->w2o[0]), __ldg(&in->w2o[1]), __ldg(&in->w2o[2])};
    setup_shear(L.rs, xf_vector(m, L.wd));
    L.shear_ok = true;
  }
  if (!hit_triangle(L.rs, f3(v0), f3(v1), f3(v2), t, b1, b2)) return false;
  if (!(t > L.tmin)) return false;
  const uint32_t prim = __float_as_uint(v0.w);
  // world-space slots carry (hit kind << 28 | instance) in v1.w (k_world_triangles_batched): the single-level kernels keep
  // the packed word, so that committing a hit needs no look-up of the instance's material type
  const uint32_t inst = SINGLE ? __float_as_uint(v1.w)
                               : (L.cur_inst == sc.world_inst ? (__float_as_uint(v1.w) & kInstMask) : L.cur_inst);
  bool closer = t < L.tmax;
  if (!closer && t == L.tmax && L.best.inst != 0xFFFFFFFFu) {  // tie: lower (instance, primitive) wins
    const uint32_t a = SINGLE ? (inst & kInstMask) : inst, b = SINGLE ? (L.best.inst & kInstMask) : L.best.inst;
    closer = a < b || (a == b && prim < L.best.prim);
  }
  if (!closer) return false;
  L.best.t = t, L.best.b1 = b1, L.best.b2 = b2, L.best.inst = inst, L.best.prim = prim;
  L.tmax = t;
  return ANY;
}

#ifndef ASUNA_TICKET_CHUNK
#define ASUNA_TICKET_CHUNK 64  // rays a warp takes from the queue per atomic (shrunk for short queues)
#endif

// Persistent warp loop.  Policy: tag = load(i, o, d, tmin, tmax) reads ray i; commit(i, tag, found, hit) stores its
// result (tag = whatever the policy wants back, e.g. the path slot).
//
// STAGE (single-level kernels): every lane also keeps one PREPARED ray (origin, 1/direction, shear rows, interval, ids:
// kStageWords words) in shared memory.  A lane whose ray finishes starts its prepared ray in the same iteration -- no
// lane waits for a refill quorum -- and the prepared slots are refilled sc.stage_lanes at a time, so the queue fetch,
// the global ray loads and the reciprocal / shear set-up run at least that wide.
// Slot layout: 20 words per lane, lane-major, moved as five 128-bit shared-memory accesses (the 80-byte lane stride maps
// the eight lanes of a quarter warp onto disjoint bank quads).  A ray hand-over is executed by one or two lanes at a
// time in most loop iterations, so its instruction count is paid almost per ray: 5 LDS.128 instead of 19 LDS.32.
constexpr int kStageWords = 20;
template <bool ANY, bool COUNT, bool SINGLE, bool STAGE, class Policy>
__device__ void trace_persistent(const SceneView& sc, Policy& pol, uint32_t count, uint32_t* ticket, uint32_t* overflow,
                                 unsigned long long* node_visits, unsigned long long* tri_tests,
                                 unsigned long long* lane_stats = nullptr, uint32_t* stage_words = nullptr) {
  static_assert(!STAGE || SINGLE, "prepared rays are implemented for the single-level kernels");
  const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
  uint4* const stg = STAGE ? reinterpret_cast<uint4*>(stage_words) + ((threadIdx.x >> 5) * 32 + lane) * (kStageWords / 4) : nullptr;
  bool staged = false;
  Lane L;
  uint2 stack_local[kStackSize + 1];
  StackMem stack;
  stack.local = stack_local;
#if ASUNA_SMEM_STACK > 0
  __shared__ uint2 stack_shared[ASUNA_SMEM_STACK * kTraceThreads];
  stack.shared = stack_shared + threadIdx.x;
#endif
  L.sp = 0, L.blas_sp = 0, L.in_blas = false, L.shear_ok = false;
  L.ng = L.tg = L.top = make_uint2(0u, 0u);
  bool active = false, exhausted = (count == 0) || sc.n_instances == 0;
  uint32_t ray = 0, tag = 0, n_nodes = 0, n_tris = 0;
  uint32_t ls_iter = 0, ls_act = 0, ls_node = 0, ls_want = 0, ls_fired = 0, ls_tri = 0;  // COUNT only (warp-uniform)
  if (sc.n_instances == 0) {  // nothing to hit: every ray misses
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x); i < count; i += gridDim.x * blockDim.x) {
      float3 o, d;
      float t0, t1;
      const uint32_t g = pol.load(i, o, d, t0, t1);
      lane_begin<SINGLE>(L, sc, o, d, t0, t1);
      pol.commit(i, g, false, L.best, kKindUnknown);
    }
    return;
  }
  // work fetch: the warp owns the index range [w_next, w_end) of the queue and takes a new chunk with one atomic
  // when it runs dry; chunks shrink for short queues so that every warp of the grid still gets work
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t chunk = min((uint32_t)ASUNA_TICKET_CHUNK, max(count / (n_warps * 4u), 1u));
  uint32_t w_next = 0, w_end = 0;
  // start the prepared ray of this lane (STAGE)
  auto take_staged = [&]() {
    const uint4 a = stg[0], b = stg[1], c = stg[2], d = stg[3], e = stg[4];
    L.rs.o = f3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
    L.rs.idir = f3(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
    L.rs.sx = f3(__uint_as_float(b.z), __uint_as_float(b.w), __uint_as_float(c.x));
    L.rs.sy = f3(__uint_as_float(c.y), __uint_as_float(c.z), __uint_as_float(c.w));
    L.rs.sz = f3(__uint_as_float(d.x), __uint_as_float(d.y), __uint_as_float(d.z));
    L.rs.oct = (a.w >> 31) | ((b.x >> 31) << 1) | ((b.y >> 31) << 2);
    L.tmin = __uint_as_float(d.w), L.tmax = __uint_as_float(e.x);
    ray = e.y, tag = e.z;
    L.sp = 0;
    L.ng = make_uint2(sc.single_root, 0x80000000u);
    L.tg = make_uint2(0u, 0u);
    L.best.inst = 0xFFFFFFFFu, L.best.prim = 0xFFFFFFFFu, L.best.b1 = L.best.b2 = L.best.t = 0.f;
    staged = false;
    active = true;
  };
  for (;;) {
    // ---- refill from the queue: idle lanes directly, or (STAGE) the empty prepared-ray slots
    const uint32_t idle = __ballot_sync(0xFFFFFFFFu, STAGE ? !staged : !active);
    if (!exhausted && idle) {
      const uint32_t n_idle = __popc(idle), avail = w_end - w_next;
      uint32_t i = w_next + __popc(idle & lt);
      if (avail < n_idle) {  // warp-uniform: finish the old chunk, continue in a fresh one
        const uint32_t take = max(chunk, n_idle);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(ticket, take);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (i >= w_end) i = base + (i - w_end);
        w_next = base + (n_idle - avail), w_end = base + take;
      } else {
        w_next += n_idle;
      }
      if (STAGE) {
        if (!staged && i < count) {
          float3 o, d;
          float t0, t1;
          const uint32_t g = pol.load(i, o, d, t0, t1);
          RaySpace rs;
          setup_space(rs, o, d);
          setup_shear(rs, d);
          stg[0] = make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(rs.idir.x));
          stg[1] = make_uint4(__float_as_uint(rs.idir.y), __float_as_uint(rs.idir.z), __float_as_uint(rs.sx.x), __float_as_uint(rs.sx.y));
          stg[2] = make_uint4(__float_as_uint(rs.sx.z), __float_as_uint(rs.sy.x), __float_as_uint(rs.sy.y), __float_as_uint(rs.sy.z));
          stg[3] = make_uint4(__float_as_uint(rs.sz.x), __float_as_uint(rs.sz.y), __float_as_uint(rs.sz.z), __float_as_uint(t0));
          stg[4] = make_uint4(__float_as_uint(t1), i, g, 0u);
          staged = true;
        }
      } else if (!active && i < count) {
        float3 o, d;
        float t0, t1;
        tag = pol.load(i, o, d, t0, t1);
        lane_begin<SINGLE>(L, sc, o, d, t0, t1);
        ray = i;
        active = true;
      }
      if (w_next >= count) exhausted = true;
    }
    if (STAGE && !active && staged) take_staged();
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) {
      if (!STAGE || (exhausted && __ballot_sync(0xFFFFFFFFu, staged) == 0u)) break;
      continue;
    }
    // ---- traverse until enough lanes have finished to make a refill worthwhile
    for (;;) {
      // pop / leave instance / terminate; lanes holding only a postponed triangle group swap it under the next
      // node group of their own mesh-level stack and keep traversing
      if (active && L.ng.y <= 0x00FFFFFFu) {
        if (L.tg.y == 0u) {
          if (!SINGLE && L.in_blas && L.sp == L.blas_sp) {
            L.in_blas = false;
            setup_space(L.rs, L.wo, L.wd);
          }
          if (L.sp == 0) {
            const bool found = L.best.inst != 0xFFFFFFFFu;
            uint32_t kind = kKindUnknown;
            if (SINGLE && found) kind = L.best.inst >> 28, L.best.inst &= kInstMask;
            pol.commit(ray, tag, found, L.best, kind);
            active = false;
            if (STAGE && staged) take_staged();
          } else {
            const uint2 e = lane_pop(L, stack);
            if (e.y > 0x00FFFFFFu) L.ng = e;
            else L.tg = e;
          }
        } else if ((SINGLE || L.in_blas) && L.sp > (SINGLE ? 0 : L.blas_sp) && L.top.y > 0x00FFFFFFu) {
          L.ng = L.top;
          L.top = L.tg;
          L.tg = make_uint2(0u, 0u);
        }
      }
      const uint32_t m_act = __ballot_sync(0xFFFFFFFFu, active);
      if (m_act == 0u) break;
      if (COUNT) ls_iter++, ls_act += __popc(m_act), ls_node += __popc(__ballot_sync(0xFFFFFFFFu, active && L.ng.y > 0x00FFFFFFu));
      if (active && L.ng.y > 0x00FFFFFFu) lane_node_step<COUNT, SINGLE>(L, stack, sc, overflow, n_nodes);
      if (!SINGLE && active && !L.in_blas && L.tg.y) lane_enter_instance(L, stack, sc, overflow);
      const bool want_tri = active && (SINGLE || L.in_blas) && L.tg.y != 0u;
      const uint32_t m_tri = __ballot_sync(0xFFFFFFFFu, want_tri);
      if (COUNT) ls_want += __popc(m_tri);
      if (m_tri) {
        const uint32_t m_node = __ballot_sync(0xFFFFFFFFu, active && L.ng.y > 0x00FFFFFFu);
        if (m_node == 0u || __popc(m_tri) >= (__popc(m_act) >> sc.tri_vote_shift)) {
          if (COUNT) ls_fired++, ls_tri += __popc(m_tri);
          if (want_tri && lane_triangle_step<ANY, COUNT, SINGLE>(L, sc, n_tris)) {
            if (SINGLE) L.best.inst &= kInstMask;
            pol.commit(ray, tag, true, L.best, kKindUnknown);
            active = false;
            if (STAGE && staged) take_staged();
          }
        }
      }
      if (STAGE) {
        if (!exhausted && (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, !staged)) >= sc.stage_lanes) break;
      } else if (!exhausted && 32u - __popc(m_act) >= sc.refill_lanes) {
        break;
      }
    }
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      n_nodes += __shfl_xor_sync(0xFFFFFFFFu, n_nodes, o);
      n_tris += __shfl_xor_sync(0xFFFFFFFFu, n_tris, o);
    }
    if (lane == 0) {
      atomicAdd(node_visits, (unsigned long long)n_nodes);
      atomicAdd(tri_tests, (unsigned long long)n_tris);
      if (lane_stats) {
        atomicAdd(lane_stats + 0, (unsigned long long)ls_iter), atomicAdd(lane_stats + 1, (unsigned long long)ls_act);
        atomicAdd(lane_stats + 2, (unsigned long long)ls_node), atomicAdd(lane_stats + 3, (unsigned long long)ls_want);
        atomicAdd(lane_stats + 4, (unsigned long long)ls_fired), atomicAdd(lane_stats + 5, (unsigned long long)ls_tri);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// EXPERIMENT (ASUNA_TRI_POOL, single-level closest-hit kernel only): triangle tests decoupled from the lane that found
// them.  A node step's leaf hits are appended to a per-warp pool of (owner lane, triangle slot) items in shared memory;
// once 32 are pooled (or nobody has node work left) the whole warp tests one triangle per lane with the OWNER's ray
// constants fetched by indexed shuffles, and the per-ray minima are merged with shared-memory atomics under the same
// tie-break as the per-lane form (lowest t, then instance, then primitive).  Lanes never hold triangle groups, so the
// postponing / swapping logic disappears.  Measured against the per-lane form in profiles/README.md.
#ifndef ASUNA_TRI_POOL
#define ASUNA_TRI_POOL 0
#endif
constexpr int kPoolCap = 128;  // items per warp, power of two (ring buffer)
#ifndef ASUNA_POOL_FLUSH_MIN
#define ASUNA_POOL_FLUSH_MIN 16  // a batch also runs with this many items when some lane waits for its triangles
#endif
struct PoolMem {  // per warp
  uint32_t item[kPoolCap];
  unsigned long long best_key[32];
  uint32_t best_t[32];
  float best_b1[32], best_b2[32];
  uint32_t done[32];
};
constexpr int kPoolWordsPerWarp = sizeof(PoolMem) / 4;

template <bool COUNT, class Policy>
__device__ void trace_persistent_pool(const SceneView& sc, Policy& pol, uint32_t count, uint32_t* ticket, uint32_t* overflow,
                                      unsigned long long* node_visits, unsigned long long* tri_tests,
                                      uint32_t* stage_words, uint32_t* pool_words) {
  const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u, warp = threadIdx.x >> 5;
  uint4* const stg = reinterpret_cast<uint4*>(stage_words) + (warp * 32 + lane) * (kStageWords / 4);
  PoolMem& P = *reinterpret_cast<PoolMem*>(pool_words + warp * kPoolWordsPerWarp);
  bool staged = false;
  Lane L;
  uint2 stack_local[kStackSize + 1];
  StackMem stack;
  stack.local = stack_local;
#if ASUNA_SMEM_STACK > 0
  __shared__ uint2 stack_shared_p[ASUNA_SMEM_STACK * kTraceThreads];
  stack.shared = stack_shared_p + threadIdx.x;
#endif
  L.sp = 0, L.blas_sp = 0, L.in_blas = true, L.shear_ok = true;
  L.ng = L.tg = L.top = make_uint2(0u, 0u);
  bool active = false, exhausted = (count == 0);
  uint32_t ray = 0, tag = 0, n_nodes = 0, n_tris = 0, pending = 0;
  uint32_t p_head = 0, p_count = 0;  // warp-uniform ring state
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t chunk = min((uint32_t)ASUNA_TICKET_CHUNK, max(count / (n_warps * 4u), 1u));
  uint32_t w_next = 0, w_end = 0;
  auto take_staged = [&]() {
    const uint4 a = stg[0], b = stg[1], c = stg[2], d = stg[3], e = stg[4];
    L.rs.o = f3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
    L.rs.idir = f3(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
    L.rs.sx = f3(__uint_as_float(b.z), __uint_as_float(b.w), __uint_as_float(c.x));
    L.rs.sy = f3(__uint_as_float(c.y), __uint_as_float(c.z), __uint_as_float(c.w));
    L.rs.sz = f3(__uint_as_float(d.x), __uint_as_float(d.y), __uint_as_float(d.z));
    L.rs.oct = (a.w >> 31) | ((b.x >> 31) << 1) | ((b.y >> 31) << 2);
    L.tmin = __uint_as_float(d.w), L.tmax = __uint_as_float(e.x);
    ray = e.y, tag = e.z;
    L.sp = 0;
    L.ng = make_uint2(sc.single_root, 0x80000000u);
    L.tg = make_uint2(0u, 0u);
    L.best.inst = 0xFFFFFFFFu, L.best.prim = 0xFFFFFFFFu, L.best.b1 = L.best.b2 = L.best.t = 0.f;
    pending = 0;
    staged = false;
    active = true;
  };
  for (;;) {
    // ---- refill the empty prepared-ray slots from the queue
    const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !staged);
    if (!exhausted && idle) {
      const uint32_t n_idle = __popc(idle), avail = w_end - w_next;
      uint32_t i = w_next + __popc(idle & lt);
      if (avail < n_idle) {
        const uint32_t take = max(chunk, n_idle);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(ticket, take);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (i >= w_end) i = base + (i - w_end);
        w_next = base + (n_idle - avail), w_end = base + take;
      } else {
        w_next += n_idle;
      }
      if (!staged && i < count) {
        float3 o, d;
        float t0, t1;
        const uint32_t g = pol.load(i, o, d, t0, t1);
        RaySpace rs;
        setup_space(rs, o, d);
        setup_shear(rs, d);
        stg[0] = make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(rs.idir.x));
        stg[1] = make_uint4(__float_as_uint(rs.idir.y), __float_as_uint(rs.idir.z), __float_as_uint(rs.sx.x), __float_as_uint(rs.sx.y));
        stg[2] = make_uint4(__float_as_uint(rs.sx.z), __float_as_uint(rs.sy.x), __float_as_uint(rs.sy.y), __float_as_uint(rs.sy.z));
        stg[3] = make_uint4(__float_as_uint(rs.sz.x), __float_as_uint(rs.sz.y), __float_as_uint(rs.sz.z), __float_as_uint(t0));
        stg[4] = make_uint4(__float_as_uint(t1), i, g, 0u);
        staged = true;
      }
      if (w_next >= count) exhausted = true;
    }
    if (!active && staged) take_staged();
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) {
      if (exhausted && __ballot_sync(0xFFFFFFFFu, staged) == 0u) break;
      continue;
    }
    for (;;) {
      // ---- pop / finish: a ray is done when it has no node group, an empty stack, nothing held back and no pooled triangle
      if (active && L.ng.y <= 0x00FFFFFFu && L.tg.y == 0u) {
        if (L.sp > 0) {
          L.ng = lane_pop(L, stack);
        } else if (pending == 0u) {
          const bool found = L.best.inst != 0xFFFFFFFFu;
          uint32_t kind = kKindUnknown;
          if (found) kind = L.best.inst >> 28, L.best.inst &= kInstMask;
          pol.commit(ray, tag, found, L.best, kind);
          active = false;
          if (staged) take_staged();
        }
      }
      const uint32_t m_act = __ballot_sync(0xFFFFFFFFu, active);
      if (m_act == 0u) break;
      // ---- one wide-node step for every lane that has a node group (and is not holding triangles back)
      uint32_t leaf_base = L.tg.x, leaf_mask = L.tg.y;  // held back from an iteration in which the pool was full
      const bool stepping = active && L.ng.y > 0x00FFFFFFu && L.tg.y == 0u;
      if (stepping) {
        const uint32_t hits = L.ng.y;
        const uint32_t bit = 31u - (uint32_t)__clz(hits);
        const uint2 rest = make_uint2(L.ng.x, hits & ~(1u << bit));
        const uint32_t slot = (bit - 24u) ^ 7u ^ L.rs.oct;
        const uint32_t rel = __popc(hits & 0xFFu & ~(0xFFFFFFFFu << slot));
        const uint4* np = reinterpret_cast<const uint4*>(sc.blas_nodes + L.ng.x + rel);
        const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
        if (COUNT) n_nodes++;
        if (rest.y > 0x00FFFFFFu) lane_push(L, stack, rest, overflow);
        const uint32_t mask = intersect_wide_node(n0, n1, n2, n3, n4, L.rs, L.tmin, L.tmax, sc.magic);
        L.ng = make_uint2(n1.x, (mask & 0xFF000000u) | (n0.w >> 24));
        leaf_base = n1.y, leaf_mask = mask & 0x00FFFFFFu;
      }
      // ---- append the leaf hits to the warp's pool (whole lanes, in lane order, while there is room)
      if (__ballot_sync(0xFFFFFFFFu, leaf_mask != 0u)) {
        const uint32_t c = __popc(leaf_mask);
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
          if (lane >= (uint32_t)o) x += y;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, x, 31);
        if (total) {
          const uint32_t room = (uint32_t)kPoolCap - p_count;
          const bool fits = x <= room;  // monotone in the lane index: the lanes that fit form a prefix
          if (fits) {
            uint32_t pos = p_head + p_count + (x - c), m = leaf_mask;
            while (m) {
              const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
              m &= m - 1u;
              P.item[pos++ & (kPoolCap - 1)] = (lane << 27) | (leaf_base + b);
            }
            pending += c;
            L.tg = make_uint2(0u, 0u);
          } else {
            L.tg = make_uint2(leaf_base, leaf_mask);
          }
          const uint32_t fit_mask = __ballot_sync(0xFFFFFFFFu, fits);
          const uint32_t added = fit_mask ? __shfl_sync(0xFFFFFFFFu, x, 31 - __clz(fit_mask)) : 0u;
          p_count += added;
          __syncwarp();
        }
      }
      // ---- pooled triangle step
      {
        const uint32_t m_node = __ballot_sync(0xFFFFFFFFu, active && L.ng.y > 0x00FFFFFFu && L.tg.y == 0u);
        const uint32_t m_wait = __ballot_sync(0xFFFFFFFFu, active && (L.tg.y != 0u || (L.ng.y <= 0x00FFFFFFu && L.sp == 0 && pending != 0u)));
        while (p_count >= 32u || (p_count > 0u && (m_node == 0u || (m_wait != 0u && p_count >= (uint32_t)ASUNA_POOL_FLUSH_MIN)))) {
          const uint32_t nb = min(p_count, 32u);
          const bool have = lane < nb;
          const uint32_t it = have ? P.item[(p_head + lane) & (kPoolCap - 1)] : 0u;
          const uint32_t src = have ? it >> 27 : lane;
          const TriSlot* tp = sc.tris + (it & 0x07FFFFFFu);
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0;
          if (have) v0 = __ldg(&tp->v0), v1 = __ldg(&tp->v1), v2 = __ldg(&tp->v2);
          RaySpace rs;
          rs.o = f3(__shfl_sync(0xFFFFFFFFu, L.rs.o.x, src), __shfl_sync(0xFFFFFFFFu, L.rs.o.y, src), __shfl_sync(0xFFFFFFFFu, L.rs.o.z, src));
          rs.sx = f3(__shfl_sync(0xFFFFFFFFu, L.rs.sx.x, src), __shfl_sync(0xFFFFFFFFu, L.rs.sx.y, src), __shfl_sync(0xFFFFFFFFu, L.rs.sx.z, src));
          rs.sy = f3(__shfl_sync(0xFFFFFFFFu, L.rs.sy.x, src), __shfl_sync(0xFFFFFFFFu, L.rs.sy.y, src), __shfl_sync(0xFFFFFFFFu, L.rs.sy.z, src));
          rs.sz = f3(__shfl_sync(0xFFFFFFFFu, L.rs.sz.x, src), __shfl_sync(0xFFFFFFFFu, L.rs.sz.y, src), __shfl_sync(0xFFFFFFFFu, L.rs.sz.z, src));
          const float src_tmin = __shfl_sync(0xFFFFFFFFu, L.tmin, src), src_tmax = __shfl_sync(0xFFFFFFFFu, L.tmax, src);
          // owners publish their current best; candidates compete in shared memory
          const unsigned long long old_key = L.best.inst != 0xFFFFFFFFu
                                                 ? ((unsigned long long)(L.best.inst & kInstMask) << 32) | (L.best.inst & ~kInstMask) | L.best.prim
                                                 : ~0ull;
          const uint32_t old_t = __float_as_uint(L.tmax);
          P.best_t[lane] = old_t, P.best_key[lane] = old_key, P.done[lane] = 0u;
          __syncwarp();
          float t = 0.f, b1 = 0.f, b2 = 0.f;
          bool cand = false;
          unsigned long long key = ~0ull;
          if (have) {
            if (COUNT) n_tris++;
            atomicAdd(&P.done[src], 1u);
            cand = hit_triangle(rs, f3(v0), f3(v1), f3(v2), t, b1, b2) && t > src_tmin && t <= src_tmax;
            if (cand) {
              const uint32_t pk = __float_as_uint(v1.w);  // kind << 28 | instance
              key = ((unsigned long long)(pk & kInstMask) << 32) | (pk & ~kInstMask) | __float_as_uint(v0.w);
              atomicMin(&P.best_t[src], __float_as_uint(t));
            }
          }
          __syncwarp();
          if (P.best_t[lane] < old_t) P.best_key[lane] = ~0ull;  // strictly nearer: the old hit no longer competes
          __syncwarp();
          cand = cand && __float_as_uint(t) == P.best_t[src];
          if (cand) atomicMin(&P.best_key[src], key);
          __syncwarp();
          if (cand && key == P.best_key[src]) P.best_b1[src] = b1, P.best_b2[src] = b2;
          __syncwarp();
          const unsigned long long new_key = P.best_key[lane];
          if (new_key != old_key) {
            L.tmax = L.best.t = __uint_as_float(P.best_t[lane]);
            L.best.b1 = P.best_b1[lane], L.best.b2 = P.best_b2[lane];
            L.best.inst = (uint32_t)(new_key >> 32) | ((uint32_t)new_key & ~kInstMask);
            L.best.prim = (uint32_t)new_key & 0x07FFFFFFu;
          }
          pending -= P.done[lane];
          __syncwarp();
          p_head = (p_head + nb) & (kPoolCap - 1), p_count -= nb;
        }
      }
      if (!exhausted && (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, !staged)) >= sc.stage_lanes) break;
    }
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      n_nodes += __shfl_xor_sync(0xFFFFFFFFu, n_nodes, o);
      n_tris += __shfl_xor_sync(0xFFFFFFFFu, n_tris, o);
    }
    if (lane == 0) {
      atomicAdd(node_visits, (unsigned long long)n_nodes);
      atomicAdd(tri_tests, (unsigned long long)n_tris);
    }
  }
}

}  // namespace asuna
