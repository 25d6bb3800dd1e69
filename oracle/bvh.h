// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// The reference delegates acceleration-structure build and traversal to the Vulkan driver
// (vkCmdBuildAccelerationStructuresKHR, reference ext/nvpro_core/nvvk/raytraceKHR_vk.cpp:216,375;
// traceRayEXT, reference src/shaders/raytrace.projective.rgen:108,121).  What that black box
// *returns* is mathematically defined: the nearest triangle hit of a two-level (instance ->
// mesh) scene with rays transformed into object space and t preserved.  This file is a
// deliberately plain CPU implementation of that definition: a binned-SAH BVH2 per mesh, one
// over the instance boxes, a watertight ray/triangle test (Woop, Benthin, Wald 2013) and the
// tie-break the driver leaves unspecified (lowest t, then lowest instance, then lowest
// primitive -- SURVEY.md section 7 "hard parts" item 3).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <vector>

#include "vecmath.h"

namespace orc {

struct Box {
  vec3 lo{FLT_MAX, FLT_MAX, FLT_MAX}, hi{-FLT_MAX, -FLT_MAX, -FLT_MAX};
  void grow(vec3 p) {
    lo = {std::fmin(lo.x, p.x), std::fmin(lo.y, p.y), std::fmin(lo.z, p.z)};
    hi = {std::fmax(hi.x, p.x), std::fmax(hi.y, p.y), std::fmax(hi.z, p.z)};
  }
  void grow(const Box& b) {
    grow(b.lo);
    grow(b.hi);
  }
  float half_area() const {
    vec3 e = hi - lo;
    return e.x * e.y + e.y * e.z + e.z * e.x;
  }
  vec3 centre() const { return (lo + hi) * 0.5f; }
};

struct BvhNode {
  Box box;
  uint32_t left = 0;   // inner: index of left child (right = left + 1); leaf: first prim slot
  uint32_t count = 0;  // 0 for inner nodes, else number of prims in the leaf
};

struct Bvh {
  std::vector<BvhNode> nodes;
  std::vector<uint32_t> prims;  // permutation of primitive ids

  // SAH cost of the best 8-wide BVH obtainable from this binary tree by the collapse of Ylitie, Karras, Laine 2017
  // section 4.1 (dynamic programme over "forest of at most i roots", leaves of at most pmax primitives):
  // C(root, 1) / area(root).  The yardstick for the GPU builder's tree (SAH <= 1.15 x CPU SAH).
  double wide_sah_cost(double c_node, double c_prim, uint32_t pmax) const {
    const size_t n = nodes.size();
    if (n == 0 || prims.empty()) return 0.0;
    const double INF = 1e300;
    std::vector<double> C(n * 8, INF);  // C[node * 8 + i], i = 1..7
    std::vector<uint32_t> P(n, 0);
    for (size_t k = n; k-- > 0;) {  // children are appended after their parent: descending order is post-order
      const BvhNode& nd = nodes[k];
      const double A = nd.box.half_area();
      if (nd.count) {
        P[k] = nd.count;
        for (int i = 1; i <= 7; i++) C[k * 8 + i] = A * nd.count * c_prim;
        continue;
      }
      const size_t l = nd.left, r = nd.left + 1;
      P[k] = P[l] + P[r];
      double D[9];
      for (int j = 2; j <= 8; j++) {
        D[j] = INF;
        for (int a = 1; a < j; a++) D[j] = std::min(D[j], C[l * 8 + std::min(a, 7)] + C[r * 8 + std::min(j - a, 7)]);
      }
      C[k * 8 + 1] = std::min(P[k] <= pmax ? A * P[k] * c_prim : INF, D[8] + A * c_node);
      for (int i = 2; i <= 7; i++) C[k * 8 + i] = std::min(D[i], C[k * 8 + i - 1]);
    }
    return C[1] / nodes[0].box.half_area();
  }

  void build(const std::vector<Box>& boxes) {
    const uint32_t n = (uint32_t)boxes.size();
    prims.resize(n);
    for (uint32_t i = 0; i < n; i++) prims[i] = i;
    nodes.clear();
    nodes.reserve(2 * n + 1);
    nodes.emplace_back();
    if (n == 0) return;
    std::vector<vec3> cent(n);
    for (uint32_t i = 0; i < n; i++) cent[i] = boxes[i].centre();
    split(0, 0, n, boxes, cent);
  }

 private:
  void split(uint32_t node, uint32_t first, uint32_t count, const std::vector<Box>& boxes,
             const std::vector<vec3>& cent) {
    Box nb, cb;
    for (uint32_t i = first; i < first + count; i++) {
      nb.grow(boxes[prims[i]]);
      cb.grow(cent[prims[i]]);
    }
    nodes[node].box = nb;
    if (count <= 2) {
      nodes[node].left = first;
      nodes[node].count = count;
      return;
    }
    constexpr int NB = 16;
    float best = FLT_MAX;
    int best_axis = -1, best_bin = 0;
    for (int a = 0; a < 3; a++) {
      float lo = cb.lo[a], ext = cb.hi[a] - lo;
      if (!(ext > 0.0f)) continue;
      Box bb[NB];
      uint32_t bc[NB] = {0};
      float scale = NB / ext;
      for (uint32_t i = first; i < first + count; i++) {
        int b = std::min(NB - 1, (int)((cent[prims[i]][a] - lo) * scale));
        bb[b].grow(boxes[prims[i]]);
        bc[b]++;
      }
      float right_area[NB];
      uint32_t right_cnt[NB];
      Box acc;
      uint32_t c = 0;
      for (int b = NB - 1; b > 0; b--) {
        if (bc[b]) acc.grow(bb[b]);
        c += bc[b];
        right_area[b] = c ? acc.half_area() : 0.0f;
        right_cnt[b] = c;
      }
      acc = Box();
      c = 0;
      for (int b = 0; b < NB - 1; b++) {
        if (bc[b]) acc.grow(bb[b]);
        c += bc[b];
        if (c == 0 || right_cnt[b + 1] == 0) continue;
        float cost = acc.half_area() * c + right_area[b + 1] * right_cnt[b + 1];
        if (cost < best) {
          best = cost;
          best_axis = a;
          best_bin = b;
        }
      }
    }
    uint32_t mid;
    if (best_axis < 0 || (count <= 4 && best >= nb.half_area() * count)) {
      if (count <= 4) {
        nodes[node].left = first;
        nodes[node].count = count;
        return;
      }
      // all centroids coincide: median split
      mid = first + count / 2;
    } else {
      float lo = cb.lo[best_axis], scale = NB / (cb.hi[best_axis] - lo);
      auto it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](uint32_t p) {
        int b = std::min(NB - 1, (int)((cent[p][best_axis] - lo) * scale));
        return b <= best_bin;
      });
      mid = (uint32_t)(it - prims.begin());
      if (mid == first || mid == first + count) mid = first + count / 2;
    }
    uint32_t l = (uint32_t)nodes.size();
    nodes.emplace_back();
    nodes.emplace_back();
    nodes[node].left = l;
    nodes[node].count = 0;
    split(l, first, mid - first, boxes, cent);
    split(l + 1, mid, first + count - mid, boxes, cent);
  }
};

// Slab test against [tmin, tmax]; the far plane is widened by 2 ulp so the box test never
// rejects a triangle the watertight triangle test would accept.
inline bool hit_box(const Box& b, vec3 o, vec3 inv_d, float tmin, float tmax, float& tnear) {
  float tx0 = (b.lo.x - o.x) * inv_d.x, tx1 = (b.hi.x - o.x) * inv_d.x;
  float ty0 = (b.lo.y - o.y) * inv_d.y, ty1 = (b.hi.y - o.y) * inv_d.y;
  float tz0 = (b.lo.z - o.z) * inv_d.z, tz1 = (b.hi.z - o.z) * inv_d.z;
  float tn = std::fmax(std::fmax(std::fmin(tx0, tx1), std::fmin(ty0, ty1)), std::fmax(std::fmin(tz0, tz1), tmin));
  float tf = std::fmin(std::fmin(std::fmax(tx0, tx1), std::fmax(ty0, ty1)), std::fmin(std::fmax(tz0, tz1), tmax));
  tnear = tn;
  return tn <= tf * 1.0000004f;
}

// Per-ray constants of the watertight test: dominant axis kz, kx = kz+1, ky = kz+2 (cyclic), shear
// Sx = d[kx]/d[kz], Sy = d[ky]/d[kz], Sz = 1/d[kz], kept as the three rows of the shear matrix
// (e_kx - Sx e_kz, e_ky - Sy e_kz, Sz e_kz) so the test needs no permutation.  The kx/ky swap of the paper is
// dropped: it negates all three edge functions together (exactly), which changes nothing without culling.
struct RayShear {
  vec3 sx, sy, sz;
  explicit RayShear(vec3 d) {
    float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    int kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    int kx = (kz + 1) % 3, ky = (kx + 1) % 3;
    float Sx = d[kx] / d[kz], Sy = d[ky] / d[kz], Sz = 1.0f / d[kz];
    float m[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    m[0][kx] = 1.f, m[0][kz] = -Sx;
    m[1][ky] = 1.f, m[1][kz] = -Sy;
    m[2][kz] = Sz;
    sx = vec3{m[0][0], m[0][1], m[0][2]};
    sy = vec3{m[1][0], m[1][1], m[1][2]};
    sz = vec3{m[2][0], m[2][1], m[2][2]};
  }
};

// fma(s.z, a.z, fma(s.y, a.y, s.x * a.x)): the exact operation order of the CUDA kernel (traverse.cuh dot_chain)
inline float dot_chain(vec3 s, vec3 a) { return std::fmaf(s.z, a.z, std::fmaf(s.y, a.y, s.x * a.x)); }

// Watertight ray/triangle test.  No back-face culling (the reference sets
// VK_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE, src/pipeline/pipeline_raytrace.cpp:132).
// Returns t and the Vulkan barycentrics (b1, b2): hit = (1-b1-b2) v0 + b1 v1 + b2 v2.
inline bool hit_triangle(vec3 o, const RayShear& rs, vec3 v0, vec3 v1, vec3 v2, float& t, float& b1, float& b2) {
  vec3 A = v0 - o, B = v1 - o, C = v2 - o;
  float Ax = dot_chain(rs.sx, A), Ay = dot_chain(rs.sy, A);
  float Bx = dot_chain(rs.sx, B), By = dot_chain(rs.sy, B);
  float Cx = dot_chain(rs.sx, C), Cy = dot_chain(rs.sy, C);
  float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  float det = U + V + W;
  if (det == 0.0f) return false;
  float Az = dot_chain(rs.sz, A), Bz = dot_chain(rs.sz, B), Cz = dot_chain(rs.sz, C);
  float T = U * Az + V * Bz + W * Cz;
  float inv = 1.0f / det;
  t = T * inv;
  b1 = V * inv;
  b2 = W * inv;
  return true;
}

}  // namespace orc
