// Device-side data layout of the B200 path-tracing core (see DESIGN.md "data layout in HBM").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/asuna_b200.h"

#define ASUNA_CUDA_CHECK(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      asuna_set_cuda_error(ctx, _e, #expr, __FILE__, __LINE__);                           \
      return ASUNA_E_CUDA;                                                                \
    }                                                                                     \
  } while (0)

namespace asuna {

// ---- acceleration structure -------------------------------------------------------------
// Binary BVH node in the Aila-Laine layout: the node carries the boxes of its two children so
// one 64-byte fetch (4 x LDG.128) decides both.  link >= 0: absolute index of an inner node;
// link < 0: ~link = (first leaf slot << 3) | (slot count - 1), 1..8 consecutive slots
// (0x80000000 is reserved as the traversal's leave-instance sentinel, so slots < 2^27).
struct __align__(16) BvhNode {
  float4 c0xy;   // child0: lo.x, hi.x, lo.y, hi.y
  float4 c1xy;   // child1: lo.x, hi.x, lo.y, hi.y
  float4 cz;     // child0 lo.z, hi.z, child1 lo.z, hi.z
  int4 link;     // child0, child1, count0, count1
};
static_assert(sizeof(BvhNode) == 64, "node is one 64-byte line half");

// Triangle slot in BVH-leaf order: three float4 (48 B), prim id in v0.w.
struct __align__(16) TriSlot {
  float4 v0;  // xyz, w = __uint_as_float(primitive id in the mesh)
  float4 v1;
  float4 v2;
};

struct __align__(16) DInstance {
  float4 w2o[3];      // world->object rows (3x4)
  float4 o2w[3];      // object->world rows (3x4)
  int32_t blas_root;  // absolute node index of the mesh BVH root
  uint32_t mesh;
  uint32_t material;
  int32_t light;      // >= 0: emitter instance
  uint32_t mat_type;  // AsunaMaterialType, or 0xFFFFFFFF for emitters
  uint32_t pad[3];
};
static_assert(sizeof(DInstance) == 128, "instance record is one 128-byte line");

struct DMesh {
  const AsunaVertex* vertices;
  const uint32_t* indices;
  uint32_t n_tris;
  uint32_t pad;
};

struct DTexture {
  const float4* texels;
  int32_t w, h;
};

// Everything a render kernel needs, passed by value (fits the 4 KB kernel-parameter space).
struct SceneView {
  const BvhNode* tlas_nodes;      // top-level nodes; leaves index tlas_leaf_inst
  const uint32_t* tlas_leaf_inst;
  const BvhNode* blas_nodes;      // all mesh BVHs, absolute indices
  const TriSlot* tris;            // all mesh triangles in leaf order, absolute indices
  const DInstance* instances;
  const DMesh* meshes;
  const AsunaMaterial* materials;
  const AsunaLight* lights;
  const DTexture* textures;
  DTexture env[3];                // env, marginal, conditional
  uint32_t n_instances;
};

// ---- wavefront path state (structure of float4 arrays, one slot per in-flight path) -------
struct PathState {
  float4* ray_o;   // o.xyz, w = RNG state (bits)
  float4* ray_d;   // d.xyz, w = pdf of the BSDF sample that produced the ray
  float4* thr;     // throughput.xyz, w = packed {depth:16, bsdf flags:8}
  float4* rad;     // radiance.xyz, w unused
  uint4* hit;      // b1 bits, b2 bits, instance, primitive   (instance 0xFFFFFFFF = miss)
  float4* sh_o;    // shadow queue: o.xyz, tmax
  float4* sh_d;    //               d.xyz, w = path slot (bits)
  float4* sh_l;    //               NEE radiance to add if unoccluded
  uint32_t* queue[2];
};

// Per-iteration device counters; one slot per bounce iteration so no reset kernel is needed.
#define ASUNA_MAX_ITERS 256
struct Counters {
  uint32_t queue[ASUNA_MAX_ITERS + 1];   // paths alive entering iteration i
  uint32_t shadow[ASUNA_MAX_ITERS + 1];  // shadow rays emitted by iteration i
  uint32_t incoherent[ASUNA_MAX_ITERS + 1];  // closest rays at depth >= 2 in iteration i
  uint32_t ticket_closest[ASUNA_MAX_ITERS + 1];  // dynamic work-fetch tickets of the persistent trace kernels
  uint32_t ticket_shadow[ASUNA_MAX_ITERS + 1];
  uint32_t stack_overflow;
  unsigned long long node_visits;   // instrumented traversal only (asuna_set_counting)
  unsigned long long tri_tests;
};

// Totals folded from `Counters` at the end of every batch by k_fold_counters (no host sync needed).
struct Totals {
  unsigned long long closest_rays, shadow_rays, incoherent_rays, node_visits, tri_tests;
  unsigned long long stack_overflow;
};

struct FrameParams {
  AsunaCamera cam;
  AsunaSunSky sunsky;
  AsunaState pc;
  uint32_t width, height;
  uint32_t n_pixels;
  uint32_t n_frames;                 // frames in this batch
  int32_t frame_ids[64];             // curFrame value of each batch frame
  uint32_t first_is_replace;         // first frame of the batch replaces the accumulation buffer
};
#define ASUNA_MAX_BATCH_FRAMES 64

}  // namespace asuna

struct asuna_ctx;
void asuna_set_cuda_error(asuna_ctx* ctx, cudaError_t e, const char* expr, const char* file, int line);
