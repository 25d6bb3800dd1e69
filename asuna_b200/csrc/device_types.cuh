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
// Compressed 8-wide BVH node (80 bytes = 5 x LDG.128), after Ylitie, Karras, Laine 2017
// "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs".  The eight child boxes are
// quantised to 8 bits per plane on a per-node grid: plane = p + q * 2^e (conservative: lo
// rounded down, hi rounded up, after a 2-ulp pad).  Children sit in octant-ordered slots so that
// visiting slot (s ^ ray octant) approximates front-to-back order without sorting.
//   meta[s] = 0                         empty slot
//   meta[s] = 001 | (24 + s)            inner child; its node index is child_base + popc(imask & below(s))
//   meta[s] = unary(count) | offset     leaf child: `count` (1..3) consecutive primitive slots from
//                                       prim_base + offset (offset < 24)
// The same node type serves both levels: in a mesh BVH (BLAS) primitive slots are TriSlot entries,
// in the instance BVH (TLAS) they index tlas_leaf_inst.
struct __align__(16) WideNode {
  float px, py, pz;      // grid origin
  uint8_t ex, ey, ez;    // biased exponents: scale = uint_as_float(e << 23)
  uint8_t imask;         // which slots hold inner children
  uint32_t child_base;   // absolute index of the first inner child
  uint32_t prim_base;    // absolute index of the first primitive slot
  uint8_t meta[8];
  uint8_t qlox[8], qloy[8], qloz[8];
  uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(WideNode) == 80, "compressed wide node is 80 bytes");
constexpr int kMaxLeafPrims = 3;

// Triangle slot in BVH-leaf order: three float4 (48 B), prim id in v0.w.
struct __align__(16) TriSlot {
  float4 v0;  // xyz, w = __uint_as_float(primitive id in the mesh)
  float4 v1;  // xyz, w = __uint_as_float(instance id) in the world BLAS, unused in a mesh BLAS
  float4 v2;
};

struct __align__(16) DInstance {
  float4 w2o[3];      // world->object rows (3x4)
  float4 o2w[3];      // object->world rows (3x4)
  int32_t blas_root;  // absolute wide-node index of the mesh BVH root
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
  const WideNode* tlas_nodes;     // top-level nodes (root 0); leaf slots index tlas_leaf_inst
  const uint32_t* tlas_leaf_inst;
  const WideNode* blas_nodes;     // all mesh BVHs, absolute indices
  const TriSlot* tris;            // all mesh triangles in leaf order, absolute indices
  const DInstance* instances;
  const DMesh* meshes;
  const AsunaMaterial* materials;
  const AsunaLight* lights;
  const DTexture* textures;
  DTexture env[3];                // env, marginal, conditional
  uint32_t n_instances;
  uint32_t world_inst;            // pseudo-instance of the merged world-space BLAS (hits take their instance id from the triangle), or ~0
  uint32_t single_root;           // every instance merged: root of the world BLAS, traversal is single-level; else ~0
  uint32_t magic;                 // 0x4B000000 from the constant bank: keeps the PRMT selectors immediate (traverse.cuh)
  uint32_t refill_lanes;          // traversal tuning (ASUNA_TUNE): refill when this many lanes are idle
  uint32_t tri_vote_shift;        // triangle step quorum = live lanes >> shift
  uint32_t stage_lanes;           // prepared-ray slots are refilled once this many are empty (single-level kernels)
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
  uint8_t* kind;     // per queue index: what the closest-hit ray found (HitKind), written by the trace kernel
  uint32_t* sorted;  // the current queue regrouped by kind (counting sort), consumed by the per-kind shade kernels
  uint32_t* bin_hist;  // [kNumKinds][bin blocks] counts, then exclusive offsets
};

// What a closest-hit ray found: the reference dispatches on instanceShaderBindingTableRecordOffset = material type
// (src/pipeline/pipeline_raytrace.cpp:134-140); here the hit queue is regrouped by this key instead.
enum HitKind : uint32_t { kKindMiss = 0, kKindLight = 1, kKindMaterial0 = 2 /* + AsunaMaterialType */, kNumKinds = 16 };
constexpr uint32_t kBinTile = 8192;  // queue entries per block of the binning kernels

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
  // warp-loop occupancy of the instrumented traversal: iterations, sum of live lanes, lanes in node steps, lanes
  // wanting a triangle step, triangle steps run, lanes in them
  unsigned long long lane_stats[6];
};

// Totals folded from `Counters` at the end of every batch by k_fold_counters (no host sync needed).
struct Totals {
  unsigned long long closest_rays, shadow_rays, incoherent_rays, node_visits, tri_tests;
  unsigned long long stack_overflow;
  unsigned long long lane_stats[6];
};

// Sun & sky model: the per-setting part of sun_and_sky (shading.cuh), evaluated on the device once per setting.
struct SkyPre {
  float rgb_scale[3], sat, horiz, haze, factor;
  float sun[3], real_sun[3], sun_color_up[3], sun_color_down[3];
  float sun_radius, disk_scale, glow_scale;
  float lum0, zx, zy, perez[3][5], perez_den[3];  // sky luminance / chromaticity coefficients for T = haze
  float ground_irrad[3];                           // calc_irrad
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
  SkyPre sky;                        // sun & sky: what depends on the setting only (shading.cuh sky_prepare)
};
#define ASUNA_MAX_BATCH_FRAMES 64

}  // namespace asuna

struct asuna_ctx;
void asuna_set_cuda_error(asuna_ctx* ctx, cudaError_t e, const char* expr, const char* file, int line);
