// Launch interface of the sm_100a BVH builder (bvh_build.cu).
#pragma once
#include "device_types.cuh"

namespace asuna {

// Device temporaries of one build, sized for the largest primitive count seen so far.
struct BuildScratch {
  float4 *blo = nullptr, *bhi = nullptr;   // primitive boxes
  uint64_t* keys[2] = {nullptr, nullptr};  // Morton keys (ping-pong)
  uint32_t* vals[2] = {nullptr, nullptr};  // primitive ids (ping-pong)
  uint32_t* hist = nullptr;                // [256][sort blocks]
  int2* children = nullptr;                // binary hierarchy: children of inner node i
  float4 *nlo = nullptr, *nhi = nullptr;   // [2n-1] binary-node boxes (lo.w = primitive count)
  float* cost = nullptr;                   // [2n-1][7] SAH collapse table
  uint64_t* dec = nullptr;                 // [2n-1] its argmin bookkeeping
  int* cid[2] = {nullptr, nullptr};        // PLOC cluster arrays (ping-pong)
  float4* clo[2] = {nullptr, nullptr};
  float4* chi[2] = {nullptr, nullptr};
  int* nn = nullptr;
  uint2* pre = nullptr;
  uint2* block_sums = nullptr;
  int* root_of = nullptr;                  // binary subtree root of each wide node
  uint32_t* counters = nullptr;
  int* bounds = nullptr;                   // ordered-int centroid bounds (6 words)
  void* arena = nullptr;                   // all of the above live in ONE device allocation (a build pays one cudaMalloc)
  uint32_t capacity = 0;
  uint32_t coop_blocks = 0;                // co-resident grid size of k_ploc ...
  uint32_t coop_blocks_emit = 0;           // ... and of k_emit_wide<4> / <2>
  uint32_t coop_blocks_emit_small = 0;
  ~BuildScratch();
  cudaError_t reserve(uint32_t n);
  void release();
};

// What the primitive slots of a BVH hold: triangles of one mesh (tris != nullptr) or instance ids.
struct PrimPayload {
  const AsunaVertex* vertices = nullptr;
  const uint32_t* indices = nullptr;
  const TriSlot* soup = nullptr;       // world-space triangles (merged BLAS): primitive i = soup[i], vertices/indices unused
  const uint32_t* prim_ids = nullptr;  // top level: primitive -> instance index (nullptr = identity)
  TriSlot* tris = nullptr;             // absolute array
  uint32_t* leaf_inst = nullptr;       // absolute array
};

struct BuildResult {  // written by the emit kernel, read back once after all builds
  uint32_t wide_nodes;
  uint32_t prim_slots;
  float sah_cost;  // C(root, 1) / area(root), c_node = 1
  uint32_t pad;
};

// One instance of the merged world-space BLAS: `n` triangles of a mesh taken through o2w into soup[offset ...).
struct WorldJob {
  const AsunaVertex* v;
  const uint32_t* idx;
  float4 r0, r1, r2;  // object -> world rows
  uint32_t n, offset, inst;
  uint32_t kind;  // HitKind of the instance (emitter / material type): stored above the instance id in the slot's v1.w
};
void launch_world_triangles_batched(cudaStream_t s, const WorldJob* d_jobs, uint32_t n_jobs, uint32_t max_n, TriSlot* soup,
                                    BuildScratch& sc);
void launch_tri_boxes(cudaStream_t s, const AsunaVertex* v, const uint32_t* idx, uint32_t n, BuildScratch& sc);
void launch_instance_boxes(cudaStream_t s, const DInstance* inst, const uint32_t* ids, const float4* mesh_lo,
                           const float4* mesh_hi, uint32_t n, BuildScratch& sc);
cudaError_t launch_build_wide(cudaStream_t s, uint32_t n, WideNode* nodes, uint32_t node_base, uint32_t prim_base,
                              BuildScratch& sc, const PrimPayload& payload, float cost_prim, float4* root_lo,
                              float4* root_hi, BuildResult* result);
cudaError_t radix_sort_pairs(cudaStream_t s, uint64_t* keys_io, uint32_t* vals_io, uint32_t n, BuildScratch& sc);

}  // namespace asuna
