// Launch interface of the sm_100a BVH builder (bvh_build.cu).
#pragma once
#include "device_types.cuh"

namespace asuna {

// Device temporaries of one build, sized for the largest primitive count seen so far.
struct BuildScratch {
  float4 *blo = nullptr, *bhi = nullptr;  // primitive boxes
  uint64_t* keys[2] = {nullptr, nullptr};  // Morton keys (ping-pong)
  uint32_t* vals[2] = {nullptr, nullptr};  // primitive ids (ping-pong)
  uint32_t* hist = nullptr;                // [256][sort blocks]
  int2* children = nullptr;                // LBVH inner-node children (build numbering)
  int* parent = nullptr;                   // [2n-1]
  float4 *nlo = nullptr, *nhi = nullptr;   // [2n-1] fitted boxes
  uint32_t* flags = nullptr;               // arrival counters
  int* bounds = nullptr;                   // ordered-int centroid bounds (6 words)
  uint32_t capacity = 0;
  ~BuildScratch();
  cudaError_t reserve(uint32_t n);
  void release();
};

void launch_tri_boxes(cudaStream_t s, const AsunaVertex* v, const uint32_t* idx, uint32_t n, BuildScratch& sc);
void launch_instance_boxes(cudaStream_t s, const DInstance* inst, const float4* mesh_lo, const float4* mesh_hi,
                           uint32_t n, BuildScratch& sc);
void launch_lbvh(cudaStream_t s, uint32_t n, BvhNode* nodes, int node_base, int leaf_base, BuildScratch& sc,
                 float4* root_lo, float4* root_hi);
void launch_emit_tris(cudaStream_t s, const AsunaVertex* v, const uint32_t* idx, uint32_t n, const uint32_t* order,
                      TriSlot* tris);
void launch_sah_cost(cudaStream_t s, const BvhNode* nodes, int node_base, int n_nodes, double* cost_sum);
cudaError_t radix_sort_pairs(cudaStream_t s, uint64_t* keys_io, uint32_t* vals_io, uint32_t n, BuildScratch& sc);

}  // namespace asuna
