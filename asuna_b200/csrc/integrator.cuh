// Launch interface of the wavefront integrator (integrator.cu).
#pragma once
#include "device_types.cuh"

namespace asuna {

struct OutputImages {
  float4* img[ASUNA_NUM_OUTPUT_IMAGES];  // 0 radiance mean, 1..7 AOVs, 8 filter-weight sum
};

struct LaunchDims {
  uint32_t trace_blocks = 0;  // persistent grids: SM count x resident blocks per SM
  uint32_t trace_blocks_single = 0;  // same for the single-level kernels (fewer registers, one more block per SM)
  uint32_t shadow_blocks_single = 0;  // ... and for the single-level any-hit kernel (one more again)
  uint32_t shade_blocks[kNumKinds] = {};  // per hit kind (register use differs per material)
};

cudaError_t query_launch_dims(LaunchDims& ld, int sm_count);

void launch_raygen(cudaStream_t s, const FrameParams& fp, const PathState& ps, const OutputImages& out, Counters* cnt);
void launch_trace_closest(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const PathState& ps, Counters* cnt,
                          int iter, int qsel, bool counting);
void launch_fold_counters(cudaStream_t s, const Counters* cnt, Totals* tot, int iters);
void launch_sky_prepare(cudaStream_t s, const AsunaSunSky& ss, SkyPre* out_device);
int launch_shade(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const FrameParams& fp, const PathState& ps,
                 const OutputImages& out, Counters* cnt, int iter, int qsel, uint32_t kind_mask, uint32_t n_paths);
void launch_trace_shadow(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const PathState& ps, Counters* cnt,
                         int iter);
void launch_accumulate(cudaStream_t s, const FrameParams& fp, const PathState& ps, const OutputImages& out);
void launch_export_partial(cudaStream_t s, const OutputImages& out, float4* partial, uint32_t n, int have_accum);
void launch_import_partial(cudaStream_t s, const OutputImages& out, const float4* partial, uint32_t n);
// block_sums: 3 x 296 doubles of scratch (only touched for the custom tone mapper with auto-exposure)
void launch_post_process(cudaStream_t s, const float4* hdr, float4* ldr, uint32_t w, uint32_t h, const AsunaPost& tm, double* block_sums);
void launch_primary_rays(cudaStream_t s, const FrameParams& fp, float4* rays);
void launch_trace_user(cudaStream_t s, const LaunchDims& ld, const SceneView& sc, const float4* rays, uint32_t n, float* tuv,
                       uint32_t* inst_prim, uint8_t* occluded, Counters* cnt);

}  // namespace asuna
