// TEST INFRASTRUCTURE ONLY.  Plain-C seam between the oracle's scene store / BVH (oracle.cpp built with
// -DASUNA_REF_SHADERS) and the translation unit generated from the reference's GLSL (build_ref.py).
// oracle.cpp answers traceRayEXT's geometric question (nearest hit / any hit, the part the Vulkan driver
// answers in the reference); every line of per-pixel arithmetic runs the reference's own shader text.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RefHit {
  float t, b1, b2;
  uint32_t inst, prim;
  float o2w[12], w2o[12];  // 3 rows x 4 columns, VkTransformMatrixKHR order
} RefHit;

typedef int (*ref_trace_fn)(void* ctx, const float o[3], const float d[3], float tmin, float tmax, int any,
                            RefHit* out);

typedef struct RefTex {
  const float* rgba;
  int32_t w, h;
} RefTex;

typedef struct RefInst {
  const void* vertices;     // GpuVertex[], 44 B stride (src/shared/vertex.h)
  const uint32_t* indices;  // 3 per triangle
  const void* material;     // GpuMaterial, 132 B (unused when light_id >= 0)
  int32_t light_id;
  uint32_t sbt_offset;  // material type; emitters use the lambertian hit group (pipeline_raytrace.cpp:134-140)
} RefInst;

typedef struct RefBind {
  void* ctx;
  ref_trace_fn trace;
  const void* camera;  // GpuCamera 224 B
  const void* sunsky;  // GpuSunAndSky 96 B
  const void* pc;      // GpuPushConstantRaytrace 84 B
  const void* lights;  // GpuLight[], index 0 = dummy
  uint32_t n_textures;
  const RefTex* textures;
  RefTex env[3];  // env map, marginal, conditional (binding order rchit_layouts.glsl:28)
  uint32_t n_instances;
  const RefInst* instances;
  float* images[9];
  uint32_t width, height;
} RefBind;

// Scene-constant state is copied into the generated unit's globals: call before the parallel pixel loop.
void refglsl_bind(const RefBind* b);
// One invocation of raytrace.projective.rgen main() for gl_LaunchIDEXT = (x, y, 0).
void refglsl_render_pixel(uint32_t x, uint32_t y);

// Probe of ONE shader invocation (tests/test_ref_pins.py): the payload as it enters a closest-hit or miss
// shader and as it leaves.  Layout shared by liboracle.so and libref.so (oracle_shade_probe).
typedef struct ShadeProbe {
  // in / out
  float ray_o[3], ray_d[3], radiance[3], throughput[3];
  uint32_t depth, seed, stop;
  float brec_d[3], brec_pdf;
  uint32_t brec_flags;
  // out only
  float drec_radiance[3], drec_dist, drec_o[3], drec_d[3];
  uint32_t drec_skip;
  float channel[8][3];
} ShadeProbe;
// hit == NULL runs raytrace.default.rmiss, otherwise the hit group of the instance.
void refglsl_shade_probe(const RefHit* hit, ShadeProbe* p);


// raytrace post stage: post.idle.frag over a w x h RGBA32F image; tm = GpuPushConstantPost (48 B)
void refglsl_post_process(const float* hdr, uint32_t w, uint32_t h, const void* tm48, float* out);

#ifdef __cplusplus
}
#endif
