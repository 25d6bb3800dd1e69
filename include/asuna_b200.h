/*
 * asuna_b200.h -- C ABI of the B200-native path-tracing core for the Asuna renderer.
 *
 * This is the drop-in boundary.  The reference has no FFI: the seam is the
 * PipelineRaytrace object (reference src/pipeline/pipeline_raytrace.h:15-50), whose
 * inputs are the flat "Gpu*" PODs of reference src/shared/ headers and whose outputs are the
 * nine RGBA32F images of src/shared/binding.h:70.  Every entry point below cites the
 * reference interface it replaces.  All structs are tightly packed 4-byte words
 * ("scalar layout"), little-endian, matrices column-major exactly like nvmath::mat4f.
 *
 * Conventions: every call returns 0 on success or a negative ASUNA_E_* code (the
 * reference calls exit(1) instead, e.g. src/scene/scene.cpp:324-327).  The library owns
 * all device memory; host pointers are only borrowed for the duration of a call.  One
 * context drives one GPU; a context is not re-entrant.  Multi-GPU runs use one context
 * per GPU (one process per GPU) with asuna_set_partition + asuna_export_partial /
 * asuna_import_partial around the caller's NCCL reduce (bench.py).
 */
#ifndef ASUNA_B200_H
#define ASUNA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- wire structs (reference src/shared/ headers; sizes checked in asuna_abi_check) ---- */

/* reference src/shared/vertex.h:6-11 -- 44 bytes */
typedef struct AsunaVertex {
  float pos[3];
  float uv[2];
  float normal[3];
  float tangent[3];
} AsunaVertex;

/* reference src/shared/material.h:7-21 */
enum AsunaMaterialType {
  ASUNA_MAT_LAMBERTIAN = 0,
  ASUNA_MAT_KANG18 = 1,
  ASUNA_MAT_EMISSIVE = 2,
  ASUNA_MAT_PBR_METALNESS_ROUGHNESS = 3,
  ASUNA_MAT_PLASTIC = 4,
  ASUNA_MAT_ROUGH_PLASTIC = 5,
  ASUNA_MAT_CONDUCTOR = 6,
  ASUNA_MAT_ROUGH_CONDUCTOR = 7,
  ASUNA_MAT_MIRROR = 8,
  ASUNA_MAT_DISNEY = 9,
  ASUNA_MAT_DIELECTRIC = 10,
  ASUNA_MAT_PHONG = 11,
  ASUNA_MAT_NUM = 12
};

/* reference src/shared/material.h:24-49 -- 132 bytes */
typedef struct AsunaMaterial {
  float diffuse[3];
  float rhoSpec[3];
  float anisoAlpha[2];
  float ior;
  float roughness;
  float subsurface;
  float specular;
  float specularTint;
  float anisotropic;
  float sheen;
  float sheenTint;
  float clearcoat;
  float clearcoatGloss;
  float radiance[3];
  float metalness;
  float radianceFactor[3];
  int32_t diffuseTextureId;
  int32_t roughnessTextureId;
  int32_t metalnessTextureId;
  int32_t radianceTextureId;
  int32_t normalTextureId;
  int32_t tangentTextureId;
  int32_t opacityTextureId;
  uint32_t type;
} AsunaMaterial;

/* reference src/shared/light.h:7-13 */
enum AsunaLightType {
  ASUNA_LIGHT_DIRECTIONAL = 0,
  ASUNA_LIGHT_RECT = 1,
  ASUNA_LIGHT_TRIANGLE = 2,
  ASUNA_LIGHT_POINT = 3,
  ASUNA_LIGHT_UNDEFINED = 4
};

/* reference src/shared/light.h:16-25 -- 76 bytes */
typedef struct AsunaLight {
  int32_t type;
  float position[3];
  float direction[3];
  float radiance[3];
  float u[3];
  float v[3];
  float radius;
  float area;
  uint32_t doubleSide;
} AsunaLight;

/* reference src/shared/camera.h:7-11 */
enum AsunaCameraType { ASUNA_CAMERA_PERSPECTIVE = 0, ASUNA_CAMERA_OPENCV = 1 };

/* reference src/shared/camera.h:15-24 -- 224 bytes; matrices column-major (m[col*4+row]) */
typedef struct AsunaCamera {
  float rasterToCamera[16];
  float cameraToWorld[16];
  float envTransform[16];
  float fxfycxcy[4];
  uint32_t type;
  float aperture;
  float focalDistance;
  float padding;
} AsunaCamera;

/* reference src/shared/pushconstant.h:10-34 -- 84 bytes */
typedef struct AsunaState {
  int32_t spp;
  int32_t curFrame;
  int32_t maxPathDepth;
  int32_t numLights;
  float bgColor[3];
  uint32_t useFaceNormal;
  uint32_t ignoreEmissive;
  uint32_t hasEnvMap;
  float envMapResolution[2];
  float envMapIntensity;
  uint32_t nMultiChannel;
  int32_t diffuseOutChannel;
  int32_t specularOutChannel;
  int32_t roughnessOutChannel;
  int32_t normalOutChannel;
  int32_t positionOutChannel;
  int32_t tangentOutChannel;
  int32_t uvOutChannel;
} AsunaState;

/* reference src/shared/sun_and_sky.h:6-28 -- 96 bytes */
typedef struct AsunaSunSky {
  float rgb_unit_conversion[3];
  float multiplier;
  float haze;
  float redblueshift;
  float saturation;
  float horizon_height;
  float ground_color[3];
  float horizon_blur;
  float night_color[3];
  float sun_disk_intensity;
  float sun_direction[3];
  float sun_disk_scale;
  float sun_glow_intensity;
  int32_t y_is_up;
  int32_t physically_scaled_sun;
  int32_t in_use;
} AsunaSunSky;

#define ASUNA_NUM_OUTPUT_IMAGES 9 /* reference src/shared/binding.h:70 */

enum AsunaError {
  ASUNA_OK = 0,
  ASUNA_E_INVALID = -1,  /* bad argument / id out of range / wrong call order */
  ASUNA_E_CUDA = -2,     /* CUDA runtime error; text in asuna_last_error */
  ASUNA_E_NO_DEVICE = -3,
  ASUNA_E_UNSUPPORTED = -4 /* material type outside the hot-path scope */
};

/* GpuPushConstantPost, reference src/shared/pushconstant.h:49-62 (48 bytes): the post-process / tone-mapping state.
 * Defaults: reference src/core/state.h:46-58 (everything 1 / 0, Ywhite = key = 0.5, tmType Filmic). */
typedef struct AsunaPost {
  float brightness, contrast, saturation, vignette, avgLum, zoom;
  float renderingRatio[2];
  int32_t autoExposure; /* bit 0: global exposure from the image mean; bit 1: local (mip-chain) exposure -- unsupported */
  float Ywhite, key;
  uint32_t tmType; /* AsunaToneMapping */
} AsunaPost;
typedef enum AsunaToneMapping { /* reference src/shared/pushconstant.h:36-46 */
  ASUNA_TM_NONE = 0, ASUNA_TM_GAMMA = 1, ASUNA_TM_REINHARD = 2, ASUNA_TM_ACES = 3, ASUNA_TM_FILMIC = 4,
  ASUNA_TM_PBRT = 5, ASUNA_TM_CUSTOM = 6, ASUNA_TM_NUM = 7
} AsunaToneMapping;

/* Counters a caller may read after a render; all are totals since the last asuna_reset_stats.
 * Times are device times from CUDA events recorded on the context's stream around each launch, kept only while
 * asuna_set_profiling is on (build_ms is always measured). */
typedef struct AsunaStats {
  uint64_t paths;          /* pixel-samples started                         */
  uint64_t closest_rays;   /* closest-hit rays traced (rgen:108)            */
  uint64_t shadow_rays;    /* shadow rays traced (rgen:121)                 */
  uint64_t incoherent_closest_rays; /* closest-hit rays at depth >= 2       */
  uint64_t kernel_launches;  /* kernels launched by asuna_render_frames     */
  uint64_t closest_launches; /* launches of the closest-hit trace kernel    */
  uint64_t node_visits;    /* BVH nodes fetched by closest-hit rays (only with asuna_set_counting) */
  uint64_t tri_tests;      /* triangle tests by closest-hit rays   (only with asuna_set_counting) */
  float trace_ms;          /* closest-hit + shadow trace kernels            */
  float shade_ms;          /* raygen + shade + accumulate kernels           */
  float total_ms;          /* whole batches, launch gaps included           */
  float build_ms;          /* the last asuna_build_accel                    */
  float closest_ms;        /* closest-hit trace kernel alone                */
  float shadow_ms;         /* shadow trace kernel alone                     */
  float pad[2];
} AsunaStats;

typedef struct asuna_ctx asuna_ctx; /* opaque */

/* Struct sizes as compiled into the library: {vertex, material, light, camera, state,
 * sunsky}.  A binding calls this first to prove both sides agree on the wire format. */
void asuna_abi_sizes(uint32_t out_sizes[6]);

/* ≙ ContextAware::init({offline,gpuId}) reference src/context/context.cpp:8-27,304 */
int asuna_create(asuna_ctx** out, int gpu_id);
void asuna_destroy(asuna_ctx* ctx);
const char* asuna_last_error(asuna_ctx* ctx);

/* ---- scene upload (≙ Scene::submit, reference src/scene/scene.cpp:26-85) ---- */

/* ≙ film resolution, reference src/loader/loader.cpp:40-65 */
int asuna_set_film(asuna_ctx* ctx, uint32_t width, uint32_t height);
/* ≙ Scene::allocTexture (RGBA32F, LINEAR/REPEAT, LOD 0), reference src/core/texture.cpp:99-130.
 * Returns the texture id (insertion order; the host adds the dummy first, scene.cpp:93-98). */
int asuna_add_texture(asuna_ctx* ctx, const float* rgba32f, uint32_t width, uint32_t height);
/* ≙ Scene::allocEnvMap, three w*h RGBA32F tables, reference src/core/texture.cpp:144-226,240-296 */
int asuna_set_envmap(asuna_ctx* ctx, const float* rgba, const float* marginal,
                     const float* conditional, uint32_t width, uint32_t height);
/* ≙ Scene::allocMesh + MeshBufferToBlas input, reference src/core/mesh.cpp:79-101,147-192.  Returns mesh id. */
int asuna_add_mesh(asuna_ctx* ctx, const AsunaVertex* vertices, uint32_t n_vertices,
                   const uint32_t* indices, uint32_t n_indices);
/* ≙ Scene::allocMaterial, reference src/core/material.cpp:5-14.  Returns material id. */
int asuna_add_material(asuna_ctx* ctx, const AsunaMaterial* material);
/* ≙ Scene::allocLights; index 0 is the dummy, reference src/scene/scene.cpp:100-111,33 */
int asuna_set_lights(asuna_ctx* ctx, const AsunaLight* lights, uint32_t n_including_dummy);
/* ≙ one VkAccelerationStructureInstanceKHR + GpuInstance, reference
 * src/pipeline/pipeline_raytrace.cpp:120-142, src/core/instance.cpp:4-28.
 * light_id >= 0 marks an emitter instance (material is then ignored).  Returns instance id. */
int asuna_add_instance(asuna_ctx* ctx, const float xform_colmajor[16], uint32_t mesh_id,
                       uint32_t material_id, int32_t light_id);
/* ≙ createBottomLevelAS + createTopLevelAS, reference src/pipeline/pipeline_raytrace.cpp:107-147
 * (driver vkCmdBuildAccelerationStructuresKHR, ext/nvpro_core/nvvk/raytraceKHR_vk.cpp:216,375). */
int asuna_build_accel(asuna_ctx* ctx, float* out_build_ms);

/* ---- per shot / per frame ---- */

/* ≙ PipelineGraphics::run camera UBO update, reference src/pipeline/pipeline_graphics.cpp:50-98 */
int asuna_set_camera(asuna_ctx* ctx, const AsunaCamera* camera);
/* ≙ sun/sky UBO update, reference src/pipeline/pipeline_graphics.cpp:100-103 */
int asuna_set_sunsky(asuna_ctx* ctx, const AsunaSunSky* sunsky);
/* ≙ vkCmdPushConstants of rtxState, reference src/pipeline/pipeline_raytrace.cpp:58-60.
 * curFrame is taken from the struct; spp must be 1 (offline mode, tracer.cpp:211). */
int asuna_set_state(asuna_ctx* ctx, const AsunaState* state);
/* ≙ PipelineRaytrace::resetFrame (curFrame = -1), reference src/pipeline/pipeline_raytrace.cpp:80-82 */
int asuna_reset_frame(asuna_ctx* ctx);
/* ≙ n x PipelineRaytrace::run (incrementFrame + vkCmdTraceRaysKHR(w,h,1)), reference
 * src/pipeline/pipeline_raytrace.cpp:36-74 driven by tracer.cpp:218-226.  Asynchronous. */
int asuna_render_frames(asuna_ctx* ctx, uint32_t n_frames);
/* Multi-GPU sample-range split: this context renders only frames f with f % world == rank
 * (asuna_render_frames still advances curFrame by n_frames).  Default rank 0 / world 1. */
int asuna_set_partition(asuna_ctx* ctx, uint32_t rank, uint32_t world);
/* Waits for all queued work (≙ submitAndWait, reference src/tracer/tracer.cpp:226). */
int asuna_sync(asuna_ctx* ctx);

/* ≙ vkTextureToBuffer + map, reference src/tracer/tracer.cpp:313-342,365.
 * channel 0 = radiance mean, 1..7 = AOVs (frame 0), 8 = filter-weight sum.  w*h*4 floats. */
int asuna_read_channel(asuna_ctx* ctx, int channel, float* rgba32f_out);
/* The same read without stalling the context: the image is snapshotted on the context's stream (so later
 * asuna_reset_frame / asuna_render_frames calls may overwrite it at once) and the snapshot travels to the host on a
 * second stream while that later work runs -- the multi-shot loop of reference src/tracer/tracer.cpp:205-262 with the
 * copy of shot k hidden behind the rendering of shot k+1.  `rgba32f_out` should be page-locked (asuna_host_alloc) and
 * must not be read, written or freed before asuna_wait_reads returns; one read may be in flight per context, a second
 * one queues behind it on the device. */
int asuna_read_channel_async(asuna_ctx* ctx, int channel, float* rgba32f_out);
int asuna_wait_reads(asuna_ctx* ctx);

/* Page-locked host memory for the read-back / upload buffers of a caller (≙ the host-visible, host-coherent
 * staging buffer the reference maps in src/tracer/tracer.cpp:317-336): asuna_read_channel into such a buffer is
 * one DMA, into pageable memory the driver stages it.  Free with asuna_host_free. */
int asuna_host_alloc(asuna_ctx* ctx, size_t bytes, void** out_host_ptr);
int asuna_host_free(asuna_ctx* ctx, void* host_ptr);

/* Multi-GPU combine.  export: writes (sum_w*L.rgb, sum_w) per pixel into a device buffer
 * owned by the library and returns its device pointer (w*h*4 floats) for the caller's
 * reduce (NCCL sum).  import: reads that buffer back after the reduce and stores
 * L = sum_wL / sum_w into image 0 and sum_w into image 8.
 * Both are ordered on the context's stream (asuna_stream_handle) and do not synchronise the host: a collective issued on
 * that stream needs nothing else; one issued on another stream must wait for it and be waited for (events /
 * wait_stream), or bracket the call with asuna_sync. */
int asuna_export_partial(asuna_ctx* ctx, void** out_device_ptr);
int asuna_import_partial(asuna_ctx* ctx);

/* ≙ PipelinePost::run (reference src/pipeline/pipeline_post.cpp:24-43, src/shaders/post.idle.frag:71-133) followed by
 * the read of the offline colour image (src/tracer/tracer.cpp:236-255, 358-359): tone-maps radiance image 0 on the
 * device with all seven tone mappers of the reference (custom: Uncharted-2 curve, dithering, contrast / brightness /
 * saturation / vignette, optional global auto-exposure) and returns w*h RGBA32F in host memory. */
int asuna_post_process(asuna_ctx* ctx, const AsunaPost* tm, float* rgba32f_out);

/* Device pointer of output image `channel` (w*h float4), for zero-copy consumers. */
int asuna_channel_device_ptr(asuna_ctx* ctx, int channel, void** out_device_ptr);

/* CUstream the context launches on (so a caller can bracket work with its own events). */
int asuna_stream_handle(asuna_ctx* ctx, void** out_stream);
/* Instrumented traversal: when on, closest-hit launches also count node visits / triangle tests
 * (slower; used to derive the algorithmic bytes per ray of the roofline, never while timing). */
int asuna_set_counting(asuna_ctx* ctx, int on);

/* Profiling: when on, every kernel group of asuna_render_frames is bracketed by a CUDA-event pair feeding the *_ms
 * fields of AsunaStats.  Off by default (counts are always kept): an integration that renders frame after frame
 * (reference src/tracer/tracer.cpp:218-228) records no events. */
int asuna_set_profiling(asuna_ctx* ctx, int on);

int asuna_get_stats(asuna_ctx* ctx, AsunaStats* out);
int asuna_reset_stats(asuna_ctx* ctx);

/* Introspection used by the parity tests and the roofline arithmetic. */
/* Primary-visibility query: traces the frame-0 camera rays (pixel centres) and returns
 * per pixel {instance id, primitive id} (0xFFFFFFFF on miss) and hit distance t. */
int asuna_trace_primary(asuna_ctx* ctx, uint32_t* inst_prim_out /* w*h*2 */, float* t_out /* w*h */);
/* Traces caller-supplied rays (n x {ox,oy,oz,tmin,dx,dy,dz,tmax}) with the closest-hit
 * kernel; outputs n x {t,u,v} and n x {inst,prim}.  Host pointers. */
int asuna_trace_rays(asuna_ctx* ctx, const float* rays, uint32_t n, float* tuv_out,
                     uint32_t* inst_prim_out);
/* Same rays through the any-hit (shadow) kernel; out[i] = 1 if occluded. */
int asuna_occlusion_rays(asuna_ctx* ctx, const float* rays, uint32_t n, uint8_t* occluded_out);
/* BVH statistics: {wide nodes in use over all mesh BVHs, primitive slots, wide nodes of the instance BVH,
 * sum over mesh BVHs of SAH cost C(root)/area(root) x 1000 (c_node 1, c_triangle 0.3)}. */
int asuna_accel_stats(asuna_ctx* ctx, uint64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* ASUNA_B200_H */
