// C ABI of libasuna_b200.so (include/asuna_b200.h): context, scene upload, acceleration-structure
// build and the per-frame driver that sequences the wavefront kernels.
//
// This object sits where the reference's PipelineRaytrace sits (reference
// src/pipeline/pipeline_raytrace.{h,cpp}): init() ≙ asuna_build_accel, run() ≙ asuna_render_frames,
// the nine storage images of PipelineGraphics (reference src/pipeline/pipeline_graphics.cpp:126-138)
// ≙ OutputImages.  There is no CPU fallback: without a CUDA device asuna_create fails.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

#include "bvh_build.cuh"
#include "integrator.cuh"

using namespace asuna;

// ASUNA_TIMING=1: host-side wall time of the steps of asuna_create / asuna_build_accel on stderr (developer probe for
// the serial part of a job, tools/upload_probe.py)
static double wall_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
struct StepTimer {
  bool on = getenv("ASUNA_TIMING") != nullptr;
  double t = wall_ms();
  const char* what;
  explicit StepTimer(const char* w) : what(w) {}
  void step(const char* name) {
    if (!on) return;
    double n = wall_ms();
    fprintf(stderr, "[asuna timing] %s: %s %.3f ms\n", what, name, n - t);
    t = n;
  }
};

namespace {
struct NvtxRange {  // one named range per pipeline stage (tracing / profiling hooks)
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
struct HostMesh {
  AsunaVertex* d_vertices = nullptr;
  uint32_t* d_indices = nullptr;
  uint32_t n_vertices = 0, n_tris = 0;
  int node_base = 0, tri_base = 0;
};
struct HostTexture {
  float4* d_texels = nullptr;
  uint32_t w = 0, h = 0;
};
struct HostInstance {
  float xform[16];
  uint32_t mesh, material;
  int32_t light;
};
struct TimedEvent {
  cudaEvent_t a, b;
  int kind;  // 0 trace, 1 shade/other
};
}  // namespace

struct asuna_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::string err;

  uint32_t W = 0, H = 0;
  std::vector<HostTexture> textures;
  HostTexture env[3];
  std::vector<HostMesh> meshes;
  std::vector<AsunaMaterial> materials;
  std::vector<AsunaLight> lights;
  std::vector<HostInstance> instances;
  bool scene_dirty = true;

  // device scene
  WideNode *d_blas_nodes = nullptr, *d_tlas_nodes = nullptr;
  BuildResult* d_build_results = nullptr;
  TriSlot* d_tris = nullptr;
  uint32_t *d_tlas_leaf_inst = nullptr, *d_tlas_ids = nullptr;
  DInstance* d_instances = nullptr;
  DMesh* d_meshes = nullptr;
  AsunaMaterial* d_materials = nullptr;
  AsunaLight* d_lights = nullptr;
  DTexture* d_textures = nullptr;
  float4 *d_mesh_lo = nullptr, *d_mesh_hi = nullptr;
  SceneView view{};
  BuildScratch scratch;
  uint64_t accel_stats[4] = {0, 0, 0, 0};
  size_t pool_nodes = 0, pool_tris = 0;  // allocated entries of d_blas_nodes / d_tris (asuna_debug_download_accel)
  bool may_pass_through = false;
  uint32_t kind_mask = 0;  // hit kinds that can occur in this scene (which shade kernels to launch)

  // frame state
  AsunaCamera cam{};
  AsunaSunSky sunsky{};
  SkyPre sky{};  // the per-setting part of the sun & sky model for `sunsky`, evaluated on the device when it changes
  SkyPre* d_sky = nullptr;
  bool sky_valid = false;
  AsunaState pc{};
  uint32_t rank = 0, world = 1;
  bool have_accum = false;
  OutputImages out{};
  float4* d_partial = nullptr;
  float4* d_ldr = nullptr;        // post-processed (tone-mapped) image
  double* d_post_sums = nullptr;  // block sums of the image mean (custom tone mapper's auto-exposure)
  void* film_arena = nullptr;   // the nine planes + the partial-sum plane: one allocation
  void* path_arena = nullptr;   // the whole wavefront state: one allocation
  void* scene_arena = nullptr;  // nodes, triangle slots and the flat scene tables: one allocation per build
  cudaStream_t upload_stream = nullptr;  // H2D copies of textures / env tables / meshes overlap the BVH build
  // asuna_read_channel_async: the image is snapshotted into d_readback on `stream`, the D2H copy of the snapshot runs on
  // read_stream while later work on `stream` proceeds
  cudaStream_t read_stream = nullptr;
  cudaEvent_t ev_read_ready = nullptr, ev_read_done = nullptr;
  float4* d_readback = nullptr;
  size_t readback_px = 0;
  bool read_pending = false;
  cudaEvent_t ev_geometry = nullptr, ev_textures = nullptr, ev_partial = nullptr;
  bool textures_pending = false;

  // wavefront buffers
  PathState ps{};
  Counters* d_counters = nullptr;
  Totals* d_totals = nullptr;
  Totals* h_totals = nullptr;  // pinned
  bool counting = false;
  uint32_t path_capacity = 0;
  uint32_t max_batch_frames = ASUNA_MAX_BATCH_FRAMES;  // capped below so that a batch stays under ~16 M paths (8 frames at 1080p, 64 at 512x512)
  LaunchDims dims;

  // user-ray scratch
  float4* d_user_rays = nullptr;
  float* d_user_tuv = nullptr;
  uint32_t* d_user_ip = nullptr;
  uint8_t* d_user_occ = nullptr;
  uint32_t user_capacity = 0;

  AsunaStats stats{};
  bool profiling = false;  // asuna_set_profiling: per-launch CUDA-event pairs feeding the *_ms statistics
  std::vector<TimedEvent> pending;
  std::vector<cudaEvent_t> event_pool;
};

void asuna_set_cuda_error(asuna_ctx* ctx, cudaError_t e, const char* expr, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, expr);
  if (ctx) ctx->err = buf;
}

namespace {

int fail(asuna_ctx* ctx, int code, const char* msg) {
  ctx->err = msg;
  return code;
}

template <class T>
void free_dev(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

// Sizes first, then one cudaMalloc and the pointers into it (a 1080p context used to pay ~50 cudaMalloc calls before its
// first frame; each costs 0.1-2 ms of host time that no amount of GPU parallelism gives back).
struct ArenaPlan {
  struct Item {
    void** p;
    size_t bytes;
  };
  std::vector<Item> items;
  template <class T>
  void add(T*& p, size_t bytes) {
    items.push_back({reinterpret_cast<void**>(&p), bytes});
  }
  cudaError_t commit(void** arena) {
    size_t total = 0;
    for (auto& it : items) total += (it.bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(arena, std::max<size_t>(total, 256));
    if (e != cudaSuccess) return e;
    size_t off = 0;
    for (auto& it : items) {
      *it.p = static_cast<char*>(*arena) + off;
      off += (it.bytes + 255) & ~(size_t)255;
    }
    return cudaSuccess;
  }
};

cudaEvent_t get_event(asuna_ctx* ctx) {
  if (!ctx->event_pool.empty()) {
    cudaEvent_t e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void collect_timers(asuna_ctx* ctx);
// Brackets stream work with two events when profiling is on (asuna_set_profiling; off by default, so an
// integration that renders frame after frame without ever asking for statistics records nothing).  Elapsed times
// are collected wherever the stream is synchronised anyway; the backlog is bounded.
struct ScopedTimer {
  asuna_ctx* ctx;
  TimedEvent te;
  bool on;
  ScopedTimer(asuna_ctx* c, int kind) : ctx(c), on(c->profiling) {
    if (!on) return;
    te.a = get_event(c);
    te.b = get_event(c);
    te.kind = kind;
    cudaEventRecord(te.a, c->stream);
  }
  ~ScopedTimer() {
    if (!on) return;
    cudaEventRecord(te.b, ctx->stream);
    ctx->pending.push_back(te);
    if (ctx->pending.size() >= 4096) {  // cap: drain instead of growing without bound
      cudaEventSynchronize(te.b);
      collect_timers(ctx);
    }
  }
};
void collect_timers(asuna_ctx* ctx) {
  for (auto& te : ctx->pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, te.a, te.b) == cudaSuccess) {
      if (te.kind == 0) ctx->stats.trace_ms += ms, ctx->stats.closest_ms += ms;
      else if (te.kind == 3) ctx->stats.trace_ms += ms, ctx->stats.shadow_ms += ms;
      else if (te.kind == 1) ctx->stats.shade_ms += ms;
      else if (te.kind == 2) ctx->stats.total_ms += ms;
    }
    ctx->event_pool.push_back(te.a);
    ctx->event_pool.push_back(te.b);
  }
  ctx->pending.clear();
}

void free_scene_device(asuna_ctx* ctx) {
  free_dev(ctx->scene_arena);
  ctx->d_blas_nodes = ctx->d_tlas_nodes = nullptr;
  ctx->d_tris = nullptr;
  ctx->d_tlas_leaf_inst = ctx->d_tlas_ids = nullptr;
  ctx->d_instances = nullptr;
  ctx->d_meshes = nullptr;
  ctx->d_materials = nullptr;
  ctx->d_lights = nullptr;
  ctx->d_textures = nullptr;
  ctx->d_mesh_lo = ctx->d_mesh_hi = nullptr;
  ctx->d_build_results = nullptr;
}

void free_path_buffers(asuna_ctx* ctx) {
  free_dev(ctx->path_arena);
  ctx->ps = PathState{};
  ctx->path_capacity = 0;
}

int ensure_path_buffers(asuna_ctx* ctx, uint32_t n_paths) {
  if (n_paths <= ctx->path_capacity) return 0;
  StepTimer tm("path state");
  free_path_buffers(ctx);
  size_t n = n_paths;
  ArenaPlan plan;
  plan.add(ctx->ps.ray_o, n * sizeof(float4)), plan.add(ctx->ps.ray_d, n * sizeof(float4));
  plan.add(ctx->ps.thr, n * sizeof(float4)), plan.add(ctx->ps.rad, n * sizeof(float4));
  plan.add(ctx->ps.hit, n * sizeof(uint4));
  plan.add(ctx->ps.sh_o, n * sizeof(float4)), plan.add(ctx->ps.sh_d, n * sizeof(float4)), plan.add(ctx->ps.sh_l, n * sizeof(float4));
  plan.add(ctx->ps.queue[0], n * sizeof(uint32_t)), plan.add(ctx->ps.queue[1], n * sizeof(uint32_t));
  plan.add(ctx->ps.kind, n), plan.add(ctx->ps.sorted, n * sizeof(uint32_t));
  plan.add(ctx->ps.bin_hist, ((n + kBinTile - 1) / kBinTile) * kNumKinds * sizeof(uint32_t));
  ASUNA_CUDA_CHECK(plan.commit(&ctx->path_arena));
  tm.step("allocation");
  ctx->path_capacity = n_paths;
  return 0;
}

// world->object of a column-major 4x4 affine transform, inverted in double (what the driver
// derives from VkAccelerationStructureInstanceKHR::transform)
void make_instance_matrices(const float x[16], DInstance& d) {
  float o2w[12];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 4; c++) o2w[r * 4 + c] = x[c * 4 + r];
  double a[9], inv[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) a[r * 3 + c] = o2w[r * 4 + c];
  double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  double id = 1.0 / det;
  inv[0] = (a[4] * a[8] - a[5] * a[7]) * id, inv[1] = (a[2] * a[7] - a[1] * a[8]) * id, inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  inv[3] = (a[5] * a[6] - a[3] * a[8]) * id, inv[4] = (a[0] * a[8] - a[2] * a[6]) * id, inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  inv[6] = (a[3] * a[7] - a[4] * a[6]) * id, inv[7] = (a[1] * a[6] - a[0] * a[7]) * id, inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  float w2o[12];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) w2o[r * 4 + c] = (float)inv[r * 3 + c];
    w2o[r * 4 + 3] = (float)-(inv[r * 3 + 0] * o2w[3] + inv[r * 3 + 1] * o2w[7] + inv[r * 3 + 2] * o2w[11]);
  }
  for (int r = 0; r < 3; r++) {
    d.o2w[r] = make_float4(o2w[r * 4 + 0], o2w[r * 4 + 1], o2w[r * 4 + 2], o2w[r * 4 + 3]);
    d.w2o[r] = make_float4(w2o[r * 4 + 0], w2o[r * 4 + 1], w2o[r * 4 + 2], w2o[r * 4 + 3]);
  }
}

bool in_scope(uint32_t type) { return type < ASUNA_MAT_NUM; }  // all twelve closest-hit shaders of the reference

FrameParams make_frame_params(asuna_ctx* ctx) {
  FrameParams fp{};
  fp.cam = ctx->cam;
  fp.sunsky = ctx->sunsky;
  fp.sky = ctx->sky;
  fp.pc = ctx->pc;
  fp.width = ctx->W;
  fp.height = ctx->H;
  fp.n_pixels = ctx->W * ctx->H;
  return fp;
}

int upload_user_rays(asuna_ctx* ctx, const float* rays, uint32_t n) {
  if (n > ctx->user_capacity) {
    free_dev(ctx->d_user_rays);
    free_dev(ctx->d_user_tuv);
    free_dev(ctx->d_user_ip);
    free_dev(ctx->d_user_occ);
    ASUNA_CUDA_CHECK(cudaMalloc(&ctx->d_user_rays, (size_t)n * 2 * sizeof(float4)));
    ASUNA_CUDA_CHECK(cudaMalloc(&ctx->d_user_tuv, (size_t)n * 3 * sizeof(float)));
    ASUNA_CUDA_CHECK(cudaMalloc(&ctx->d_user_ip, (size_t)n * 2 * sizeof(uint32_t)));
    ASUNA_CUDA_CHECK(cudaMalloc(&ctx->d_user_occ, (size_t)n));
    ctx->user_capacity = n;
  }
  if (rays)
    ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_user_rays, rays, (size_t)n * 8 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

}  // namespace

extern "C" {

void asuna_abi_sizes(uint32_t out[6]) {
  out[0] = sizeof(AsunaVertex), out[1] = sizeof(AsunaMaterial), out[2] = sizeof(AsunaLight);
  out[3] = sizeof(AsunaCamera), out[4] = sizeof(AsunaState), out[5] = sizeof(AsunaSunSky);
}

int asuna_create(asuna_ctx** out, int gpu_id) {
  *out = nullptr;
  StepTimer tm("create");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || gpu_id < 0 || gpu_id >= n) return ASUNA_E_NO_DEVICE;
  if (cudaSetDevice(gpu_id) != cudaSuccess) return ASUNA_E_NO_DEVICE;
  asuna_ctx* ctx = new asuna_ctx();
  ctx->device = gpu_id;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, gpu_id) != cudaSuccess) {
    delete ctx;
    return ASUNA_E_NO_DEVICE;
  }
  ctx->sm_count = prop.multiProcessorCount;
  tm.step("device properties");
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_geometry, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_textures, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_partial, cudaEventDisableTiming) != cudaSuccess ||
      cudaMalloc(&ctx->d_counters, sizeof(Counters)) != cudaSuccess ||
      cudaMalloc(&ctx->d_totals, sizeof(Totals)) != cudaSuccess ||
      cudaMemset(ctx->d_totals, 0, sizeof(Totals)) != cudaSuccess ||
      cudaMallocHost(&ctx->h_totals, sizeof(Totals)) != cudaSuccess) {
    delete ctx;
    return ASUNA_E_CUDA;
  }
  tm.step("streams, events, counters");
  if (query_launch_dims(ctx->dims, ctx->sm_count) != cudaSuccess) {
    delete ctx;
    return ASUNA_E_CUDA;
  }
  tm.step("occupancy queries");
  ctx->pc.curFrame = -1;
  ctx->pc.spp = 1;
  ctx->pc.maxPathDepth = 3;
  ctx->pc.envMapIntensity = 1.f;
  const float id[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  memcpy(ctx->cam.envTransform, id, sizeof id);
  *out = ctx;
  return 0;
}

void asuna_destroy(asuna_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->upload_stream) cudaStreamSynchronize(ctx->upload_stream);
  if (ctx->read_stream) cudaStreamSynchronize(ctx->read_stream);
  collect_timers(ctx);
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : {ctx->ev_geometry, ctx->ev_textures, ctx->ev_partial, ctx->ev_read_ready, ctx->ev_read_done})
    if (e) cudaEventDestroy(e);
  free_scene_device(ctx);
  free_path_buffers(ctx);
  for (auto& t : ctx->textures) free_dev(t.d_texels);
  for (auto& t : ctx->env) free_dev(t.d_texels);
  for (auto& m : ctx->meshes) {
    free_dev(m.d_vertices);
    free_dev(m.d_indices);
  }
  free_dev(ctx->film_arena);
  free_dev(ctx->d_counters);
  free_dev(ctx->d_totals);
  free_dev(ctx->d_sky);
  free_dev(ctx->d_user_rays);
  free_dev(ctx->d_user_tuv);
  free_dev(ctx->d_user_ip);
  free_dev(ctx->d_user_occ);
  if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
  ctx->scratch.release();
  cudaStreamDestroy(ctx->stream);
  if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
  if (ctx->read_stream) cudaStreamDestroy(ctx->read_stream);
  free_dev(ctx->d_readback);
  delete ctx;
}

const char* asuna_last_error(asuna_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int asuna_set_film(asuna_ctx* ctx, uint32_t w, uint32_t h) {
  if (w == 0 || h == 0) return fail(ctx, ASUNA_E_INVALID, "film resolution must be non-zero");
  cudaSetDevice(ctx->device);
  ctx->W = w, ctx->H = h;
  size_t bytes = (size_t)w * h * sizeof(float4);
  StepTimer tm("set_film");
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  free_dev(ctx->film_arena);
  ArenaPlan plan;
  for (auto& p : ctx->out.img) plan.add(p, bytes);
  plan.add(ctx->d_partial, bytes);
  plan.add(ctx->d_ldr, bytes);
  plan.add(ctx->d_post_sums, 3 * 296 * sizeof(double));
  ASUNA_CUDA_CHECK(plan.commit(&ctx->film_arena));
  ASUNA_CUDA_CHECK(cudaMemsetAsync(ctx->film_arena, 0, ((bytes + 255) & ~(size_t)255) * ASUNA_NUM_OUTPUT_IMAGES, ctx->stream));
  tm.step("allocation + clear enqueue");
  ctx->have_accum = false;
  return 0;
}

static int upload_texture(asuna_ctx* ctx, HostTexture& t, const float* rgba, uint32_t w, uint32_t h) {
  StepTimer tm("texture");
  free_dev(t.d_texels);
  t.w = w, t.h = h;
  ASUNA_CUDA_CHECK(cudaMalloc(&t.d_texels, (size_t)w * h * sizeof(float4)));
  tm.step("allocation");
  // On the upload stream: the call returns once the host buffer has been consumed (pageable memory is staged by the
  // driver, so the pointer may be reused at once), the DMA itself overlaps whatever the caller does next -- usually
  // asuna_build_accel.  Rendering waits for ev_textures.
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(t.d_texels, rgba, (size_t)w * h * sizeof(float4), cudaMemcpyHostToDevice, ctx->upload_stream));
  ctx->textures_pending = true;
  if (w * h > 65536) tm.step("copy (host side)");
  return 0;
}

int asuna_add_texture(asuna_ctx* ctx, const float* rgba, uint32_t w, uint32_t h) {
  if (!rgba || w == 0 || h == 0) return fail(ctx, ASUNA_E_INVALID, "empty texture");
  cudaSetDevice(ctx->device);
  ctx->textures.emplace_back();
  int rc = upload_texture(ctx, ctx->textures.back(), rgba, w, h);
  if (rc) {
    ctx->textures.pop_back();
    return rc;
  }
  ctx->scene_dirty = true;
  return (int)ctx->textures.size() - 1;
}

int asuna_set_envmap(asuna_ctx* ctx, const float* rgba, const float* marginal, const float* conditional, uint32_t w,
                     uint32_t h) {
  if (!rgba || !marginal || !conditional || w == 0 || h == 0) return fail(ctx, ASUNA_E_INVALID, "empty env map");
  cudaSetDevice(ctx->device);
  const float* src[3] = {rgba, marginal, conditional};
  for (int k = 0; k < 3; k++) {
    int rc = upload_texture(ctx, ctx->env[k], src[k], w, h);
    if (rc) return rc;
  }
  ctx->scene_dirty = true;
  return 0;
}

int asuna_add_mesh(asuna_ctx* ctx, const AsunaVertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni) {
  if (!v || !idx || nv == 0 || ni < 3) return fail(ctx, ASUNA_E_INVALID, "empty mesh");
  uint32_t nt = ni / 3;
  for (uint32_t i = 0; i < nt * 3; i++)
    if (idx[i] >= nv) return fail(ctx, ASUNA_E_INVALID, "index out of range");
  cudaSetDevice(ctx->device);
  HostMesh m;
  m.n_vertices = nv, m.n_tris = nt;
  ASUNA_CUDA_CHECK(cudaMalloc(&m.d_vertices, (size_t)nv * sizeof(AsunaVertex)));
  ASUNA_CUDA_CHECK(cudaMalloc(&m.d_indices, (size_t)nt * 3 * sizeof(uint32_t)));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(m.d_vertices, v, (size_t)nv * sizeof(AsunaVertex), cudaMemcpyHostToDevice, ctx->stream));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(m.d_indices, idx, (size_t)nt * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  ctx->meshes.push_back(m);
  ctx->scene_dirty = true;
  return (int)ctx->meshes.size() - 1;
}

int asuna_add_material(asuna_ctx* ctx, const AsunaMaterial* m) {
  if (!m) return fail(ctx, ASUNA_E_INVALID, "null material");
  if (!in_scope(m->type)) return fail(ctx, ASUNA_E_UNSUPPORTED, "unknown material type");
  ctx->materials.push_back(*m);
  ctx->scene_dirty = true;
  return (int)ctx->materials.size() - 1;
}

int asuna_set_lights(asuna_ctx* ctx, const AsunaLight* l, uint32_t n) {
  if (!l || n == 0) return fail(ctx, ASUNA_E_INVALID, "light table must hold at least the dummy");
  ctx->lights.assign(l, l + n);
  ctx->scene_dirty = true;
  return 0;
}

int asuna_add_instance(asuna_ctx* ctx, const float x[16], uint32_t mesh, uint32_t material, int32_t light) {
  if (mesh >= ctx->meshes.size()) return fail(ctx, ASUNA_E_INVALID, "instance refers to unknown mesh");
  if (light < 0 && material >= ctx->materials.size()) return fail(ctx, ASUNA_E_INVALID, "instance refers to unknown material");
  HostInstance in;
  memcpy(in.xform, x, sizeof in.xform);
  in.mesh = mesh, in.material = material, in.light = light;
  ctx->instances.push_back(in);
  ctx->scene_dirty = true;
  return (int)ctx->instances.size() - 1;
}

int asuna_build_accel(asuna_ctx* ctx, float* out_ms) {
  NvtxRange nvtx("asuna_build_accel");
  StepTimer tm("build_accel");
  cudaSetDevice(ctx->device);
  if (ctx->instances.empty() || ctx->meshes.empty()) return fail(ctx, ASUNA_E_INVALID, "scene has no instances");
  for (auto& in : ctx->instances)
    if (in.light >= 0 && (size_t)in.light >= ctx->lights.size()) return fail(ctx, ASUNA_E_INVALID, "emitter instance refers to unknown light");
  for (auto& m : ctx->materials) {
    const int32_t ids[7] = {m.diffuseTextureId, m.roughnessTextureId, m.metalnessTextureId, m.radianceTextureId,
                            m.normalTextureId,  m.tangentTextureId,   m.opacityTextureId};
    for (int32_t id : ids)
      if (id >= (int32_t)ctx->textures.size()) return fail(ctx, ASUNA_E_INVALID, "material refers to unknown texture");
  }
  cudaStreamSynchronize(ctx->stream);
  free_scene_device(ctx);
  cudaStream_t s = ctx->stream;

  // Flattening: meshes used by exactly one instance are taken to world space at build time and share ONE merged
  // BLAS ("world BLAS"), so a ray pays neither a transform nor a second tree for them; instanced meshes keep
  // their own object-space BLAS.  When every instance is merged the top level disappears (single-level
  // traversal).  ASUNA_FLATTEN=0 keeps the plain two-level structure (A/B experiments, tests).
  uint32_t n_inst = (uint32_t)ctx->instances.size(), n_mesh = (uint32_t)ctx->meshes.size();
  std::vector<uint32_t> uses(n_mesh, 0);
  for (auto& in : ctx->instances) uses[in.mesh]++;
  bool flatten = true;
  if (const char* f = getenv("ASUNA_FLATTEN")) flatten = atoi(f) != 0;
  // Instanced meshes are flattened too (one world-space copy per instance) while the whole scene stays under a triangle
  // budget: 180 GB of HBM buy single-level traversal where a Vulkan driver instances to save memory.  48 B per triangle
  // slot + ~12 B of nodes: the default 64 M triangles are 4 GB (plus the builder's scratch); ASUNA_FLATTEN_MAX_TRIS=0
  // restricts flattening to single-use meshes.
  uint64_t flat_budget = 64ull << 20, instanced_tris = 0;
  if (const char* f = getenv("ASUNA_FLATTEN_MAX_TRIS")) flat_budget = strtoull(f, nullptr, 10);
  for (auto& in : ctx->instances) instanced_tris += ctx->meshes[in.mesh].n_tris;
  const bool flatten_all = flatten && instanced_tris <= flat_budget && instanced_tris < (1u << 27);
  std::vector<uint32_t> merged, tlas_ids;  // instance ids in the world BLAS / instance records under the top level
  for (uint32_t i = 0; i < n_inst; i++)
    (flatten && (flatten_all || uses[ctx->instances[i].mesh] == 1) ? merged : tlas_ids).push_back(i);
  if (merged.size() == 1 && !tlas_ids.empty()) tlas_ids.push_back(merged[0]), merged.clear();  // nothing to merge with
  std::vector<char> mesh_merged(n_mesh, 0);
  for (uint32_t i : merged) mesh_merged[ctx->instances[i].mesh] = 1;
  std::sort(tlas_ids.begin(), tlas_ids.end());
  const bool have_world = !merged.empty();
  const bool single_level = have_world && tlas_ids.empty();
  if (have_world) tlas_ids.push_back(n_inst);  // the world BLAS enters the top level as pseudo-instance n_inst
  const uint32_t n_tlas = (uint32_t)tlas_ids.size();

  // layout of the node / triangle pools
  size_t total_nodes = 0, total_tris = 0;
  uint32_t max_prims = n_tlas, world_tris = 0;
  for (uint32_t i = 0; i < n_mesh; i++) {
    HostMesh& m = ctx->meshes[i];
    if (mesh_merged[i]) {  // lives in the world BLAS only (one copy per instance, counted below)
      m.node_base = m.tri_base = -1;
      continue;
    }
    m.node_base = (int)total_nodes;
    m.tri_base = (int)total_tris;
    total_nodes += std::max<uint32_t>(m.n_tris - 1, 1);
    total_tris += m.n_tris;
    max_prims = std::max(max_prims, m.n_tris);
  }
  const size_t world_node_base = total_nodes, world_tri_base = total_tris;
  for (uint32_t i : merged) world_tris += ctx->meshes[ctx->instances[i].mesh].n_tris;
  if (have_world) {
    total_nodes += std::max<uint32_t>(world_tris - 1, 1);
    total_tris += world_tris;
    max_prims = std::max(max_prims, world_tris);
  }
  if (total_tris >= (1u << 27)) return fail(ctx, ASUNA_E_INVALID, "more than 2^27 triangles");
  if (n_inst >= (1u << 28)) return fail(ctx, ASUNA_E_INVALID, "more than 2^28 instances");
  ctx->pool_nodes = total_nodes, ctx->pool_tris = total_tris;
  ASUNA_CUDA_CHECK(ctx->scratch.reserve(max_prims));
  {
    ArenaPlan plan;
    plan.add(ctx->d_blas_nodes, total_nodes * sizeof(WideNode));
    plan.add(ctx->d_build_results, (n_mesh + 2) * sizeof(BuildResult));
    plan.add(ctx->d_tris, total_tris * sizeof(TriSlot));
    plan.add(ctx->d_tlas_nodes, std::max<uint32_t>(n_tlas - 1, 1) * sizeof(WideNode));
    plan.add(ctx->d_tlas_leaf_inst, n_tlas * sizeof(uint32_t));
    plan.add(ctx->d_tlas_ids, n_tlas * sizeof(uint32_t));
    plan.add(ctx->d_instances, (n_inst + 1) * sizeof(DInstance));
    plan.add(ctx->d_meshes, n_mesh * sizeof(DMesh));
    plan.add(ctx->d_mesh_lo, (n_mesh + 1) * sizeof(float4));
    plan.add(ctx->d_mesh_hi, (n_mesh + 1) * sizeof(float4));
    plan.add(ctx->d_materials, std::max<size_t>(ctx->materials.size(), 1) * sizeof(AsunaMaterial));
    plan.add(ctx->d_lights, std::max<size_t>(ctx->lights.size(), 1) * sizeof(AsunaLight));
    plan.add(ctx->d_textures, std::max<size_t>(ctx->textures.size(), 1) * sizeof(DTexture));
    ASUNA_CUDA_CHECK(plan.commit(&ctx->scene_arena));
  }
  ASUNA_CUDA_CHECK(cudaMemsetAsync(ctx->d_build_results, 0, (n_mesh + 2) * sizeof(BuildResult), s));
  TriSlot* d_soup = nullptr;  // world-space triangles in input order; the emit kernel copies them into leaf order
  WorldJob* d_jobs = nullptr;
  if (have_world) {
    // one temporary allocation: the soup and, behind it, the per-instance job table of k_world_triangles_batched
    const size_t soup_bytes = ((size_t)world_tris * sizeof(TriSlot) + 255) & ~(size_t)255;
    ASUNA_CUDA_CHECK(cudaMalloc(&d_soup, soup_bytes + merged.size() * sizeof(WorldJob)));
    d_jobs = reinterpret_cast<WorldJob*>(reinterpret_cast<char*>(d_soup) + soup_bytes);
  }

  // flat tables
  std::vector<DInstance> hinst(n_inst + 1);
  for (uint32_t i = 0; i < n_inst; i++) {
    const HostInstance& in = ctx->instances[i];
    DInstance d{};
    make_instance_matrices(in.xform, d);
    d.blas_root = ctx->meshes[in.mesh].node_base;  // -1: lives in the world BLAS
    d.mesh = in.mesh, d.material = in.material, d.light = in.light;
    d.mat_type = in.light >= 0 ? 0xFFFFFFFFu : ctx->materials[in.material].type;
    hinst[i] = d;
  }
  {  // pseudo-instance of the world BLAS: identity transform, "mesh" n_mesh (its box slot)
    static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    DInstance d{};
    make_instance_matrices(ident, d);
    d.blas_root = (int32_t)world_node_base;
    d.mesh = n_mesh, d.material = 0, d.light = -1, d.mat_type = 0;
    hinst[n_inst] = d;
  }
  std::vector<DMesh> hmesh(n_mesh);
  for (uint32_t i = 0; i < n_mesh; i++) hmesh[i] = DMesh{ctx->meshes[i].d_vertices, ctx->meshes[i].d_indices, ctx->meshes[i].n_tris, 0};
  std::vector<DTexture> htex(ctx->textures.size());
  for (size_t i = 0; i < htex.size(); i++) htex[i] = DTexture{ctx->textures[i].d_texels, (int)ctx->textures[i].w, (int)ctx->textures[i].h};
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_instances, hinst.data(), (n_inst + 1) * sizeof(DInstance), cudaMemcpyHostToDevice, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_tlas_ids, tlas_ids.data(), n_tlas * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_meshes, hmesh.data(), n_mesh * sizeof(DMesh), cudaMemcpyHostToDevice, s));
  if (!ctx->materials.empty())
    ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_materials, ctx->materials.data(), ctx->materials.size() * sizeof(AsunaMaterial), cudaMemcpyHostToDevice, s));
  if (!ctx->lights.empty())
    ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_lights, ctx->lights.data(), ctx->lights.size() * sizeof(AsunaLight), cudaMemcpyHostToDevice, s));
  if (!htex.empty())
    ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_textures, htex.data(), htex.size() * sizeof(DTexture), cudaMemcpyHostToDevice, s));
  tm.step("allocation + table upload enqueue");
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(s));  // host vectors go out of scope below; also excludes H2D from build time
  tm.step("wait for uploads");

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  // bottom level: one wide BVH per mesh (≙ createBottomLevelAS)
  float cost_prim = 1.0f;  // SAH: one triangle test relative to one wide-node step (measured: 0.3 -> 1.0 = +3-6 % rays/s)
  if (const char* t = getenv("ASUNA_TUNE")) {
    unsigned a, b, c = 0;
    if (sscanf(t, "%u,%u,%u", &a, &b, &c) == 3 && c > 0) cost_prim = 0.1f * (float)c;
  }
  for (uint32_t i = 0; i < n_mesh; i++) {
    HostMesh& m = ctx->meshes[i];
    if (mesh_merged[i]) continue;
    launch_tri_boxes(s, m.d_vertices, m.d_indices, m.n_tris, ctx->scratch);
    PrimPayload pl;
    pl.vertices = m.d_vertices, pl.indices = m.d_indices, pl.tris = ctx->d_tris;
    ASUNA_CUDA_CHECK(launch_build_wide(s, m.n_tris, ctx->d_blas_nodes, (uint32_t)m.node_base, (uint32_t)m.tri_base, ctx->scratch,
                                       pl, cost_prim, ctx->d_mesh_lo + i, ctx->d_mesh_hi + i, ctx->d_build_results + i));
  }
  std::vector<WorldJob> jobs;  // must outlive the enqueued copy: the stream is synchronised before this function returns
  if (have_world) {  // the world BLAS over the flattened instances: every instance's triangles to world space, one launch
    uint32_t off = 0, max_n = 0;
    for (uint32_t i : merged) {
      const HostMesh& m = ctx->meshes[ctx->instances[i].mesh];
      jobs.push_back(WorldJob{m.d_vertices, m.d_indices, hinst[i].o2w[0], hinst[i].o2w[1], hinst[i].o2w[2], m.n_tris, off, i,
                              hinst[i].mat_type == 0xFFFFFFFFu ? (uint32_t)kKindLight : kKindMaterial0 + hinst[i].mat_type});
      off += m.n_tris;
      max_n = std::max(max_n, m.n_tris);
    }
    ASUNA_CUDA_CHECK(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(WorldJob), cudaMemcpyHostToDevice, s));
    launch_world_triangles_batched(s, d_jobs, (uint32_t)jobs.size(), max_n, d_soup, ctx->scratch);
    PrimPayload pl;
    pl.soup = d_soup, pl.tris = ctx->d_tris;
    ASUNA_CUDA_CHECK(launch_build_wide(s, world_tris, ctx->d_blas_nodes, (uint32_t)world_node_base, (uint32_t)world_tri_base,
                                       ctx->scratch, pl, cost_prim, ctx->d_mesh_lo + n_mesh, ctx->d_mesh_hi + n_mesh,
                                       ctx->d_build_results + n_mesh));
  }
  // top level over the instance boxes (≙ createTopLevelAS); entering an instance costs a ray transform
  // plus a whole mesh BVH, so leaves are kept to single instances wherever the SAH allows
  launch_instance_boxes(s, ctx->d_instances, ctx->d_tlas_ids, ctx->d_mesh_lo, ctx->d_mesh_hi, n_tlas, ctx->scratch);
  {
    PrimPayload pl;
    pl.leaf_inst = ctx->d_tlas_leaf_inst, pl.prim_ids = ctx->d_tlas_ids;
    ASUNA_CUDA_CHECK(launch_build_wide(s, n_tlas, ctx->d_tlas_nodes, 0, 0, ctx->scratch, pl, 4.0f, nullptr, nullptr,
                                       ctx->d_build_results + n_mesh + 1));
  }
  cudaEventRecord(e1, s);
  tm.step("build enqueue");
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(s));
  ASUNA_CUDA_CHECK(cudaGetLastError());
  tm.step("build on the device");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->stats.build_ms = ms;
  if (out_ms) *out_ms = ms;

  // statistics for asuna_accel_stats: wide nodes in use, primitive slots, SAH cost summed over the BLASes
  {
    std::vector<BuildResult> res(n_mesh + 2);
    ASUNA_CUDA_CHECK(cudaMemcpy(res.data(), ctx->d_build_results, res.size() * sizeof(BuildResult), cudaMemcpyDeviceToHost));
    uint64_t nodes = 0, prims = 0;
    double cost = 0.0;
    for (uint32_t i = 0; i <= n_mesh; i++) nodes += res[i].wide_nodes, prims += res[i].prim_slots, cost += res[i].sah_cost;
    ctx->accel_stats[0] = nodes;
    ctx->accel_stats[1] = prims;
    ctx->accel_stats[2] = res[n_mesh + 1].wide_nodes;
    ctx->accel_stats[3] = (uint64_t)(cost * 1000.0);
  }

  ctx->view.tlas_nodes = ctx->d_tlas_nodes;
  ctx->view.tlas_leaf_inst = ctx->d_tlas_leaf_inst;
  ctx->view.blas_nodes = ctx->d_blas_nodes;
  ctx->view.tris = ctx->d_tris;
  ctx->view.instances = ctx->d_instances;
  ctx->view.meshes = ctx->d_meshes;
  ctx->view.materials = ctx->d_materials;
  ctx->view.lights = ctx->d_lights;
  ctx->view.textures = ctx->d_textures;
  for (int k = 0; k < 3; k++) ctx->view.env[k] = DTexture{ctx->env[k].d_texels, (int)ctx->env[k].w, (int)ctx->env[k].h};
  ctx->view.n_instances = n_inst;
  ctx->view.world_inst = have_world ? n_inst : 0xFFFFFFFFu;
  ctx->view.single_root = single_level ? (uint32_t)world_node_base : 0xFFFFFFFFu;
  free_dev(d_soup);
  if (ctx->scratch.capacity > (4u << 20)) ctx->scratch.release();  // ~330 B per primitive: give large builds' temporaries back
  ctx->view.magic = 0x4B000000u;
  ctx->view.refill_lanes = 8, ctx->view.tri_vote_shift = 3;
  ctx->view.stage_lanes = 16;
  if (const char* t = getenv("ASUNA_STAGE_LANES")) ctx->view.stage_lanes = std::min(std::max(atoi(t), 1), 32);
  if (const char* t = getenv("ASUNA_TUNE")) {  // "refill,shift[,cost_prim x10]" -- traversal tuning experiments
    unsigned a = 8, b = 2, c = 0;
    if (sscanf(t, "%u,%u,%u", &a, &b, &c) >= 1) ctx->view.refill_lanes = std::min(std::max(a, 1u), 32u), ctx->view.tri_vote_shift = std::min(b, 5u);
  }
  ctx->kind_mask = 1u << kKindMiss;
  for (auto& in : ctx->instances)
    ctx->kind_mask |= in.light >= 0 ? (1u << kKindLight) : (1u << (kKindMaterial0 + ctx->materials[in.material].type));
  ctx->may_pass_through = false;
  for (auto& m : ctx->materials)
    if ((m.type == ASUNA_MAT_PBR_METALNESS_ROUGHNESS && (m.opacityTextureId >= 0 || m.specular > 0.f)) ||
        (m.type == ASUNA_MAT_KANG18 && (m.opacityTextureId >= 0 || m.metalness > 0.f)) ||
        (m.type == ASUNA_MAT_DISNEY && (m.opacityTextureId >= 0 || m.rhoSpec[0] > 0.f)))
      ctx->may_pass_through = true;
  // texture / env-table copies ran on the upload stream while the BVH was being built; rendering waits for them
  if (ctx->textures_pending) {
    ASUNA_CUDA_CHECK(cudaEventRecord(ctx->ev_textures, ctx->upload_stream));
    ASUNA_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_textures, 0));
    ctx->textures_pending = false;
  }
  ctx->scene_dirty = false;
  tm.step("statistics read-back, frees");
  return 0;
}

int asuna_set_camera(asuna_ctx* ctx, const AsunaCamera* c) {
  if (!c) return fail(ctx, ASUNA_E_INVALID, "null camera");
  if (c->type != ASUNA_CAMERA_PERSPECTIVE && c->type != ASUNA_CAMERA_OPENCV) return fail(ctx, ASUNA_E_INVALID, "unknown camera type");
  ctx->cam = *c;
  return 0;
}
int asuna_set_sunsky(asuna_ctx* ctx, const AsunaSunSky* s) {
  if (!s) return fail(ctx, ASUNA_E_INVALID, "null sunsky");
  const bool changed = memcmp(&ctx->sunsky, s, sizeof *s) != 0;
  ctx->sunsky = *s;
  if (s->in_use == 1 && (changed || !ctx->sky_valid)) {
    // most of the model depends on the setting only: one device evaluation per setting, not per lookup
    cudaSetDevice(ctx->device);
    if (!ctx->d_sky) ASUNA_CUDA_CHECK(cudaMalloc(&ctx->d_sky, sizeof(SkyPre)));
    launch_sky_prepare(ctx->stream, ctx->sunsky, ctx->d_sky);
    ASUNA_CUDA_CHECK(cudaMemcpyAsync(&ctx->sky, ctx->d_sky, sizeof(SkyPre), cudaMemcpyDeviceToHost, ctx->stream));
    ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->sky_valid = true;
  }
  return 0;
}
int asuna_set_state(asuna_ctx* ctx, const AsunaState* st) {
  if (!st) return fail(ctx, ASUNA_E_INVALID, "null state");
  if (st->spp != 1) return fail(ctx, ASUNA_E_INVALID, "spp must be 1 per frame (reference tracer.cpp:211)");
  if (st->nMultiChannel > ASUNA_NUM_OUTPUT_IMAGES - 1) return fail(ctx, ASUNA_E_INVALID, "more than 8 output channels");
  if (st->maxPathDepth > ASUNA_MAX_ITERS) return fail(ctx, ASUNA_E_INVALID, "maxPathDepth exceeds ASUNA_MAX_ITERS (256)");
  if (st->numLights < 0 || (size_t)st->numLights + 1 > std::max<size_t>(ctx->lights.size(), 1))
    return fail(ctx, ASUNA_E_INVALID, "numLights exceeds the uploaded light table");
  if (st->hasEnvMap == 1 && !ctx->env[0].d_texels) return fail(ctx, ASUNA_E_INVALID, "hasEnvMap set but no env map uploaded");
  ctx->pc = *st;
  return 0;
}
int asuna_reset_frame(asuna_ctx* ctx) {
  ctx->pc.curFrame = -1;
  ctx->have_accum = false;
  return 0;
}
int asuna_set_partition(asuna_ctx* ctx, uint32_t rank, uint32_t world) {
  if (world == 0 || rank >= world) return fail(ctx, ASUNA_E_INVALID, "bad partition");
  ctx->rank = rank, ctx->world = world;
  return 0;
}

static int render_batch(asuna_ctx* ctx, const int* frames, uint32_t n_frames) {
  NvtxRange nvtx("asuna render_batch");
  cudaStream_t s = ctx->stream;
  FrameParams fp = make_frame_params(ctx);
  fp.n_frames = n_frames;
  for (uint32_t i = 0; i < n_frames; i++) fp.frame_ids[i] = frames[i];
  fp.first_is_replace = ctx->have_accum ? 0u : 1u;
  uint32_t n_paths = fp.n_pixels * n_frames;
  int rc = ensure_path_buffers(ctx, n_paths);
  if (rc) return rc;
  ScopedTimer total(ctx, 2);
  ASUNA_CUDA_CHECK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(Counters), s));
  {
    ScopedTimer t(ctx, 1);
    launch_raygen(s, fp, ctx->ps, ctx->out, ctx->d_counters);
  }
  int iter = 0, qsel = 0;
  auto bounce = [&]() {
    NvtxRange nvtx_b("bounce: trace closest / regroup + shade / trace shadow");
    {
      ScopedTimer t(ctx, 0);
      launch_trace_closest(s, ctx->dims, ctx->view, ctx->ps, ctx->d_counters, iter, qsel, ctx->counting);
    }
    {
      ScopedTimer t(ctx, 1);
      ctx->stats.kernel_launches += launch_shade(s, ctx->dims, ctx->view, fp, ctx->ps, ctx->out, ctx->d_counters, iter, qsel,
                                                 ctx->kind_mask, n_paths);
    }
    {
      ScopedTimer t(ctx, 3);
      launch_trace_shadow(s, ctx->dims, ctx->view, ctx->ps, ctx->d_counters, iter);
    }
    ctx->stats.kernel_launches += 2;
    ctx->stats.closest_launches += 1;
    iter++;
    qsel ^= 1;
  };
  int planned = std::min(std::max(ctx->pc.maxPathDepth, 0), ASUNA_MAX_ITERS);
  for (int d = 0; d < planned; d++) bounce();
  // opacity pass-through keeps the depth (brdf_pbr_metalness_roughness.rchit:156-160), so a path may
  // need more iterations than maxPathDepth; only then is a host round trip needed.
  while (ctx->may_pass_through && planned > 0 && iter < ASUNA_MAX_ITERS) {
    uint32_t alive = 0;
    ASUNA_CUDA_CHECK(cudaMemcpyAsync(&alive, &ctx->d_counters->queue[iter], sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    ASUNA_CUDA_CHECK(cudaStreamSynchronize(s));
    if (alive == 0) break;
    if (iter + 1 >= ASUNA_MAX_ITERS)  // the reference would keep looping; say so instead of dropping live paths silently
      return fail(ctx, ASUNA_E_UNSUPPORTED, "opacity pass-through chain exceeds ASUNA_MAX_ITERS bounce iterations");
    bounce();
  }
  {
    ScopedTimer t(ctx, 1);
    launch_accumulate(s, fp, ctx->ps, ctx->out);
  }
  launch_fold_counters(s, ctx->d_counters, ctx->d_totals, iter);
  ctx->stats.kernel_launches += 3;  // raygen, accumulate, fold
  ctx->stats.paths += n_paths;
  ctx->have_accum = true;
  ASUNA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int asuna_render_frames(asuna_ctx* ctx, uint32_t n) {
  cudaSetDevice(ctx->device);
  if (ctx->scene_dirty) return fail(ctx, ASUNA_E_INVALID, "scene changed since the last asuna_build_accel");
  if (ctx->W == 0) return fail(ctx, ASUNA_E_INVALID, "asuna_set_film not called");
  std::vector<int> mine;
  for (uint32_t f = 0; f < n; f++) {
    ctx->pc.curFrame++;  // ≙ incrementFrame(), pipeline_raytrace.cpp:84-86
    if ((uint32_t)ctx->pc.curFrame % ctx->world == ctx->rank) mine.push_back(ctx->pc.curFrame);
  }
  uint32_t batch = std::min<uint32_t>(ctx->max_batch_frames, ASUNA_MAX_BATCH_FRAMES);
  // keep a batch under ~16 M paths
  uint64_t px = (uint64_t)ctx->W * ctx->H;
  while (batch > 1 && px * batch > (16ull << 20)) batch--;
  for (size_t i = 0; i < mine.size(); i += batch) {
    uint32_t nb = (uint32_t)std::min<size_t>(batch, mine.size() - i);
    int rc = render_batch(ctx, mine.data() + i, nb);
    if (rc) return rc;
  }
  return 0;
}

static int pull_totals(asuna_ctx* ctx);

// User-ray launches do not fold into Totals: fetch their overflow flag directly (the stream is idle here).
static int check_user_overflow(asuna_ctx* ctx) {
  uint32_t ovf = 0;
  ASUNA_CUDA_CHECK(cudaMemcpy(&ovf, &ctx->d_counters->stack_overflow, sizeof ovf, cudaMemcpyDeviceToHost));
  if (ovf) return fail(ctx, ASUNA_E_CUDA, "traversal stack overflow: results are incomplete");
  return 0;
}

int asuna_sync(asuna_ctx* ctx) {
  cudaSetDevice(ctx->device);
  int rc = pull_totals(ctx);  // also surfaces a traversal stack overflow
  if (rc) return rc;
  ASUNA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int asuna_read_channel(asuna_ctx* ctx, int ch, float* out) {
  NvtxRange nvtx("asuna_read_channel");
  if (ch < 0 || ch >= ASUNA_NUM_OUTPUT_IMAGES || !out) return fail(ctx, ASUNA_E_INVALID, "bad channel");
  if (ctx->W == 0) return fail(ctx, ASUNA_E_INVALID, "asuna_set_film not called");
  cudaSetDevice(ctx->device);
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(out, ctx->out.img[ch], (size_t)ctx->W * ctx->H * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  collect_timers(ctx);
  return 0;
}

int asuna_read_channel_async(asuna_ctx* ctx, int ch, float* out) {
  NvtxRange nvtx("asuna_read_channel_async");
  if (ch < 0 || ch >= ASUNA_NUM_OUTPUT_IMAGES || !out) return fail(ctx, ASUNA_E_INVALID, "bad channel");
  if (ctx->W == 0) return fail(ctx, ASUNA_E_INVALID, "asuna_set_film not called");
  cudaSetDevice(ctx->device);
  const size_t px = (size_t)ctx->W * ctx->H;
  if (!ctx->read_stream) {
    ASUNA_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->read_stream, cudaStreamNonBlocking));
    ASUNA_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_read_ready, cudaEventDisableTiming));
    ASUNA_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_read_done, cudaEventDisableTiming));
  }
  if (ctx->readback_px < px) {
    if (ctx->read_pending) ASUNA_CUDA_CHECK(cudaEventSynchronize(ctx->ev_read_done));
    free_dev(ctx->d_readback);
    ctx->readback_px = 0;
    ASUNA_CUDA_CHECK(cudaMalloc(&ctx->d_readback, px * sizeof(float4)));
    ctx->readback_px = px;
  }
  // the snapshot buffer is free again once the previous copy out of it has finished (a stream-side wait, not a host one)
  if (ctx->read_pending) ASUNA_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_read_done, 0));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_readback, ctx->out.img[ch], px * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  ASUNA_CUDA_CHECK(cudaEventRecord(ctx->ev_read_ready, ctx->stream));
  ASUNA_CUDA_CHECK(cudaStreamWaitEvent(ctx->read_stream, ctx->ev_read_ready, 0));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(out, ctx->d_readback, px * sizeof(float4), cudaMemcpyDeviceToHost, ctx->read_stream));
  ASUNA_CUDA_CHECK(cudaEventRecord(ctx->ev_read_done, ctx->read_stream));
  ctx->read_pending = true;
  return 0;
}
int asuna_wait_reads(asuna_ctx* ctx) {
  cudaSetDevice(ctx->device);
  if (ctx->read_pending) ASUNA_CUDA_CHECK(cudaEventSynchronize(ctx->ev_read_done));
  ctx->read_pending = false;
  return 0;
}

int asuna_host_alloc(asuna_ctx* ctx, size_t bytes, void** out) {
  if (!out || bytes == 0) return fail(ctx, ASUNA_E_INVALID, "bad host allocation request");
  cudaSetDevice(ctx->device);
  ASUNA_CUDA_CHECK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return 0;
}
int asuna_host_free(asuna_ctx* ctx, void* p) {
  if (!p) return 0;
  ASUNA_CUDA_CHECK(cudaFreeHost(p));
  return 0;
}

int asuna_export_partial(asuna_ctx* ctx, void** out) {
  if (ctx->W == 0) return fail(ctx, ASUNA_E_INVALID, "asuna_set_film not called");
  cudaSetDevice(ctx->device);
  // Stream-ordered: the buffer is complete once the work queued on asuna_stream_handle so far has run.  A collective
  // issued on ANOTHER stream must wait for that stream (an event, torch's wait_stream); no host sync is needed here.
  launch_export_partial(ctx->stream, ctx->out, ctx->d_partial, ctx->W * ctx->H, ctx->have_accum ? 1 : 0);
  ASUNA_CUDA_CHECK(cudaGetLastError());
  *out = ctx->d_partial;
  return 0;
}
int asuna_import_partial(asuna_ctx* ctx) {
  if (ctx->W == 0) return fail(ctx, ASUNA_E_INVALID, "asuna_set_film not called");
  cudaSetDevice(ctx->device);
  launch_import_partial(ctx->stream, ctx->out, ctx->d_partial, ctx->W * ctx->H);  // stream-ordered, like the export
  ASUNA_CUDA_CHECK(cudaGetLastError());
  ctx->have_accum = true;
  return 0;
}
int asuna_post_process(asuna_ctx* ctx, const AsunaPost* tm, float* out) {
  NvtxRange nvtx("asuna_post_process");
  if (!tm || !out) return fail(ctx, ASUNA_E_INVALID, "null argument");
  if (ctx->W == 0) return fail(ctx, ASUNA_E_INVALID, "asuna_set_film not called");
  if (tm->tmType >= ASUNA_TM_NUM) return fail(ctx, ASUNA_E_INVALID, "unknown tone mapper");
  if (tm->tmType == ASUNA_TM_CUSTOM && (tm->autoExposure & 2))
    return fail(ctx, ASUNA_E_UNSUPPORTED, "local auto-exposure (autoExposure bit 1) reads the display image's mip chain and is not available offline");
  cudaSetDevice(ctx->device);
  launch_post_process(ctx->stream, ctx->out.img[0], ctx->d_ldr, ctx->W, ctx->H, *tm, ctx->d_post_sums);
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(out, ctx->d_ldr, (size_t)ctx->W * ctx->H * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ASUNA_CUDA_CHECK(cudaGetLastError());
  collect_timers(ctx);
  return 0;
}
int asuna_channel_device_ptr(asuna_ctx* ctx, int ch, void** out) {
  if (ch < 0 || ch >= ASUNA_NUM_OUTPUT_IMAGES || !out || ctx->W == 0) return fail(ctx, ASUNA_E_INVALID, "bad channel");
  *out = ctx->out.img[ch];
  return 0;
}

static int pull_totals(asuna_ctx* ctx) {
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_totals, ctx->d_totals, sizeof(Totals), cudaMemcpyDeviceToHost, ctx->stream));
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  collect_timers(ctx);
  ctx->stats.closest_rays = ctx->h_totals->closest_rays;
  ctx->stats.shadow_rays = ctx->h_totals->shadow_rays;
  ctx->stats.incoherent_closest_rays = ctx->h_totals->incoherent_rays;
  ctx->stats.node_visits = ctx->h_totals->node_visits;
  ctx->stats.tri_tests = ctx->h_totals->tri_tests;
  if (ctx->h_totals->stack_overflow) return fail(ctx, ASUNA_E_CUDA, "traversal stack overflow: results are incomplete");
  return 0;
}

int asuna_get_stats(asuna_ctx* ctx, AsunaStats* out) {
  cudaSetDevice(ctx->device);
  int rc = pull_totals(ctx);
  *out = ctx->stats;
  return rc;
}
int asuna_stream_handle(asuna_ctx* ctx, void** out) {
  *out = (void*)ctx->stream;
  return 0;
}
int asuna_set_counting(asuna_ctx* ctx, int on) {
  ctx->counting = on != 0;
  return 0;
}
int asuna_set_profiling(asuna_ctx* ctx, int on) {
  ctx->profiling = on != 0;
  return 0;
}
int asuna_reset_stats(asuna_ctx* ctx) {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  collect_timers(ctx);
  float b = ctx->stats.build_ms;
  ctx->stats = AsunaStats{};
  ctx->stats.build_ms = b;
  ASUNA_CUDA_CHECK(cudaMemsetAsync(ctx->d_totals, 0, sizeof(Totals), ctx->stream));
  return 0;
}

int asuna_trace_rays(asuna_ctx* ctx, const float* rays, uint32_t n, float* tuv, uint32_t* ip) {
  if (ctx->scene_dirty) return fail(ctx, ASUNA_E_INVALID, "scene not built");
  if (n == 0) return 0;
  cudaSetDevice(ctx->device);
  int rc = upload_user_rays(ctx, rays, n);
  if (rc) return rc;
  ASUNA_CUDA_CHECK(cudaMemsetAsync(&ctx->d_counters->stack_overflow, 0, sizeof(uint32_t), ctx->stream));
  const bool timed = getenv("ASUNA_TIME_USER_RAYS") != nullptr;  // developer probe (tools/coherence_probe.py): kernel ms on stderr
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed) cudaEventCreate(&e0), cudaEventCreate(&e1), cudaEventRecord(e0, ctx->stream);
  launch_trace_user(ctx->stream, ctx->dims, ctx->view, ctx->d_user_rays, n, ctx->d_user_tuv, ctx->d_user_ip, nullptr, ctx->d_counters);
  if (timed) {
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "asuna_trace_rays: %u rays in %.4f ms = %.1f Mrays/s\n", n, ms, n / ms / 1e3);
    cudaEventDestroy(e0), cudaEventDestroy(e1);
  }
  if (tuv) ASUNA_CUDA_CHECK(cudaMemcpyAsync(tuv, ctx->d_user_tuv, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ip, ctx->d_user_ip, (size_t)n * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ASUNA_CUDA_CHECK(cudaGetLastError());
  return check_user_overflow(ctx);
}
int asuna_occlusion_rays(asuna_ctx* ctx, const float* rays, uint32_t n, uint8_t* occ) {
  if (ctx->scene_dirty) return fail(ctx, ASUNA_E_INVALID, "scene not built");
  if (n == 0) return 0;
  cudaSetDevice(ctx->device);
  int rc = upload_user_rays(ctx, rays, n);
  if (rc) return rc;
  ASUNA_CUDA_CHECK(cudaMemsetAsync(&ctx->d_counters->stack_overflow, 0, sizeof(uint32_t), ctx->stream));
  launch_trace_user(ctx->stream, ctx->dims, ctx->view, ctx->d_user_rays, n, nullptr, nullptr, ctx->d_user_occ, ctx->d_counters);
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(occ, ctx->d_user_occ, n, cudaMemcpyDeviceToHost, ctx->stream));
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ASUNA_CUDA_CHECK(cudaGetLastError());
  return check_user_overflow(ctx);
}
int asuna_trace_primary(asuna_ctx* ctx, uint32_t* ip, float* t) {
  if (ctx->scene_dirty) return fail(ctx, ASUNA_E_INVALID, "scene not built");
  if (ctx->W == 0) return fail(ctx, ASUNA_E_INVALID, "asuna_set_film not called");
  cudaSetDevice(ctx->device);
  uint32_t n = ctx->W * ctx->H;
  int rc = upload_user_rays(ctx, nullptr, n);
  if (rc) return rc;
  FrameParams fp = make_frame_params(ctx);
  ASUNA_CUDA_CHECK(cudaMemsetAsync(&ctx->d_counters->stack_overflow, 0, sizeof(uint32_t), ctx->stream));
  launch_primary_rays(ctx->stream, fp, ctx->d_user_rays);
  launch_trace_user(ctx->stream, ctx->dims, ctx->view, ctx->d_user_rays, n, ctx->d_user_tuv, ctx->d_user_ip, nullptr, ctx->d_counters);
  std::vector<float> tuv((size_t)n * 3);
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(tuv.data(), ctx->d_user_tuv, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ip, ctx->d_user_ip, (size_t)n * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ASUNA_CUDA_CHECK(cudaGetLastError());
  for (uint32_t i = 0; i < n; i++) t[i] = tuv[3 * (size_t)i];
  return check_user_overflow(ctx);
}
int asuna_accel_stats(asuna_ctx* ctx, uint64_t out[4]) {
  if (ctx->scene_dirty) return fail(ctx, ASUNA_E_INVALID, "scene not built");
  memcpy(out, ctx->accel_stats, sizeof ctx->accel_stats);
  return 0;
}

// Test hook (not part of the reference-facing surface): sorts n (key, value) pairs given as host
// arrays with the builder's radix sort, so the sort can be checked bit-exactly on its own.
// Developer probe: warp-loop occupancy counters of the instrumented traversal (see Counters::lane_stats); valid after
// asuna_get_stats has fetched the totals.
int asuna_debug_lane_stats(asuna_ctx* ctx, uint64_t out[6]) {
  for (int k = 0; k < 6; k++) out[k] = ctx->h_totals->lane_stats[k];
  return 0;
}
// Test hook: the node and triangle pools of the mesh-level BVHs as built (tests walk the tree on the host: child boxes
// inside parent boxes, every primitive referenced once, quantised boxes contain their triangles).
// sizes = {allocated nodes, triangle slots, root of the single-level world BVH or ~0}.
int asuna_debug_accel_sizes(asuna_ctx* ctx, uint64_t sizes[3]) {
  if (ctx->scene_dirty) return fail(ctx, ASUNA_E_INVALID, "scene not built");
  sizes[0] = ctx->pool_nodes, sizes[1] = ctx->pool_tris, sizes[2] = ctx->view.single_root;
  return 0;
}
int asuna_debug_download_accel(asuna_ctx* ctx, void* nodes_out, void* tris_out) {
  if (ctx->scene_dirty) return fail(ctx, ASUNA_E_INVALID, "scene not built");
  cudaSetDevice(ctx->device);
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (nodes_out) ASUNA_CUDA_CHECK(cudaMemcpy(nodes_out, ctx->d_blas_nodes, ctx->pool_nodes * sizeof(WideNode), cudaMemcpyDeviceToHost));
  if (tris_out) ASUNA_CUDA_CHECK(cudaMemcpy(tris_out, ctx->d_tris, ctx->pool_tris * sizeof(TriSlot), cudaMemcpyDeviceToHost));
  return 0;
}

// Test hook: ONE shade-kernel invocation per caller-made path (tests/test_gpu_ref_parity.py).  Mirrors oracle_shade_probes /
// refglsl_shade_probe (oracle/refbuild/ref_bridge.h, same ShadeProbe layout): the payload as it enters a closest-hit or
// miss shader goes into path slot i, the hit record is given (inst = 0xFFFFFFFF: miss), the regroup + per-kind shade
// kernels of one bounce run, and the payload is read back from the path state, the next queue (stop), the shadow queue
// (direct-light record) and the AOV planes.  What the wavefront form does not keep is returned as it came in: the RNG
// state / next ray / throughput of a path that stopped, the radiance of a zero NEE contribution (skip = 1, A.3-4).
struct AsunaShadeProbe {
  float ray_o[3], ray_d[3], radiance[3], throughput[3];
  uint32_t depth, seed, stop;
  float brec_d[3], brec_pdf;
  uint32_t brec_flags;
  float drec_radiance[3], drec_dist, drec_o[3], drec_d[3];
  uint32_t drec_skip;
  float channel[8][3];
};
int asuna_debug_shade_probes(asuna_ctx* ctx, uint32_t n, const uint32_t* inst, const uint32_t* prim, const float* b1,
                             const float* b2, AsunaShadeProbe* q) {
  if (ctx->scene_dirty) return fail(ctx, ASUNA_E_INVALID, "scene not built");
  if (n == 0) return 0;
  if (ctx->W == 0 || n > ctx->W * ctx->H) return fail(ctx, ASUNA_E_INVALID, "shade probes need a film of at least n pixels");
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  int rc = ensure_path_buffers(ctx, ctx->W * ctx->H);
  if (rc) return rc;
  std::vector<float4> ro(n), rd(n), th(n), ra(n);
  std::vector<uint4> hit(n);
  std::vector<uint32_t> queue(n);
  std::vector<uint8_t> kind(n);
  for (uint32_t i = 0; i < n; i++) {
    const AsunaShadeProbe& p = q[i];
    uint32_t seed = p.seed, packed = (p.depth & 0xFFFFu) | (p.brec_flags << 16);
    float fs, fp_;
    memcpy(&fs, &seed, 4), memcpy(&fp_, &packed, 4);
    ro[i] = make_float4(p.ray_o[0], p.ray_o[1], p.ray_o[2], fs);
    rd[i] = make_float4(p.ray_d[0], p.ray_d[1], p.ray_d[2], p.brec_pdf);
    th[i] = make_float4(p.throughput[0], p.throughput[1], p.throughput[2], fp_);
    ra[i] = make_float4(p.radiance[0], p.radiance[1], p.radiance[2], 0.f);
    uint32_t ub1, ub2;
    memcpy(&ub1, &b1[i], 4), memcpy(&ub2, &b2[i], 4);
    hit[i] = make_uint4(ub1, ub2, inst[i], prim[i]);
    queue[i] = i;
    if (inst[i] == 0xFFFFFFFFu) {
      kind[i] = (uint8_t)kKindMiss;
    } else {
      if (inst[i] >= ctx->instances.size()) return fail(ctx, ASUNA_E_INVALID, "probe refers to unknown instance");
      const HostInstance& in = ctx->instances[inst[i]];
      if (prim[i] >= ctx->meshes[in.mesh].n_tris) return fail(ctx, ASUNA_E_INVALID, "probe refers to unknown primitive");
      kind[i] = (uint8_t)(in.light >= 0 ? (uint32_t)kKindLight : kKindMaterial0 + ctx->materials[in.material].type);
    }
  }
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->ps.ray_o, ro.data(), n * sizeof(float4), cudaMemcpyHostToDevice, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->ps.ray_d, rd.data(), n * sizeof(float4), cudaMemcpyHostToDevice, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->ps.thr, th.data(), n * sizeof(float4), cudaMemcpyHostToDevice, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->ps.rad, ra.data(), n * sizeof(float4), cudaMemcpyHostToDevice, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->ps.hit, hit.data(), n * sizeof(uint4), cudaMemcpyHostToDevice, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->ps.queue[0], queue.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ctx->ps.kind, kind.data(), n, cudaMemcpyHostToDevice, s));
  ASUNA_CUDA_CHECK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(Counters), s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(&ctx->d_counters->queue[0], &n, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  const size_t plane = (size_t)ctx->W * ctx->H * sizeof(float4);
  for (int c = 1; c < ASUNA_NUM_OUTPUT_IMAGES - 1; c++) ASUNA_CUDA_CHECK(cudaMemsetAsync(ctx->out.img[c], 0, plane, s));
  FrameParams fp = make_frame_params(ctx);
  fp.n_frames = 1;
  fp.frame_ids[0] = 0;  // frame 0: AOVs are stored for depth-1 probes, at pixel = probe index
  fp.first_is_replace = 1;
  fp.pc.curFrame = 0;
  fp.pc.maxPathDepth = 1000;  // "stop" below is then the shader's own decision, never the depth limit
  uint32_t mask = 1u << kKindMiss;
  for (uint32_t i = 0; i < n; i++) mask |= 1u << kind[i];
  launch_shade(s, ctx->dims, ctx->view, fp, ctx->ps, ctx->out, ctx->d_counters, 0, 0, mask, n);
  uint32_t n_next = 0, n_shadow = 0;
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(&n_next, &ctx->d_counters->queue[1], sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(&n_shadow, &ctx->d_counters->shadow[0], sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ro.data(), ctx->ps.ray_o, n * sizeof(float4), cudaMemcpyDeviceToHost, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(rd.data(), ctx->ps.ray_d, n * sizeof(float4), cudaMemcpyDeviceToHost, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(th.data(), ctx->ps.thr, n * sizeof(float4), cudaMemcpyDeviceToHost, s));
  ASUNA_CUDA_CHECK(cudaMemcpyAsync(ra.data(), ctx->ps.rad, n * sizeof(float4), cudaMemcpyDeviceToHost, s));
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(s));
  ASUNA_CUDA_CHECK(cudaGetLastError());
  if (n_next > n || n_shadow > n) return fail(ctx, ASUNA_E_CUDA, "shade probe: queue counters out of range");
  std::vector<uint32_t> next(n_next);
  std::vector<float4> so(n_shadow), sd(n_shadow), sl(n_shadow);
  if (n_next) ASUNA_CUDA_CHECK(cudaMemcpy(next.data(), ctx->ps.queue[1], n_next * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (n_shadow) {
    ASUNA_CUDA_CHECK(cudaMemcpy(so.data(), ctx->ps.sh_o, n_shadow * sizeof(float4), cudaMemcpyDeviceToHost));
    ASUNA_CUDA_CHECK(cudaMemcpy(sd.data(), ctx->ps.sh_d, n_shadow * sizeof(float4), cudaMemcpyDeviceToHost));
    ASUNA_CUDA_CHECK(cudaMemcpy(sl.data(), ctx->ps.sh_l, n_shadow * sizeof(float4), cudaMemcpyDeviceToHost));
  }
  std::vector<float4> aov((size_t)n);
  for (uint32_t i = 0; i < n; i++) q[i].stop = 1, q[i].drec_skip = 1;
  for (uint32_t k : next) {
    if (k >= n) return fail(ctx, ASUNA_E_CUDA, "shade probe: bad slot in the next queue");
    AsunaShadeProbe& p = q[k];
    p.stop = 0;
    uint32_t packed, seed;
    memcpy(&packed, &th[k].w, 4), memcpy(&seed, &ro[k].w, 4);
    p.ray_o[0] = ro[k].x, p.ray_o[1] = ro[k].y, p.ray_o[2] = ro[k].z, p.seed = seed;
    p.ray_d[0] = p.brec_d[0] = rd[k].x, p.ray_d[1] = p.brec_d[1] = rd[k].y, p.ray_d[2] = p.brec_d[2] = rd[k].z;
    p.brec_pdf = rd[k].w;
    p.throughput[0] = th[k].x, p.throughput[1] = th[k].y, p.throughput[2] = th[k].z;
    p.depth = (packed & 0xFFFFu) - 1u;  // the kernel stores rgen's depth++ already
    p.brec_flags = packed >> 16;
  }
  for (uint32_t i = 0; i < n; i++) q[i].radiance[0] = ra[i].x, q[i].radiance[1] = ra[i].y, q[i].radiance[2] = ra[i].z;
  for (uint32_t j = 0; j < n_shadow; j++) {
    uint32_t k;
    memcpy(&k, &sd[j].w, 4);
    if (k >= n) return fail(ctx, ASUNA_E_CUDA, "shade probe: bad slot in the shadow queue");
    AsunaShadeProbe& p = q[k];
    p.drec_skip = 0;
    p.drec_radiance[0] = sl[j].x, p.drec_radiance[1] = sl[j].y, p.drec_radiance[2] = sl[j].z;
    p.drec_o[0] = so[j].x, p.drec_o[1] = so[j].y, p.drec_o[2] = so[j].z;
    p.drec_d[0] = sd[j].x, p.drec_d[1] = sd[j].y, p.drec_d[2] = sd[j].z;
    p.drec_dist = so[j].w + 2.0f * 0.001f;  // the queue holds tmax = dist - 2 EPS (rgen:119)
  }
  for (uint32_t c = 0; c < 7; c++) {
    ASUNA_CUDA_CHECK(cudaMemcpy(aov.data(), ctx->out.img[c + 1], (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; i++) q[i].channel[c][0] = aov[i].x, q[i].channel[c][1] = aov[i].y, q[i].channel[c][2] = aov[i].z;
  }
  ctx->have_accum = false;
  return 0;
}
int asuna_debug_radix_sort(asuna_ctx* ctx, uint64_t* keys, uint32_t* vals, uint32_t n) {
  if (n == 0) return 0;
  cudaSetDevice(ctx->device);
  uint64_t* dk = nullptr;
  uint32_t* dv = nullptr;
  ASUNA_CUDA_CHECK(cudaMalloc(&dk, (size_t)n * sizeof(uint64_t)));
  ASUNA_CUDA_CHECK(cudaMalloc(&dv, (size_t)n * sizeof(uint32_t)));
  ASUNA_CUDA_CHECK(cudaMemcpy(dk, keys, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice));
  ASUNA_CUDA_CHECK(cudaMemcpy(dv, vals, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  ASUNA_CUDA_CHECK(radix_sort_pairs(ctx->stream, dk, dv, n, ctx->scratch));
  ASUNA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ASUNA_CUDA_CHECK(cudaMemcpy(keys, dk, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  ASUNA_CUDA_CHECK(cudaMemcpy(vals, dv, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  cudaFree(dk);
  cudaFree(dv);
  return 0;
}

}  // extern "C"
