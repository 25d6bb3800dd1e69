#include "tracer.h"

#include <cuda_runtime.h>
#include <nccl.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <thread>

namespace asuna_host {

namespace {
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void check(asuna_ctx* ctx, int rc, const char* what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + " failed: " + (ctx ? asuna_last_error(ctx) : "no context"));
}
std::string exe_dir() {
  char buf[4096];
  ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
  if (n <= 0) return "./";
  buf[n] = 0;
  std::string p(buf);
  return p.substr(0, p.find_last_of('/') + 1);
}
}  // namespace

// Host-side tone mappers for the image-conversion hook (main.cpp --convert) only; rendered shots are tone-mapped on the
// device by asuna_post_process.
void tonemap(const std::string& name, int n, const float* hdr, float* out) {
  auto to_srgb = [](float c) { return std::pow(c, 1.0f / 2.2f); };  // linearTosRGB, utils/tonemapping.glsl:30-33
  for (int i = 0; i < n; i++) {
    const float* h = hdr + 4 * (size_t)i;
    float* o = out + 4 * (size_t)i;
    o[3] = h[3];
    for (int c = 0; c < 3; c++) {
      float x = h[c], y;
      if (name == "none") y = x;
      else if (name == "gamma") y = to_srgb(x / (1 + x / 1.5f));
      else if (name == "reinhard" || name == "filmic") {  // toneMapHejlRichard == the Filmic branch
        float v = std::fmax(0.0f, x - 0.004f);
        y = (v * (6.2f * v + 0.5f)) / (v * (6.2f * v + 1.7f) + 0.06f);
      } else if (name == "Aces") {
        float v = (x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f);
        y = to_srgb(std::fmin(std::fmax(v, 0.0f), 1.0f));
      } else if (name == "pbrt") {
        y = x < 0.0031308f ? 12.92f * x : 1.055f * std::pow(x, 1.0f / 2.4f) - 0.055f;
      } else
        throw std::runtime_error("tone mapper [" + name + "] needs the interactive auto-exposure state and is not available offline");
      o[c] = y;
    }
  }
}

Tracer::~Tracer() {
  for (auto& w : m_writers)
    if (w.valid()) w.wait();
  for (void* c : m_comms)
    if (c) ncclCommDestroy((ncclComm_t)c);
  for (asuna_ctx* c : m_ctx)
    if (c) asuna_destroy(c);
}

void Tracer::init() {
  m_scene = Scene::from_json_file(m_tis.scenefile);
  fprintf(stderr, "[info] Scene: %zu shots, %zu instances, %zu meshes, %zu materials, %zu lights, %zu textures\n", m_scene.shots.size(),
          m_scene.instances.size(), m_scene.meshes.size(), m_scene.materials.size() - 1, m_scene.lights.size() - 1,
          m_scene.textures.size() - 1);
  int n = std::max(1, m_tis.n_gpus);
  m_ctx.assign(n, nullptr);
  std::vector<int> devs(n);
  for (int g = 0; g < n; g++) {
    devs[g] = m_tis.gpu_id + g;
    int rc = asuna_create(&m_ctx[g], devs[g]);
    if (rc < 0) throw std::runtime_error("asuna_create failed for GPU " + std::to_string(devs[g]) + " (no CUDA device? there is no CPU fallback)");
  }
  // every GPU holds the whole scene and builds its own acceleration structure (deterministic builder)
  std::vector<std::thread> th;
  std::vector<std::string> err(n);
  std::vector<float> ms(n, 0.f);
  for (int g = 0; g < n; g++)
    th.emplace_back([&, g] {
      try {
        ms[g] = m_scene.upload(m_ctx[g]);
        if (!m_tis.split_shots)  // frame ranges of every shot; with --split shots each GPU renders whole shots
          check(m_ctx[g], asuna_set_partition(m_ctx[g], (uint32_t)g, (uint32_t)n), "asuna_set_partition");
      } catch (const std::exception& e) {
        err[g] = e.what();
      }
    });
  for (auto& t : th) t.join();
  for (auto& e : err)
    if (!e.empty()) throw std::runtime_error(e);
  m_build_ms = ms[0];
  if (n > 1 && !m_tis.split_shots) {
    std::vector<ncclComm_t> comms(n);
    if (ncclCommInitAll(comms.data(), n, devs.data()) != ncclSuccess) throw std::runtime_error("ncclCommInitAll failed");
    for (auto c : comms) m_comms.push_back((void*)c);
  }
}

namespace {
// One host thread per GPU: asuna_render_frames enqueues a whole shot (and, for opacity pass-through scenes, waits on
// the device between bounces), so driving N contexts from one thread would serialise the GPUs.
template <class F>
void for_each_gpu(int n, F&& fn) {
  if (n == 1) {
    fn(0);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::string> err((size_t)n);
  for (int g = 0; g < n; g++)
    th.emplace_back([&, g] {
      try {
        fn(g);
      } catch (const std::exception& e) {
        err[(size_t)g] = e.what();
      }
    });
  for (auto& t : th) t.join();
  for (auto& e : err)
    if (!e.empty()) throw std::runtime_error(e);
}
}  // namespace

std::vector<ShotReport> Tracer::run() {
  std::vector<ShotReport> reports;
  const int n = (int)m_ctx.size();
  if (n > 1 && m_tis.split_shots) {
    // replicas: shot s goes to GPU s % n as a whole -- no communication; a group of n shots is queued on
    // the n GPUs at once (asuna_render_frames is asynchronous), then read back and saved one after the other
    for (size_t first = 0; first < m_scene.shots.size(); first += (size_t)n) {
      const int cnt = (int)std::min<size_t>((size_t)n, m_scene.shots.size() - first);
      double t0 = now_ms();
      std::vector<int> tot(cnt, 0);
      for_each_gpu(cnt, [&](int g) {
        tot[g] = m_scene.begin_shot(m_ctx[g], first + (size_t)g);
        check(m_ctx[g], asuna_render_frames(m_ctx[g], (uint32_t)tot[g]), "asuna_render_frames");
        check(m_ctx[g], asuna_sync(m_ctx[g]), "asuna_sync");
      });
      double render_ms = now_ms() - t0;
      for (int g = 0; g < cnt; g++) {
        ShotReport rep;
        rep.shot = (int)first + g, rep.spp = tot[g], rep.render_ms = render_ms / cnt;
        double t1 = now_ms();
        save_shot(rep.shot, g);
        rep.save_ms = now_ms() - t1;
        fprintf(stderr, "[info] shot %04d on GPU %d: %d spp, group of %d shots in %.1f ms, read back in %.1f ms (files are written in the background)\n", rep.shot,
                m_tis.gpu_id + g, tot[g], cnt, render_ms, rep.save_ms);
        reports.push_back(rep);
      }
    }
    drain_writers(0);
    return reports;
  }
  for (size_t shot = 0; shot < m_scene.shots.size(); shot++) {
    ShotReport rep;
    rep.shot = (int)shot;
    double t0 = now_ms();
    std::vector<int> tots((size_t)n, 0);
    std::vector<void*> part((size_t)n, nullptr), stream((size_t)n, nullptr);
    // every context walks all `tot` frames and renders the ones its partition owns; the partial sums are exported in
    // stream order right behind the last frame (no host synchronisation anywhere before the read-back)
    for_each_gpu(n, [&](int g) {
      tots[(size_t)g] = m_scene.begin_shot(m_ctx[g], shot);
      check(m_ctx[g], asuna_render_frames(m_ctx[g], (uint32_t)tots[(size_t)g]), "asuna_render_frames");
      if (n > 1) {
        check(m_ctx[g], asuna_export_partial(m_ctx[g], &part[(size_t)g]), "asuna_export_partial");
        check(m_ctx[g], asuna_stream_handle(m_ctx[g], &stream[(size_t)g]), "asuna_stream_handle");
      }
    });
    const int tot = tots[0];
    rep.spp = tot;
    if (n > 1) {
      // one NCCL sum to GPU 0 on the contexts' own streams: ordered after the export kernels, before the import kernel
      size_t count = (size_t)m_scene.camera.width * m_scene.camera.height * 4;
      ncclGroupStart();
      for (int g = 0; g < n; g++) {
        cudaSetDevice(m_tis.gpu_id + g);
        ncclReduce(part[(size_t)g], part[(size_t)g], count, ncclFloat, ncclSum, 0, (ncclComm_t)m_comms[(size_t)g], (cudaStream_t)stream[(size_t)g]);
      }
      if (ncclGroupEnd() != ncclSuccess) throw std::runtime_error("ncclReduce failed");
      check(m_ctx[0], asuna_import_partial(m_ctx[0]), "asuna_import_partial");
    }
    for (int g = 0; g < n; g++) check(m_ctx[g], asuna_sync(m_ctx[g]), "asuna_sync");
    rep.render_ms = now_ms() - t0;
    double t1 = now_ms();
    save_shot((int)shot);
    rep.save_ms = now_ms() - t1;
    fprintf(stderr, "[info] shot %04d: %d spp in %.1f ms (%.1f M samples/s), read back in %.1f ms (files are written in the background)\n", (int)shot, tot, rep.render_ms,
            (double)tot * m_scene.camera.width * m_scene.camera.height / rep.render_ms / 1e3, rep.save_ms);
    reports.push_back(rep);
  }
  double tw = now_ms();
  drain_writers(0);
  if (!reports.empty()) fprintf(stderr, "[info] image writers finished %.1f ms after the last shot\n", now_ms() - tw);
  return reports;
}

// callSavingImage, src/tracer/tracer.cpp:266-291
void Tracer::save_shot(int shot_id, int gpu) {
  char name[4096];
  const OutputOptions& o = m_scene.output;
  if (!o.render_result) return;
  if (o.hdr) {
    snprintf(name, sizeof name, "%s_shot_%04d.exr", m_tis.outputname.c_str(), shot_id);
    save_buffer(name, 0, gpu);
  } else {
    snprintf(name, sizeof name, "%s_shot_%04d.png", m_tis.outputname.c_str(), shot_id);
    save_buffer(name, -1, gpu);
  }
  for (uint32_t cid = 0; cid < m_scene.state.nMultiChannel; cid++) {
    bool ldr = cid < o.channel_ldr.size() && o.channel_ldr[cid];
    snprintf(name, sizeof name, "%s_shot_%04d_channel_%04d.%s", m_tis.outputname.c_str(), shot_id, (int)cid, ldr ? "png" : "exr");
    save_buffer(name, (int)cid + 1, gpu);
  }
}

// saveBufferToImage, src/tracer/tracer.cpp:347-395
void Tracer::save_buffer(const std::string& path_in, int channel_id, int gpu) {
  std::string path = path_in;
  if (path.empty() || path[0] != '/') path = exe_dir() + path;  // relative paths are taken from the executable's directory
  const int w = m_scene.camera.width, h = m_scene.camera.height;
  std::vector<float> data((size_t)w * h * 4);
  if (channel_id < 0) {
    // the post-processed colour: PipelinePost on the device (asuna_post_process, all seven tone mappers of
    // post.idle.frag; no denoiser on this path).  Defaults of the post state: src/core/state.h:46-58.
    static const char* names[] = {"none", "gamma", "reinhard", "Aces", "filmic", "pbrt", "custom"};
    AsunaPost post{};
    post.brightness = post.contrast = post.saturation = post.avgLum = post.zoom = 1.0f;
    post.renderingRatio[0] = post.renderingRatio[1] = 1.0f;
    post.Ywhite = post.key = 0.5f;
    post.tmType = ASUNA_TM_NONE;  // loader.cpp:206-225: an unknown name falls back to none
    for (uint32_t k = 0; k < ASUNA_TM_NUM; k++)
      if (m_scene.output.tone_mapping == names[k]) post.tmType = k;
    check(m_ctx[gpu], asuna_post_process(m_ctx[gpu], &post, data.data()), "asuna_post_process");
  } else {
    check(m_ctx[gpu], asuna_read_channel(m_ctx[gpu], channel_id, data.data()), "asuna_read_channel");
  }
  // Encoding and writing happen on worker threads while the next shot renders (a 1080p PNG is ~30 ms of deflate even in
  // bands; four files per shot used to be most of a multi-shot job's wall time).  The pixels are already in host memory
  // and owned by the task; drain_writers() bounds the number of images in flight and rethrows a writer's exception.
  auto pixels = std::make_shared<std::vector<float>>(std::move(data));
  const bool f32 = m_tis.output_f32;
  if (!m_tis.output_scanline || channel_id == 0 || channel_id == -1) {
    // (the reference also takes the scanline branch for channel -1, where its valid-pixel list is stale or empty;
    //  the tone-mapped image is written normally here)
    m_writers.push_back(std::async(std::launch::async, [=] {
      if (f32) write_npy_f32(path + ".npy", {(size_t)h, (size_t)w, 4}, pixels->data());
      write_image(path, w, h, pixels->data());
    }));
    drain_writers(8);
    return;
  }
  // --output_scanline: channels >= 1 are stored as an (n_valid, 3) float32 NPY under the image's name; the valid
  // pixel list is taken from channel 1 (pixels whose .z == 1, replaced by the pixel index) and reused afterwards
  if (channel_id == 1) {
    m_valid_pixel_index.clear();
    for (int idx = 0; idx < w * h; idx++)
      if ((*pixels)[4 * (size_t)idx + 2] == 1.0f) {
        m_valid_pixel_index.push_back(idx);
        (*pixels)[4 * (size_t)idx + 2] = (float)idx;
      }
  }
  auto valid = std::make_shared<std::vector<float>>();
  valid->reserve(m_valid_pixel_index.size() * 3);
  for (int idx : m_valid_pixel_index)
    for (int c = 0; c < 3; c++) valid->push_back((*pixels)[4 * (size_t)idx + c]);
  const size_t n_valid = m_valid_pixel_index.size();
  m_writers.push_back(std::async(std::launch::async, [=] {
    if (f32) write_npy_f32(path + ".npy", {(size_t)h, (size_t)w, 4}, pixels->data());
    write_npy_f32(path, {n_valid, 3}, valid->data());
  }));
  drain_writers(8);
}

// Waits until at most `keep` image writers are still running (0 = all of them), oldest first.
void Tracer::drain_writers(size_t keep) {
  while (m_writers.size() > keep) {
    m_writers.front().get();
    m_writers.pop_front();
  }
}

}  // namespace asuna_host
