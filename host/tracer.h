// Offline render driver: the equivalent of Tracer::init / runOffline / callSavingImage of the reference
// (src/tracer/tracer.cpp:19-61, 177-291, 347-395) on top of the C ABI.  One context per GPU; with several GPUs the
// frames of a shot are split f % N == rank (scene replicated) and the (sum w L, sum w) planes are combined with one
// NCCL reduce to GPU 0 per shot.
#pragma once
#include <deque>
#include <future>
#include <string>
#include <vector>

#include "scene.h"

namespace asuna_host {

struct TracerSettings {  // TracerInitSettings, src/tracer/tracer.h + src/main.cpp:18-24
  std::string scenefile, outputname;
  int gpu_id = 0;
  int n_gpus = 1;
  bool split_shots = false;   // extension: with n_gpus > 1 give every GPU whole shots (replicas, no reduce) instead of frame ranges
  bool offline = true;
  bool output_scanline = false;
  bool output_f32 = false;    // extension: also write <image>.npy with the full float32 RGBA plane
  std::string report;         // extension: JSON performance report
};

struct ShotReport {
  int shot = 0, spp = 0;
  double render_ms = 0, save_ms = 0;
};

// Tone mappers of src/shaders/post.idle.frag:76-133 ("custom" needs the GUI's auto-exposure state and is rejected).
void tonemap(const std::string& name, int n_pixels, const float* hdr_rgba, float* out_rgba);

class Tracer {
 public:
  explicit Tracer(const TracerSettings& s) : m_tis(s) {}
  ~Tracer();
  void init();                         // load the scene, create the contexts, upload, build the acceleration structure
  std::vector<ShotReport> run();       // runOffline: every shot, every spp, save
  const Scene& scene() const { return m_scene; }
  float build_ms() const { return m_build_ms; }

 private:
  void save_shot(int shot_id, int gpu = 0);
  void save_buffer(const std::string& outputpath, int channel_id, int gpu = 0);
  TracerSettings m_tis;
  Scene m_scene;
  std::vector<asuna_ctx*> m_ctx;
  std::vector<void*> m_comms;  // ncclComm_t per GPU when n_gpus > 1
  float m_build_ms = 0.f;
  std::vector<int> m_valid_pixel_index;  // --output_scanline state (tracer.cpp:344-345)
  std::deque<std::future<void>> m_writers;  // image files being encoded / written while the next shot renders
  void drain_writers(size_t keep);
};

}  // namespace asuna_host
