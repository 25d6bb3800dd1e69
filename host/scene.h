// Scene description host: Asuna's JSON scene format -> flat wire structs -> the C ABI of libasuna_b200.so.
//
// Mirrors the reference's Loader (src/loader/loader.cpp, src/loader/material.cpp) and Scene
// (src/scene/scene.cpp: name -> id tables with the dummy texture / material / light at index 0, rect and mesh lights
// becoming emitter instances, per-shot state) with the same JSON keys, defaults, id assignment order and error
// conditions -- except that errors are exceptions carrying the reference's message instead of exit(1).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../include/asuna_b200.h"
#include "image_io.h"
#include "json.h"
#include "linalg.h"

namespace asuna_host {

struct Shot {  // reference src/core/camera.h:8-15
  Vec3 eye, lookat, up;
  Mat4 env_transform = Mat4::identity();
  bool has_state = false;
  AsunaState state{};  // per-shot override (only six fields are honoured, scene.cpp:439-453)
};

struct CameraDesc {
  int type = ASUNA_CAMERA_PERSPECTIVE;
  int width = 0, height = 0;
  float fov = 45.0f, aperture = 0.0f, focal_distance = 0.1f;
  float fxfycxcy[4] = {0, 0, 0, 0};
};

struct MeshData {
  std::vector<AsunaVertex> vertices;
  std::vector<uint32_t> indices;
};

struct InstanceData {
  Mat4 xform = Mat4::identity();
  uint32_t mesh = 0, material = 0;
  int32_t light = -1;
};

struct OutputOptions {  // reference src/core/state.h:46-63 + loader.cpp:186-233
  bool hdr = false, render_result = true;
  std::vector<bool> channel_ldr;
  std::string tone_mapping = "filmic";
};

AsunaMaterial default_material();
AsunaState default_state();
AsunaSunSky default_sunsky();
AsunaLight dummy_light();
float compute_diffuse_fresnel(float ior, int n = 1000);
void envmap_tables(const ImageF& img, std::vector<float>& marginal, std::vector<float>& conditional);
MeshData load_obj(const std::string& path);

class Scene {
 public:
  Scene();
  // Loader::loadSceneFromJson (src/loader/loader.cpp:67-146)
  static Scene from_json_file(const std::string& path);

  int add_texture(const std::string& name, ImageF img);
  int add_material(const std::string& name, const AsunaMaterial& m);
  int add_mesh(const std::string& name, MeshData mesh);
  void add_instance(const std::string& mesh, const std::string& material, const Mat4& xform);
  int add_light(const AsunaLight& l);                                        // scene.cpp:220-246
  void add_mesh_light(const float radiance[3], const MeshData& mesh);        // scene.cpp:248-283
  void set_envmap(ImageF img);

  AsunaCamera gpu_camera(const Shot& shot) const;  // Camera::toGpuStruct, src/core/camera.cpp:77-99
  AsunaState shot_state(size_t shot_id) const;     // Scene::setShot, src/scene/scene.cpp:439-453
  void autofit_camera();                           // Scene::fitCamera when no shots are given (scene.cpp:520-553)

  // Scene::submit + PipelineRaytrace::init through the C ABI; returns the acceleration-structure build time in ms
  float upload(asuna_ctx* ctx) const;
  // Scene::setShot + setSpp(1) + resetFrame (src/tracer/tracer.cpp:206-212); returns the shot's total spp
  int begin_shot(asuna_ctx* ctx, size_t shot_id) const;
  // the flat arrays as a tagged binary blob (CPU tests compare it with the Python mirror; no GPU involved)
  void dump(const std::string& path) const;

  std::vector<ImageF> textures;
  std::map<std::string, int> texture_ids;
  std::vector<AsunaMaterial> materials;
  std::map<std::string, int> material_ids;
  std::vector<AsunaLight> lights;
  std::vector<MeshData> meshes;
  std::map<std::string, int> mesh_ids;
  std::vector<InstanceData> instances;
  bool has_envmap = false;
  ImageF envmap;
  std::vector<float> env_marginal, env_conditional;
  AsunaSunSky sunsky;
  AsunaState state;
  std::vector<Shot> shots;
  CameraDesc camera;
  OutputOptions output;
  std::string base_dir = ".";
};

}  // namespace asuna_host
