// asuna_b200 -- headless command line of the B200 path tracer.  Flags of the reference (src/main.cpp:18-24):
//   --offline  --gpu_id N  --output_scanline  --out PATH  --scene PATH  --spp N
// plus  --gpus N (1/2/4/8, frames of a shot split over the GPUs; --split shots gives every GPU whole shots instead), --output_f32, --dump-scene FILE, --report FILE.
// There is no window system on a B200 box: the online (GLFW) mode of the reference does not exist here and
// rendering is always the offline path.  `--spp` is accepted and ignored exactly like the reference does
// (Scene::setSpp drops its argument, src/scene/scene.cpp:373): spp comes from the scene file / the shot.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "tracer.h"

using namespace asuna_host;

static void usage() {
  fprintf(stderr,
          "usage: asuna_b200 --scene scene.json [--out PREFIX] [--offline] [--gpu_id N] [--gpus N] [--split frames|shots] [--output_scanline]\n"
          "                  [--output_f32] [--spp N (ignored, as in the reference)] [--dump-scene FILE] [--report FILE]\n");
}

int main(int argc, char** argv) {
  TracerSettings tis;
  tis.outputname = "asuna_out";  // src/main.cpp:22
  std::string dump;
  bool offline_given = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto value = [&]() -> std::string {
      if (i + 1 >= argc) {
        fprintf(stderr, "[x] missing value after %s\n", a.c_str());
        usage();
        exit(1);
      }
      return argv[++i];
    };
    if (a == "--offline") offline_given = true;
    else if (a == "--gpu_id") tis.gpu_id = atoi(value().c_str());
    else if (a == "--gpus") tis.n_gpus = atoi(value().c_str());
    else if (a == "--split") {  // frames (default): frame ranges of each shot + NCCL reduce; shots: whole shots per GPU
      std::string m = value();
      if (m != "frames" && m != "shots") {
        fprintf(stderr, "[x] --split takes frames or shots\n");
        exit(1);
      }
      tis.split_shots = m == "shots";
    }
    else if (a == "--output_scanline") tis.output_scanline = true;
    else if (a == "--output_f32") tis.output_f32 = true;
    else if (a == "--out") tis.outputname = value();
    else if (a == "--scene") tis.scenefile = value();
    else if (a == "--spp") (void)value();
    else if (a == "--dump-scene") dump = value();
    else if (a == "--convert") {  // image IO self-test hook: --convert IN OUT [GAMMA [TONEMAPPER]]
      std::string in = value(), out = value();
      float gamma = i + 1 < argc ? (float)atof(argv[++i]) : 1.0f;
      std::string tm = i + 1 < argc ? argv[++i] : "";
      try {
        ImageF img = read_image(in, gamma);
        if (!tm.empty()) {
          std::vector<float> o(img.px.size());
          tonemap(tm, img.w * img.h, img.px.data(), o.data());
          img.px.swap(o);
        }
        write_image(out, img.w, img.h, img.px.data());
      } catch (const std::exception& e) {
        fprintf(stderr, "[x] %s\n", e.what());
        return 1;
      }
      return 0;
    }
    else if (a == "--report") tis.report = value();
    else if (a == "--help" || a == "-h") {
      usage();
      return 0;
    } else {
      fprintf(stderr, "[x] unknown argument %s\n", a.c_str());
      usage();
      return 1;
    }
  }
  if (tis.scenefile.empty()) {
    usage();
    return 1;
  }
  try {
    if (!dump.empty()) {  // loader only: no GPU is touched
      Scene sc = Scene::from_json_file(tis.scenefile);
      sc.dump(dump);
      return 0;
    }
    if (!offline_given) fprintf(stderr, "[info] no window system on this target: running the offline path (pass --offline to silence this)\n");
    Tracer tracer(tis);
    tracer.init();
    auto reports = tracer.run();
    if (!tis.report.empty()) {
      std::ofstream f(tis.report);
      f << "{\"bvh_build_ms\": " << tracer.build_ms() << ", \"n_gpus\": " << tis.n_gpus << ", \"width\": " << tracer.scene().camera.width
        << ", \"height\": " << tracer.scene().camera.height << ", \"shots\": [";
      for (size_t i = 0; i < reports.size(); i++)
        f << (i ? ", " : "") << "{\"shot\": " << reports[i].shot << ", \"spp\": " << reports[i].spp << ", \"render_ms\": " << reports[i].render_ms
          << ", \"save_ms\": " << reports[i].save_ms << ", \"samples_per_s\": "
          << (double)reports[i].spp * tracer.scene().camera.width * tracer.scene().camera.height / (reports[i].render_ms * 1e-3) << "}";
      f << "]}\n";
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "[x] %s\n", e.what());
    return 1;
  }
  return 0;
}
