#include "scene.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace asuna_host {

namespace {

const float kNvToRad = (float)(3.14159265358979323846 / 180.0);

std::string dir_of(const std::string& path) {
  size_t s = path.find_last_of('/');
  return s == std::string::npos ? "." : path.substr(0, s);
}
std::string resolve(const std::string& base, const std::string& p) { return (!p.empty() && p[0] == '/') ? p : base + "/" + p; }

void set3(float dst[3], const Json& j) {
  std::vector<float> v = j.as_floats();
  if (v.size() < 3) throw std::runtime_error("expected a 3-vector");
  dst[0] = v[0], dst[1] = v[1], dst[2] = v[2];
}
Vec3 vec3_of(const Json& j) {
  float v[3];
  set3(v, j);
  return {v[0], v[1], v[2]};
}
Mat4 mat4_of(const Json& j) {  // row-major in the file (src/loader/utils.h:30-40)
  std::vector<float> v = j.as_floats();
  if (v.size() != 16) throw std::runtime_error("expected a 4x4 matrix (16 numbers)");
  Mat4 m;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m(r, c) = v[r * 4 + c];
  return m;
}

// reference src/loader/material.cpp:250-291 (eta, k per RGB)
struct ComplexIor {
  const char* name;
  float eta[3], k[3];
};
const ComplexIor kComplexIor[] = {
    {"a-C", {2.9440999183f, 2.2271502925f, 1.9681668794f}, {0.8874329109f, 0.7993216383f, 0.8152862927f}},
    {"Ag", {0.1552646489f, 0.1167232965f, 0.1383806959f}, {4.8283433224f, 3.1222459278f, 2.1469504455f}},
    {"Al", {1.6574599595f, 0.8803689579f, 0.5212287346f}, {9.2238691996f, 6.2695232477f, 4.8370012281f}},
    {"AlAs", {3.6051023902f, 3.2329365777f, 2.2175611545f}, {0.0006670247f, -0.0004999400f, 0.0074261204f}},
    {"AlSb", {-0.0485225705f, 4.1427547893f, 4.6697691348f}, {-0.0363741915f, 0.0937665154f, 1.3007390124f}},
    {"Au", {0.1431189557f, 0.3749570432f, 1.4424785571f}, {3.9831604247f, 2.3857207478f, 1.6032152899f}},
    {"Be", {4.1850592788f, 3.1850604423f, 2.7840913457f}, {3.8354398268f, 3.0101260162f, 2.8690088743f}},
    {"Cr", {4.3696828663f, 2.9167024892f, 1.6547005413f}, {5.2064337956f, 4.2313645277f, 3.7549467933f}},
    {"CsI", {2.1449030413f, 1.7023164587f, 1.6624194173f}, {0.0f, 0.0f, 0.0f}},
    {"Cu", {0.2004376970f, 0.9240334304f, 1.1022119527f}, {3.9129485033f, 2.4528477015f, 2.1421879552f}},
    {"Cu2O", {3.5492833755f, 2.9520622449f, 2.7369202137f}, {0.1132179294f, 0.1946659670f, 0.6001681264f}},
    {"CuO", {3.2453822204f, 2.4496293965f, 2.1974114493f}, {0.5202739621f, 0.5707372756f, 0.7172250613f}},
    {"d-C", {2.7112524747f, 2.3185812849f, 2.2288565009f}, {0.0f, 0.0f, 0.0f}},
    {"Hg", {2.3989314904f, 1.4400254917f, 0.9095512090f}, {6.3276269444f, 4.3719414152f, 3.4217899270f}},
    {"HgTe", {4.7795267752f, 3.2309984581f, 2.6600252401f}, {1.6319827058f, 1.5808189339f, 1.7295753852f}},
    {"Ir", {3.0864098394f, 2.0821938440f, 1.6178866805f}, {5.5921510077f, 4.0671757150f, 3.2672611269f}},
    {"K", {0.0640493070f, 0.0464100621f, 0.0381842017f}, {2.1042155920f, 1.3489364357f, 0.9132113889f}},
    {"Li", {0.2657871942f, 0.1956102432f, 0.2209198538f}, {3.5401743407f, 2.3111306542f, 1.6685930000f}},
    {"MgO", {2.0895885542f, 1.6507224525f, 1.5948759692f}, {0.0f, -0.0f, 0.0f}},
    {"Mo", {4.4837010280f, 3.5254578255f, 2.7760769438f}, {4.1111307988f, 3.4208716252f, 3.1506031404f}},
    {"Na", {0.0602665320f, 0.0561412435f, 0.0619909494f}, {3.1792906496f, 2.1124800781f, 1.5790940266f}},
    {"Nb", {3.4201353595f, 2.7901921379f, 2.3955856658f}, {3.4413817900f, 2.7376437930f, 2.5799132708f}},
    {"Ni", {2.3672753521f, 1.6633583302f, 1.4670554172f}, {4.4988329911f, 3.0501643957f, 2.3454274399f}},
    {"Rh", {2.5857954933f, 1.8601866068f, 1.5544279524f}, {6.7822927110f, 4.7029501026f, 3.9760892461f}},
    {"Se-e", {5.7242724833f, 4.1653992967f, 4.0816099264f}, {0.8713747439f, 1.1052845009f, 1.5647788766f}},
    {"Se", {4.0592611085f, 2.8426947380f, 2.8207582835f}, {0.7543791750f, 0.6385150558f, 0.5215872029f}},
    {"SiC", {3.1723450205f, 2.5259677964f, 2.4793623897f}, {0.0000007284f, -0.0000006859f, 0.0000100150f}},
    {"SnTe", {4.5251865890f, 1.9811525984f, 1.2816819226f}, {0.0f, 0.0f, 0.0f}},
    {"Ta", {2.0625846607f, 2.3930915569f, 2.6280684948f}, {2.4080467973f, 1.7413705864f, 1.9470377016f}},
    {"Te-e", {7.5090397678f, 4.2964603080f, 2.3698732430f}, {5.5842076830f, 4.9476231084f, 3.9975145063f}},
    {"Te", {7.3908396088f, 4.4821028985f, 2.6370708478f}, {3.2561412892f, 3.5273908133f, 3.2921683116f}},
    {"ThF4", {1.8307187117f, 1.4422274283f, 1.3876488528f}, {0.0f, 0.0f, 0.0f}},
    {"TiC", {3.7004673762f, 2.8374356509f, 2.5823030278f}, {3.2656905818f, 2.3515586388f, 2.1727857800f}},
    {"TiN", {1.6484691607f, 1.1504482522f, 1.3797795097f}, {3.3684596226f, 1.9434888540f, 1.1020123347f}},
    {"TiO2-e", {3.1065574823f, 2.5131551146f, 2.5823844157f}, {0.0000289537f, -0.0000251484f, 0.0001775555f}},
    {"TiO2", {3.4566203131f, 2.8017076558f, 2.9051485020f}, {0.0001026662f, -0.0000897534f, 0.0006356902f}},
    {"VC", {3.6575665991f, 2.7527298065f, 2.5326814570f}, {3.0683516659f, 2.1986687713f, 1.9631816252f}},
    {"VN", {2.8656011588f, 2.1191817791f, 1.9400767149f}, {3.0323264950f, 2.0561075580f, 1.6162930914f}},
    {"V", {4.2775126218f, 3.5131538236f, 2.7611257461f}, {3.4911844504f, 2.8893580874f, 3.1116965117f}},
    {"W", {4.3707029924f, 3.3002972445f, 2.9982666528f}, {3.5006778591f, 2.6048652781f, 2.2731930614f}},
};

float dielectric_reflectance(float eta, float cos_i) {  // src/loader/material.cpp:6-23
  if (cos_i < 0) {
    eta = 1.0f / eta;
    cos_i = -cos_i;
  }
  float sin2 = eta * eta * (1.0f - cos_i * cos_i);
  if (sin2 > 1) return 1.0f;
  float cos_t = std::sqrt(std::max(1.0f - sin2, 0.0f));
  float rs = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
  float rp = (eta * cos_t - cos_i) / (eta * cos_t + cos_i);
  return (rs * rs + rp * rp) * 0.5f;
}

int texture_ref(const Scene& sc, const Json& js, const char* key) {
  const std::string& name = js.at(key).as_string();
  auto it = sc.texture_ids.find(name);
  if (it == sc.texture_ids.end()) throw std::runtime_error("texture [" + name + "] does not exist");
  return it->second;
}

// src/loader/material.cpp:54-405
AsunaMaterial parse_material(const Scene& sc, const Json& js) {
  AsunaMaterial m = default_material();
  const std::string& t = js.at("type").as_string();
  auto opt3 = [&](const char* key, float* dst) {
    if (js.contains(key)) set3(dst, js.at(key));
  };
  auto opt1 = [&](const char* key, float& dst) {
    if (js.contains(key)) dst = js.at(key).as_float();
  };
  auto opt2 = [&](const char* key, float* dst) {
    if (js.contains(key)) {
      std::vector<float> v = js.at(key).as_floats();
      if (v.size() < 2) throw std::runtime_error(std::string("[") + key + "] must hold two numbers");
      dst[0] = v[0], dst[1] = v[1];
    }
  };
  auto opt_tex = [&](const char* key, int32_t& dst) {
    if (js.contains(key)) dst = texture_ref(sc, js, key);
  };
  if (t == "brdf_lambertian" || t == "brdf_mirror") {
    m.type = t == "brdf_mirror" ? ASUNA_MAT_MIRROR : ASUNA_MAT_LAMBERTIAN;
    opt3("diffuse_reflectance", m.diffuse), opt_tex("diffuse_texture", m.diffuseTextureId);
    opt_tex("normal_texture", m.normalTextureId);
  } else if (t == "brdf_pbr_metalness_roughness") {
    m.type = ASUNA_MAT_PBR_METALNESS_ROUGHNESS;
    opt_tex("normal_texture", m.normalTextureId), opt3("diffuse_reflectance", m.diffuse);
    opt_tex("diffuse_texture", m.diffuseTextureId), opt1("metalness", m.metalness);
    opt_tex("metalness_texture", m.metalnessTextureId), opt1("roughness", m.roughness);
    opt_tex("roughness_texture", m.roughnessTextureId);
    m.specular = 0.0f;  // used as opacity
    opt_tex("opacity_texture", m.opacityTextureId);
  } else if (t == "brdf_emissive") {
    m.type = ASUNA_MAT_EMISSIVE;
    opt3("radiance", m.radiance), opt3("radiance_factor", m.radianceFactor), opt_tex("radiance_texture", m.radianceTextureId);
  } else if (t == "brdf_kang18") {
    m.type = ASUNA_MAT_KANG18;
    opt_tex("normal_texture", m.normalTextureId), opt_tex("tangent_texture", m.tangentTextureId);
    if (js.contains("diffuse_texture")) m.diffuseTextureId = texture_ref(sc, js, "diffuse_texture");
    else set3(m.diffuse, js.at("diffuse_reflectance"));
    if (js.contains("specular_texture")) m.metalnessTextureId = texture_ref(sc, js, "specular_texture");
    else set3(m.rhoSpec, js.at("specular_reflectance"));
    if (js.contains("alpha_texture")) m.roughnessTextureId = texture_ref(sc, js, "alpha_texture");
    else {
      if (!js.contains("alpha")) throw std::runtime_error("missing key [\"alpha\"]");
      opt2("alpha", m.anisoAlpha);
    }
    m.metalness = 0.0f;  // used as opacity
    opt_tex("opacity_texture", m.opacityTextureId);
  } else if (t == "bsdf_dielectric") {
    m.type = ASUNA_MAT_DIELECTRIC;
    opt_tex("normal_texture", m.normalTextureId), opt1("ior", m.ior);
  } else if (t == "brdf_plastic" || t == "brdf_rough_plastic") {
    m.type = t == "brdf_plastic" ? ASUNA_MAT_PLASTIC : ASUNA_MAT_ROUGH_PLASTIC;
    opt1("ior", m.ior), opt3("diffuse_reflectance", m.diffuse), opt_tex("diffuse_texture", m.diffuseTextureId);
    opt_tex("normal_texture", m.normalTextureId);
    if (t == "brdf_rough_plastic") opt2("alpha", m.anisoAlpha), opt_tex("alpha_texture", m.roughnessTextureId);
    m.radiance[0] = compute_diffuse_fresnel(m.ior, 1000);
  } else if (t == "brdf_conductor" || t == "brdf_rough_conductor") {
    m.type = t == "brdf_conductor" ? ASUNA_MAT_CONDUCTOR : ASUNA_MAT_ROUGH_CONDUCTOR;
    std::string name = js.contains("material") ? js.at("material").as_string() : "Cu";
    const ComplexIor* found = nullptr;
    for (auto& c : kComplexIor)
      if (name == c.name) found = &c;
    if (!found) throw std::runtime_error("unrecognized material name in brdf_conductor [" + name + "]");
    std::memcpy(m.radiance, found->eta, sizeof m.radiance);
    std::memcpy(m.radianceFactor, found->k, sizeof m.radianceFactor);
    opt3("diffuse_reflectance", m.diffuse), opt_tex("diffuse_texture", m.diffuseTextureId);
    opt_tex("normal_texture", m.normalTextureId);
    if (t == "brdf_rough_conductor") opt2("alpha", m.anisoAlpha), opt_tex("alpha_texture", m.roughnessTextureId);
  } else if (t == "brdf_disney") {
    m.type = ASUNA_MAT_DISNEY;
    opt_tex("normal_texture", m.normalTextureId), opt3("diffuse_reflectance", m.diffuse);
    opt_tex("diffuse_texture", m.diffuseTextureId), opt1("metallic", m.metalness);
    opt_tex("metallic_texture", m.metalnessTextureId), opt1("roughness", m.roughness);
    opt_tex("roughness_texture", m.roughnessTextureId);
    m.rhoSpec[0] = (float)js.number_or("opacity", 0.0);
    opt_tex("opacity_texture", m.opacityTextureId);
  } else if (t == "brdf_phong") {
    m.type = ASUNA_MAT_PHONG;
    opt_tex("normal_texture", m.normalTextureId), opt3("diffuse_reflectance", m.diffuse);
    opt_tex("diffuse_texture", m.diffuseTextureId), opt3("specular_reflectance", m.rhoSpec), opt1("shininess", m.specular);
  } else
    throw std::runtime_error("unrecognized material type [" + t + "]");
  return m;
}

// src/loader/loader.cpp:148-234
void parse_state(const Json& js, AsunaState& st, OutputOptions* output) {
  static const char* names[7] = {"diffuse", "normal", "specular", "tangent", "roughness", "position", "uv"};
  const Json* pt = js.find("path_tracing");
  if (pt) {
    if (pt->contains("spp")) st.spp = pt->at("spp").as_int();
    if (pt->contains("max_path_depth")) st.maxPathDepth = pt->at("max_path_depth").as_int();
    if (pt->contains("use_face_normal")) st.useFaceNormal = pt->at("use_face_normal").as_bool() ? 1 : 0;
    if (pt->contains("ignore_emissive")) st.ignoreEmissive = pt->at("ignore_emissive").as_bool() ? 1 : 0;
    if (pt->contains("background_color")) set3(st.bgColor, pt->at("background_color"));
    if (pt->contains("envmap_intensity")) st.envMapIntensity = pt->at("envmap_intensity").as_float();
    if (pt->contains("multi_channel")) {
      const Json& mc = pt->at("multi_channel");
      if (mc.size() > ASUNA_NUM_OUTPUT_IMAGES - 1) throw std::runtime_error("channel numbers can not exceed 8");
      st.nMultiChannel = (uint32_t)mc.size();
      for (size_t cid = 0; cid < mc.size(); cid++) {
        const std::string& n = mc[cid].as_string();
        int32_t* dst[7] = {&st.diffuseOutChannel, &st.normalOutChannel,   &st.specularOutChannel, &st.tangentOutChannel,
                           &st.roughnessOutChannel, &st.positionOutChannel, &st.uvOutChannel};
        for (int k = 0; k < 7; k++)
          if (n == names[k]) *dst[k] = (int32_t)cid;
      }
    }
    if (output) {
      output->channel_ldr.assign(st.nMultiChannel, false);
      if (const Json* l = pt->find("multi_channel_ldr"))
        for (size_t cid = 0; cid < std::min<size_t>(st.nMultiChannel, l->size()); cid++) output->channel_ldr[cid] = (*l)[cid].as_bool();
    }
  }
  if (output) {
    if (const Json* pp = js.find("post_processing")) {
      std::string tm = pp->contains("tone_mapping") ? pp->at("tone_mapping").as_string() : "";
      static const char* known[] = {"none", "gamma", "reinhard", "Aces", "filmic", "pbrt", "custom"};
      bool ok = false;
      for (auto k : known) ok = ok || tm == k;
      if (!ok) {
        fprintf(stderr, "[warn] Loader: no matching tone mapper for [%s], use default\n", tm.c_str());
        tm = "none";  // loader.cpp:211: the fallback value is ToneMappingTypeNone
      }
      output->tone_mapping = tm;
    }
    if (js.contains("output_render_result")) output->render_result = js.at("output_render_result").as_bool();
    if (js.contains("output_hdr")) output->hdr = js.at("output_hdr").as_bool();
  }
}

// src/loader/loader.cpp:366-396
Mat4 parse_toworld(const Json& js, bool ban_translation) {
  Mat4 m = Mat4::identity();
  for (size_t i = 0; i < js.size(); i++) {
    const Json& s = js[i];
    const std::string& t = s.at("type").as_string();
    const Json& v = s.at("value");
    Mat4 x;
    if (t == "matrix") x = mat4_of(v);
    else if (t == "translate" && !ban_translation) x = translation(vec3_of(v));
    else if (t == "scale") x = scaling(vec3_of(v));
    else if (t == "rotx") x = rotation_x(kNvToRad * v.as_float());
    else if (t == "roty") x = rotation_y(kNvToRad * v.as_float());
    else if (t == "rotz") x = rotation_z(kNvToRad * v.as_float());
    else if (t == "rotate") {
      Vec3 a = vec3_of(v);
      x = rotation_z(kNvToRad * a.z) * rotation_y(kNvToRad * a.y) * rotation_x(kNvToRad * a.x);
    } else
      throw std::runtime_error("unrecognized toworld singleton type [" + t + "]");
    m = x * m;
  }
  return m;
}

AsunaVertex make_vertex(Vec3 p) {
  AsunaVertex v{};
  v.pos[0] = p.x, v.pos[1] = p.y, v.pos[2] = p.z;
  return v;
}

}  // namespace

AsunaMaterial default_material() {  // src/core/material.h:9-35
  AsunaMaterial m{};
  m.ior = 1.5f, m.roughness = 0.5f;
  m.radianceFactor[0] = m.radianceFactor[1] = m.radianceFactor[2] = 1.0f;
  m.diffuseTextureId = m.radianceTextureId = m.metalnessTextureId = m.normalTextureId = m.roughnessTextureId = m.tangentTextureId =
      m.opacityTextureId = -1;
  m.type = ASUNA_MAT_LAMBERTIAN;
  return m;
}
AsunaState default_state() {  // src/core/state.h:15-44
  AsunaState s{};
  s.curFrame = -1, s.spp = 1, s.maxPathDepth = 3, s.envMapIntensity = 1.0f;
  s.diffuseOutChannel = s.specularOutChannel = s.roughnessOutChannel = s.normalOutChannel = s.positionOutChannel = s.tangentOutChannel =
      s.uvOutChannel = -1;
  return s;
}
AsunaSunSky default_sunsky() {  // src/scene/scene.cpp:112-129
  AsunaSunSky s{};
  s.rgb_unit_conversion[0] = s.rgb_unit_conversion[1] = s.rgb_unit_conversion[2] = 1.0f;
  s.multiplier = 0.0000101320f, s.saturation = 1.0f;
  s.ground_color[0] = s.ground_color[1] = s.ground_color[2] = 0.4f;
  s.horizon_blur = 0.1f;
  s.night_color[2] = 0.01f;
  s.sun_disk_intensity = 0.8f;
  s.sun_direction[0] = 0.0f, s.sun_direction[1] = 0.78f, s.sun_direction[2] = 0.62f;
  s.sun_disk_scale = 5.0f, s.sun_glow_intensity = 1.0f;
  s.y_is_up = 1, s.physically_scaled_sun = 1, s.in_use = 0;
  return s;
}
AsunaLight dummy_light() {  // src/scene/scene.cpp:100-111
  AsunaLight l{};
  l.type = ASUNA_LIGHT_DIRECTIONAL;
  l.doubleSide = 1;
  return l;
}

float compute_diffuse_fresnel(float ior, int n) {  // src/loader/material.cpp:25-37: trapezoid rule, double accumulator
  double acc = 0.0;
  float fb = dielectric_reflectance(ior, 0.0f);
  for (int i = 1; i <= n; i++) {
    float cos2 = (float)i / (float)n;
    float fa = dielectric_reflectance(ior, std::min(std::sqrt(cos2), 1.0f));
    acc += (double)(fa + fb) * (0.5 / n);
    fb = fa;
  }
  return (float)acc;
}

// src/core/texture.cpp:144-226: inverse-CDF tables stored as RGBA32F images (.x = sampled coordinate, .y = pdf);
// only column 0 of the marginal table is populated
void envmap_tables(const ImageF& img, std::vector<float>& marginal, std::vector<float>& conditional) {
  const int w = img.w, h = img.h;
  std::vector<float> weight((size_t)w * h), cdf2d((size_t)w * h), pdf2d((size_t)w * h), row_sum(h), pdf1d(h), cdf1d(h);
  for (size_t i = 0; i < (size_t)w * h; i++)
    weight[i] = (float)(0.3 * (double)img.px[4 * i] + 0.6 * (double)img.px[4 * i + 1] + 0.1 * (double)img.px[4 * i + 2]);
  for (int j = 0; j < h; j++) {
    float run = 0.f;
    for (int i = 0; i < w; i++) run += weight[(size_t)j * w + i], cdf2d[(size_t)j * w + i] = run;
    row_sum[j] = run;
    double denom = (double)run + 1e-7;
    for (int i = 0; i < w; i++) {
      pdf2d[(size_t)j * w + i] = (float)((double)weight[(size_t)j * w + i] / denom);
      cdf2d[(size_t)j * w + i] = (float)((double)cdf2d[(size_t)j * w + i] / denom);
    }
  }
  float run = 0.f;
  for (int j = 0; j < h; j++) run += row_sum[j], cdf1d[j] = run;
  double total = (double)cdf1d[h - 1] + 1e-7;
  for (int j = 0; j < h; j++) pdf1d[j] = (float)((double)row_sum[j] / total), cdf1d[j] = (float)((double)cdf1d[j] / total);
  marginal.assign((size_t)w * h * 4, 0.f);
  conditional.assign((size_t)w * h * 4, 0.f);
  for (int j = 0; j < h; j++) {
    float inv = (float)(j + 1) / (float)h;
    int row = (int)(std::lower_bound(cdf1d.begin(), cdf1d.end(), inv) - cdf1d.begin());
    marginal[((size_t)j * w) * 4 + 0] = (float)row / (float)h;
    marginal[((size_t)j * w) * 4 + 1] = pdf1d[j];
    const float* c = &cdf2d[(size_t)j * w];
    for (int i = 0; i < w; i++) {
      float invw = (float)(i + 1) / (float)w;
      int col = (int)(std::lower_bound(c, c + w, invw) - c);
      conditional[((size_t)j * w + i) * 4 + 0] = (float)col / (float)w;
      conditional[((size_t)j * w + i) * 4 + 1] = pdf2d[(size_t)j * w + i];
    }
  }
}

// src/core/mesh.cpp:112-145 semantics (tinyobj, triangulate = true): fan triangulation, vertices unrolled per face
// corner (no index dedup), uv.y = 1 - v, indices = 0..n-1
MeshData load_obj(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("failed to load " + path);
  std::vector<Vec3> P, N;
  std::vector<std::pair<float, float>> T;
  struct Corner {
    int v, t, n;
  };
  std::vector<Corner> corners;
  std::string line;
  while (std::getline(f, line)) {
    std::istringstream is(line);
    std::string tag;
    if (!(is >> tag)) continue;
    if (tag == "v") {
      Vec3 p;
      is >> p.x >> p.y >> p.z;
      P.push_back(p);
    } else if (tag == "vt") {
      float u = 0, v = 0;
      is >> u >> v;
      T.emplace_back(u, v);
    } else if (tag == "vn") {
      Vec3 n;
      is >> n.x >> n.y >> n.z;
      N.push_back(n);
    } else if (tag == "f") {
      std::vector<Corner> face;
      std::string c;
      while (is >> c) {
        int vi = 0, ti = 0, ni = 0;
        size_t s1 = c.find('/');
        vi = std::atoi(c.substr(0, s1).c_str());
        if (s1 != std::string::npos) {
          size_t s2 = c.find('/', s1 + 1);
          std::string ts = c.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
          if (!ts.empty()) ti = std::atoi(ts.c_str());
          if (s2 != std::string::npos && s2 + 1 < c.size()) ni = std::atoi(c.substr(s2 + 1).c_str());
        }
        Corner k;
        k.v = vi > 0 ? vi - 1 : (int)P.size() + vi;
        k.t = ti > 0 ? ti - 1 : (ti < 0 ? (int)T.size() + ti : -1);
        k.n = ni > 0 ? ni - 1 : (ni < 0 ? (int)N.size() + ni : -1);
        face.push_back(k);
      }
      for (size_t k = 1; k + 1 < face.size(); k++) corners.push_back(face[0]), corners.push_back(face[k]), corners.push_back(face[k + 1]);
    }
  }
  MeshData m;
  m.vertices.resize(corners.size());
  m.indices.resize(corners.size());
  for (size_t i = 0; i < corners.size(); i++) {
    const Corner& c = corners[i];
    if (c.v < 0 || c.v >= (int)P.size()) throw std::runtime_error(path + ": vertex index out of range");
    AsunaVertex v = make_vertex(P[c.v]);
    if (c.t >= 0 && c.t < (int)T.size()) v.uv[0] = T[c.t].first, v.uv[1] = 1.0f - T[c.t].second;
    if (c.n >= 0 && c.n < (int)N.size()) v.normal[0] = N[c.n].x, v.normal[1] = N[c.n].y, v.normal[2] = N[c.n].z;
    m.vertices[i] = v;
    m.indices[i] = (uint32_t)i;
  }
  return m;
}

Scene::Scene() {  // src/scene/scene.cpp:87-130: dummies at index 0
  ImageF dummy;
  dummy.w = dummy.h = 1, dummy.px.assign(4, 0.f);
  textures.push_back(dummy);
  texture_ids["add_by_default_dummy_texture"] = 0;
  materials.push_back(default_material());
  material_ids["add_by_default_dummy_material"] = 0;
  lights.push_back(dummy_light());
  sunsky = default_sunsky();
  state = default_state();
}

int Scene::add_texture(const std::string& name, ImageF img) {
  int id = (int)textures.size();
  texture_ids[name] = id;
  textures.push_back(std::move(img));
  return id;
}
int Scene::add_material(const std::string& name, const AsunaMaterial& m) {
  int id = (int)materials.size();
  material_ids[name] = id;
  materials.push_back(m);
  return id;
}
int Scene::add_mesh(const std::string& name, MeshData mesh) {
  int id = (int)meshes.size();
  mesh_ids[name] = id;
  meshes.push_back(std::move(mesh));
  return id;
}
void Scene::add_instance(const std::string& mesh, const std::string& material, const Mat4& xform) {
  auto mi = mesh_ids.find(mesh);
  if (mi == mesh_ids.end()) throw std::runtime_error("mesh [" + mesh + "] does not exist");  // scene.cpp:324-327
  auto ma = material_ids.find(material);
  if (ma == material_ids.end()) throw std::runtime_error("material [" + material + "] does not exist");
  InstanceData in;
  in.xform = xform, in.mesh = (uint32_t)mi->second, in.material = (uint32_t)ma->second, in.light = -1;
  instances.push_back(in);
}
int Scene::add_light(const AsunaLight& l) {
  int light_id = (int)lights.size();
  if (l.type == ASUNA_LIGHT_RECT) {  // the light also becomes a 2-triangle emitter instance (src/core/mesh.cpp:19-32)
    Vec3 p{l.position[0], l.position[1], l.position[2]}, u{l.u[0], l.u[1], l.u[2]}, v{l.v[0], l.v[1], l.v[2]};
    MeshData m;
    m.vertices = {make_vertex(p), make_vertex(p + u), make_vertex(p + u + v), make_vertex(p + v)};
    m.indices = {0, 1, 2, 0, 2, 3};
    int mid = add_mesh("__rectLight:" + std::to_string(light_id), std::move(m));
    InstanceData in;
    in.mesh = (uint32_t)mid, in.material = 0, in.light = light_id;
    instances.push_back(in);
  }
  lights.push_back(l);
  return light_id;
}
void Scene::add_mesh_light(const float radiance[3], const MeshData& mesh) {
  for (size_t t = 0; t + 2 < mesh.indices.size(); t += 3) {
    const float* a = mesh.vertices[mesh.indices[t]].pos;
    const float* b = mesh.vertices[mesh.indices[t + 1]].pos;
    const float* c = mesh.vertices[mesh.indices[t + 2]].pos;
    AsunaLight l{};
    l.type = ASUNA_LIGHT_TRIANGLE;
    std::memcpy(l.radiance, radiance, sizeof l.radiance);
    Vec3 p{a[0], a[1], a[2]}, u = Vec3{b[0], b[1], b[2]} - p, v = Vec3{c[0], c[1], c[2]} - p;
    l.position[0] = p.x, l.position[1] = p.y, l.position[2] = p.z;
    l.u[0] = u.x, l.u[1] = u.y, l.u[2] = u.z;
    l.v[0] = v.x, l.v[1] = v.y, l.v[2] = v.z;
    l.area = length(cross(u, v)) * 0.5f;
    int light_id = (int)lights.size();
    MeshData m;
    m.vertices = {make_vertex(p), make_vertex(p + u), make_vertex(p + v)};
    m.indices = {0, 1, 2};
    int mid = add_mesh("__meshLight:" + std::to_string(light_id), std::move(m));
    InstanceData in;
    in.mesh = (uint32_t)mid, in.material = 0, in.light = light_id;
    instances.push_back(in);
    lights.push_back(l);
  }
}
void Scene::set_envmap(ImageF img) {
  envmap = std::move(img);
  envmap_tables(envmap, env_marginal, env_conditional);
  has_envmap = true;
  state.hasEnvMap = 1;
  state.envMapResolution[0] = (float)envmap.w, state.envMapResolution[1] = (float)envmap.h;
}

AsunaCamera Scene::gpu_camera(const Shot& shot) const {
  AsunaCamera c{};
  Mat4 view = scaling({1, -1, -1}) * look_at(shot.eye, shot.lookat, shot.up);  // src/core/camera.h:26-28
  invert_rot_trans(view).to_colmajor(c.cameraToWorld);
  shot.env_transform.to_colmajor(c.envTransform);
  if (camera.type == ASUNA_CAMERA_PERSPECTIVE) {
    c.type = ASUNA_CAMERA_PERSPECTIVE;
    float fov = std::min(std::max(camera.fov, 0.01f), 179.0f);  // cameramanipulator.cpp:321-324
    // src/core/camera.cpp:28-62,93-94 (fov is horizontal), assembled in double and inverted
    const double near_z = 0.1, far_z = 100.0, recip = 1.0 / (far_z - near_z);
    const double ctot = 1.0 / std::tan((double)fov * 3.14159265358979323846 / 180.0 * 0.5);
    const double aspect = camera.width / (double)camera.height, W = camera.width, H = camera.height;
    // raster = S(W,H,1) S(.5,.5,1) T(1,1,0) S(1,aspect,1) persp   applied to camera-space points
    double m[4][4] = {{0}};
    m[0][0] = ctot, m[1][1] = ctot, m[2][2] = far_z * recip, m[2][3] = -near_z * far_z * recip, m[3][2] = 1.0;
    for (int j = 0; j < 4; j++) m[1][j] *= aspect;                                 // S(1, aspect, 1)
    for (int j = 0; j < 4; j++) m[0][j] += m[3][j], m[1][j] += m[3][j];            // T(1, 1, 0)
    for (int j = 0; j < 4; j++) m[0][j] *= 0.5 * W, m[1][j] *= 0.5 * H;            // S(.5 W, .5 H, 1)
    Mat4 inv;  // invert in double, round once
    double w8[4][8];
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) w8[i][j] = m[i][j], w8[i][4 + j] = i == j ? 1.0 : 0.0;
    for (int col = 0; col < 4; col++) {
      int p = col;
      for (int r = col + 1; r < 4; r++)
        if (std::fabs(w8[r][col]) > std::fabs(w8[p][col])) p = r;
      if (p != col)
        for (int j = 0; j < 8; j++) std::swap(w8[p][j], w8[col][j]);
      double iv = 1.0 / w8[col][col];
      for (int j = 0; j < 8; j++) w8[col][j] *= iv;
      for (int r = 0; r < 4; r++)
        if (r != col) {
          double fct = w8[r][col];
          for (int j = 0; j < 8; j++) w8[r][j] -= fct * w8[col][j];
        }
    }
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) inv(i, j) = (float)w8[i][4 + j];
    inv.to_colmajor(c.rasterToCamera);
    c.focalDistance = camera.focal_distance;
    c.aperture = camera.aperture;
  } else {
    c.type = ASUNA_CAMERA_OPENCV;
    std::memcpy(c.fxfycxcy, camera.fxfycxcy, sizeof c.fxfycxcy);
  }
  return c;
}

AsunaState Scene::shot_state(size_t shot_id) const {
  AsunaState st = state;
  const Shot& sh = shots.at(shot_id);
  if (sh.has_state) {
    st.spp = sh.state.spp, st.maxPathDepth = sh.state.maxPathDepth, st.useFaceNormal = sh.state.useFaceNormal;
    st.ignoreEmissive = sh.state.ignoreEmissive, st.envMapIntensity = sh.state.envMapIntensity;
    std::memcpy(st.bgColor, sh.state.bgColor, sizeof st.bgColor);
  }
  st.numLights = (int32_t)lights.size() - 1;  // scene.cpp:33
  return st;
}

void Scene::autofit_camera() {
  // Scene::computeSceneDimensions + fitCamera (scene.cpp:520-553) with nvh::CameraManipulator::fit(min, max, true,
  // false, aspect): bounding sphere, default manipulator pose eye (10,10,10) -> centre (0,0,0), up (0,1,0)
  Vec3 lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
  for (auto& in : instances) {
    Vec3 mlo{1e30f, 1e30f, 1e30f}, mhi{-1e30f, -1e30f, -1e30f};
    for (auto& v : meshes[in.mesh].vertices)
      for (int k = 0; k < 3; k++) mlo[k] = std::min(mlo[k], v.pos[k]), mhi[k] = std::max(mhi[k], v.pos[k]);
    for (int c = 0; c < 8; c++) {
      Vec3 p = xf_point(in.xform, {c & 1 ? mhi.x : mlo.x, c & 2 ? mhi.y : mlo.y, c & 4 ? mhi.z : mlo.z}, 1.0f);
      for (int k = 0; k < 3; k++) lo[k] = std::min(lo[k], p[k]), hi[k] = std::max(hi[k], p[k]);
    }
  }
  bool volume = hi.x > lo.x && hi.y > lo.y && hi.z > lo.z;
  if (!volume) lo = {-1, -1, -1}, hi = {1, 1, 1};
  Vec3 half = (hi - lo) * 0.5f, centre = lo + half;
  float aspect = camera.width / (float)camera.height, fov = camera.fov;
  float radius = length(half);
  float offset = radius / std::sin(kNvToRad * (aspect > 1.f ? fov : fov * aspect) * 0.5f);
  Shot s;
  s.lookat = centre, s.eye = centre + normalize({10, 10, 10}) * offset, s.up = {0, 1, 0};
  shots.push_back(s);
}

Scene Scene::from_json_file(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("failed to load scene from file [" + path + "]");  // loader.cpp:73-77
  std::stringstream ss;
  ss << f.rdbuf();
  Json js = Json::parse(ss.str());
  for (const char* k : {"state", "camera", "meshes", "instances"})
    if (!js.contains(k)) throw std::runtime_error(std::string("missing key [\"") + k + "\"]");  // loader.cpp:92
  Scene sc;
  sc.base_dir = dir_of(path);
  // parse order is significant: ids are assigned by insertion (loader.cpp:91-144)
  parse_state(js.at("state"), sc.state, &sc.output);
  {
    const Json& cj = js.at("camera");
    const Json& res = cj.at("film").at("resolution");
    sc.camera.width = res[0].as_int(), sc.camera.height = res[1].as_int();
    const std::string& t = cj.at("type").as_string();
    if (t == "perspective") {
      sc.camera.type = ASUNA_CAMERA_PERSPECTIVE;
      sc.camera.fov = (float)cj.number_or("fov", 45.0);
      sc.camera.aperture = (float)cj.number_or("aperture", 0.0);
      sc.camera.focal_distance = (float)cj.number_or("focal_distance", 0.1);
    } else if (t == "opencv") {
      sc.camera.type = ASUNA_CAMERA_OPENCV;
      sc.camera.fxfycxcy[0] = cj.at("fx").as_float(), sc.camera.fxfycxcy[1] = cj.at("fy").as_float();
      sc.camera.fxfycxcy[2] = cj.at("cx").as_float(), sc.camera.fxfycxcy[3] = cj.at("cy").as_float();
    } else
      throw std::runtime_error("unrecognized camera type [" + t + "]");
  }
  if (const Json* tj = js.find("textures"))
    for (size_t i = 0; i < tj->size(); i++)
      sc.add_texture((*tj)[i].at("name").as_string(),
                     read_image(resolve(sc.base_dir, (*tj)[i].at("path").as_string()), (float)(*tj)[i].number_or("gamma", 1.0)));
  if (const Json* mj = js.find("materials"))
    for (size_t i = 0; i < mj->size(); i++) sc.add_material((*mj)[i].at("name").as_string(), parse_material(sc, (*mj)[i]));
  if (const Json* lj = js.find("lights"))
    for (size_t i = 0; i < lj->size(); i++) {
      const Json& j = (*lj)[i];
      AsunaLight l{};
      l.type = ASUNA_LIGHT_UNDEFINED;
      set3(l.radiance, j.at("radiance"));
      const std::string& t = j.at("type").as_string();
      if (t == "rect" || t == "triangle") {  // A.3-3: "triangle" is stored as a rect with halved area
        Vec3 p = vec3_of(j.at("position")), u = vec3_of(j.at("v1")) - p, v = vec3_of(j.at("v2")) - p;
        l.position[0] = p.x, l.position[1] = p.y, l.position[2] = p.z;
        l.u[0] = u.x, l.u[1] = u.y, l.u[2] = u.z;
        l.v[0] = v.x, l.v[1] = v.y, l.v[2] = v.z;
        l.area = length(cross(u, v));
        l.type = ASUNA_LIGHT_RECT;
        l.doubleSide = j.bool_or("double_side", false) ? 1 : 0;
        if (t == "triangle") l.area *= 0.5f;
      } else if (t == "point") {
        set3(l.position, j.at("position"));
        l.type = ASUNA_LIGHT_POINT;
      } else if (t == "distant") {
        set3(l.direction, j.at("direction"));
        l.type = ASUNA_LIGHT_DIRECTIONAL;
      } else if (t == "mesh") {
        sc.add_mesh_light(l.radiance, load_obj(resolve(sc.base_dir, j.at("path").as_string())));
        continue;
      } else
        throw std::runtime_error("unrecognized light type [" + t + "]");
      sc.add_light(l);
    }
  if (const Json* ej = js.find("envmap")) sc.set_envmap(read_image(resolve(sc.base_dir, ej->at("path").as_string()), 1.0f));
  if (const Json* sj = js.find("sunsky")) {  // documented extension: the reference can only enable it from the GUI
    struct F {
      const char* key;
      float* dst;
      int n;
    };
    AsunaSunSky& s = sc.sunsky;
    F fields[] = {{"rgb_unit_conversion", s.rgb_unit_conversion, 3}, {"multiplier", &s.multiplier, 1}, {"haze", &s.haze, 1},
                  {"redblueshift", &s.redblueshift, 1}, {"saturation", &s.saturation, 1}, {"horizon_height", &s.horizon_height, 1},
                  {"ground_color", s.ground_color, 3}, {"horizon_blur", &s.horizon_blur, 1}, {"night_color", s.night_color, 3},
                  {"sun_disk_intensity", &s.sun_disk_intensity, 1}, {"sun_direction", s.sun_direction, 3},
                  {"sun_disk_scale", &s.sun_disk_scale, 1}, {"sun_glow_intensity", &s.sun_glow_intensity, 1}};
    for (auto& fd : fields)
      if (const Json* v = sj->find(fd.key)) {
        if (fd.n == 3) set3(fd.dst, *v);
        else *fd.dst = v->as_float();
      }
    if (const Json* v = sj->find("y_is_up")) s.y_is_up = v->as_int();
    if (const Json* v = sj->find("physically_scaled_sun")) s.physically_scaled_sun = v->as_int();
    s.in_use = sj->contains("in_use") ? (sj->at("in_use").as_bool() ? 1 : 0) : 1;
  }
  {
    const Json& mj = js.at("meshes");
    for (size_t i = 0; i < mj.size(); i++) {
      MeshData m = load_obj(resolve(sc.base_dir, mj[i].at("path").as_string()));
      if (mj[i].bool_or("recompute_normal", false))  // src/core/mesh.cpp:61-67
        for (size_t t = 0; t + 2 < m.vertices.size(); t += 3) {
          Vec3 a{m.vertices[t].pos[0], m.vertices[t].pos[1], m.vertices[t].pos[2]};
          Vec3 b{m.vertices[t + 1].pos[0], m.vertices[t + 1].pos[1], m.vertices[t + 1].pos[2]};
          Vec3 c{m.vertices[t + 2].pos[0], m.vertices[t + 2].pos[1], m.vertices[t + 2].pos[2]};
          Vec3 n = cross(b - a, c - a);
          float l = length(n);
          n = n * (1.0f / l);
          for (int k = 0; k < 3; k++) m.vertices[t + k].normal[0] = n.x, m.vertices[t + k].normal[1] = n.y, m.vertices[t + k].normal[2] = n.z;
        }
      if (const Json* us = mj[i].find("uv_scale")) {
        std::vector<float> s = us->as_floats();
        for (auto& v : m.vertices) v.uv[0] *= s.at(0), v.uv[1] *= s.at(1);
      }
      sc.add_mesh(mj[i].at("name").as_string(), std::move(m));
    }
  }
  {
    const Json& ij = js.at("instances");
    for (size_t i = 0; i < ij.size(); i++) {
      if (!ij[i].contains("material")) throw std::runtime_error("instance without material (the reference dereferences a null default, A.3-13)");
      Mat4 x = ij[i].contains("toworld") ? parse_toworld(ij[i].at("toworld"), false) : Mat4::identity();
      sc.add_instance(ij[i].at("mesh").as_string(), ij[i].at("material").as_string(), x);
    }
  }
  if (const Json* sj = js.find("shots"))
    for (size_t i = 0; i < sj->size(); i++) {
      const Json& j = (*sj)[i];
      const std::string& t = j.at("type").as_string();
      Shot s;
      if (t == "lookat") {
        s.eye = vec3_of(j.at("eye")), s.lookat = vec3_of(j.at("lookat")), s.up = vec3_of(j.at("up"));
      } else if (t == "toworld") {  // loader.cpp:420-427
        Mat4 c2w = mat4_of(j.at("matrix"));
        s.eye = {c2w(0, 3), c2w(1, 3), c2w(2, 3)};
        s.up = xf_point(c2w, {0, 1, 0}, 0.f), s.lookat = xf_point(c2w, {0, 0, 1}, 1.f);
      } else if (t == "opencv") {  // loader.cpp:428-435
        Mat4 c2w = invert_rot_trans(mat4_of(j.at("matrix")));
        s.eye = {c2w(0, 3), c2w(1, 3), c2w(2, 3)};
        s.up = xf_point(c2w, {0, -1, 0}, 0.f), s.lookat = xf_point(c2w, {0, 0, 1}, 1.f);
      } else
        throw std::runtime_error("unrecognized shot type [" + t + "]");
      if (const Json* st = j.find("state")) {
        s.has_state = true;
        s.state = sc.state;
        parse_state(*st, s.state, nullptr);
      }
      if (const Json* e = j.find("env_toworld")) s.env_transform = parse_toworld(*e, true);
      sc.shots.push_back(s);
    }
  if (sc.shots.empty()) sc.autofit_camera();  // scene.cpp:78-81
  return sc;
}

static void check(asuna_ctx* ctx, int rc, const char* what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + " failed: " + asuna_last_error(ctx));
}

float Scene::upload(asuna_ctx* ctx) const {
  check(ctx, asuna_set_film(ctx, (uint32_t)camera.width, (uint32_t)camera.height), "asuna_set_film");
  for (auto& t : textures) check(ctx, asuna_add_texture(ctx, t.px.data(), (uint32_t)t.w, (uint32_t)t.h), "asuna_add_texture");
  for (auto& m : materials) check(ctx, asuna_add_material(ctx, &m), "asuna_add_material");
  check(ctx, asuna_set_lights(ctx, lights.data(), (uint32_t)lights.size()), "asuna_set_lights");
  if (has_envmap)
    check(ctx, asuna_set_envmap(ctx, envmap.px.data(), env_marginal.data(), env_conditional.data(), (uint32_t)envmap.w, (uint32_t)envmap.h),
          "asuna_set_envmap");
  for (auto& m : meshes)
    check(ctx, asuna_add_mesh(ctx, m.vertices.data(), (uint32_t)m.vertices.size(), m.indices.data(), (uint32_t)m.indices.size()),
          "asuna_add_mesh");
  for (auto& in : instances) {
    float x[16];
    in.xform.to_colmajor(x);
    check(ctx, asuna_add_instance(ctx, x, in.mesh, in.material, in.light), "asuna_add_instance");
  }
  check(ctx, asuna_set_sunsky(ctx, &sunsky), "asuna_set_sunsky");
  float ms = 0.f;
  check(ctx, asuna_build_accel(ctx, &ms), "asuna_build_accel");
  return ms;
}

int Scene::begin_shot(asuna_ctx* ctx, size_t shot_id) const {
  AsunaState st = shot_state(shot_id);
  int tot = st.spp;
  st.spp = 1, st.curFrame = -1;
  AsunaCamera cam = gpu_camera(shots.at(shot_id));
  check(ctx, asuna_set_camera(ctx, &cam), "asuna_set_camera");
  check(ctx, asuna_set_sunsky(ctx, &sunsky), "asuna_set_sunsky");
  check(ctx, asuna_set_state(ctx, &st), "asuna_set_state");
  check(ctx, asuna_reset_frame(ctx), "asuna_reset_frame");
  return tot;
}

void Scene::dump(const std::string& path) const {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot write " + path);
  auto blob = [&](const std::string& name, const void* data, size_t bytes) {
    uint32_t nl = (uint32_t)name.size();
    uint64_t nb = bytes;
    f.write((const char*)&nl, 4), f.write(name.data(), nl), f.write((const char*)&nb, 8);
    if (bytes) f.write((const char*)data, (std::streamsize)bytes);
  };
  int32_t film[2] = {camera.width, camera.height};
  blob("film", film, sizeof film);
  blob("materials", materials.data(), materials.size() * sizeof(AsunaMaterial));
  blob("lights", lights.data(), lights.size() * sizeof(AsunaLight));
  blob("sunsky", &sunsky, sizeof sunsky);
  for (size_t i = 0; i < textures.size(); i++) {
    int32_t wh[2] = {textures[i].w, textures[i].h};
    blob("texture_size:" + std::to_string(i), wh, sizeof wh);
    blob("texture:" + std::to_string(i), textures[i].px.data(), textures[i].px.size() * 4);
  }
  if (has_envmap) {
    int32_t wh[2] = {envmap.w, envmap.h};
    blob("envmap_size", wh, sizeof wh);
    blob("envmap", envmap.px.data(), envmap.px.size() * 4);
    blob("env_marginal", env_marginal.data(), env_marginal.size() * 4);
    blob("env_conditional", env_conditional.data(), env_conditional.size() * 4);
  }
  for (size_t i = 0; i < meshes.size(); i++) {
    blob("mesh_vertices:" + std::to_string(i), meshes[i].vertices.data(), meshes[i].vertices.size() * sizeof(AsunaVertex));
    blob("mesh_indices:" + std::to_string(i), meshes[i].indices.data(), meshes[i].indices.size() * 4);
  }
  for (size_t i = 0; i < instances.size(); i++) {
    float x[16];
    instances[i].xform.to_colmajor(x);
    blob("instance_xform:" + std::to_string(i), x, sizeof x);
    int32_t ids[3] = {(int32_t)instances[i].mesh, (int32_t)instances[i].material, instances[i].light};
    blob("instance_ids:" + std::to_string(i), ids, sizeof ids);
  }
  for (size_t i = 0; i < shots.size(); i++) {
    AsunaCamera c = gpu_camera(shots[i]);
    AsunaState st = shot_state(i);
    blob("shot_camera:" + std::to_string(i), &c, sizeof c);
    blob("shot_state:" + std::to_string(i), &st, sizeof st);
  }
  uint8_t out[2] = {(uint8_t)output.hdr, (uint8_t)output.render_result};
  blob("output", out, sizeof out);
  blob("tone_mapping", output.tone_mapping.data(), output.tone_mapping.size());
}

}  // namespace asuna_host
