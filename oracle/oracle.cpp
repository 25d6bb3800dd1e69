// TEST INFRASTRUCTURE ONLY.  CPU oracle for the asuna_b200 hot path.
//
// This file restates, in scalar fp32 C++, the per-pixel program the reference runs on the
// Vulkan ray-tracing pipeline:
//   reference src/shaders/raytrace.projective.rgen          -> Oracle::raygen / accumulate
//   reference src/shaders/raytrace.default.rmiss            -> Oracle::miss
//   reference src/shaders/raytrace.shadow.rmiss + rgen:117  -> Oracle::occluded
//   reference src/shaders/utils/rchit_layouts.glsl          -> getHitState / sampleLights
//   reference src/shaders/bxdf/raytrace.{brdf_lambertian,brdf_emissive,brdf_pbr_metalness_roughness,
//             brdf_plastic,brdf_rough_plastic,bsdf_dielectric,brdf_conductor,brdf_kang18}.rchit
//   reference src/pipeline/pipeline_raytrace.cpp:36-147     -> frame counter, instance table
// It exports a C ABI (oracle_*) that mirrors include/asuna_b200.h one to one so the parity
// tests can feed both sides the same flat scene.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product
// (asuna_b200/) never does.
//
// PARITY PINNED TO THE REFERENCE'S OWN SHADERS: the reference ships no golden vectors, tests or scenes for this path
// and cannot be built or run as a whole in this environment (no Vulkan loader/ICD, no glslang), but its
// GLSL can be compiled as C++ against the GLM it vendors (oracle/refbuild/build_ref.py -> oracle/_ref/libref.so = this
// file built with -DASUNA_REF_SHADERS + the generated shader unit).  tests/test_ref_pins.py holds every function and every
// shader main() below to that library (integer work bit-exact, fp32 within stated bounds), and tests/golden/ref_* holds
// vectors frozen FROM it.  Further anchors: published PCG / xxHash32 constants, closed forms of the rendering equation.
//
// Deliberate deviations from the shader text, all behaviour-neutral
// on the measured configs:
//   A.3-5  use_face_normal reads the geometric normal instead of an unassigned field;
//   A.3-7  with neither env nor lights (or a selector draw of exactly 1.0) lRec is zeroed,
//          so `visible` is false instead of undefined.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "bvh.h"
#include "shading.h"
// ref_bridge.h: plain-C structs shared with oracle/_ref/libref.so (oracle/refbuild/build_ref.py), which is this
// file compiled with -DASUNA_REF_SHADERS: same scene store, BVH and C ABI, but the per-pixel program is the
// reference's own GLSL compiled as C++ instead of the restatement below.
#include "refbuild/ref_bridge.h"

using namespace orc;

// Minimal dynamic-schedule parallel loop over [0,n) in chunks (std::thread; no OpenMP runtime needed).
template <class F>
static void parallel_for(int64_t n, int64_t chunk, int nthreads, F&& body) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads <= 1 || n <= chunk) {
    for (int64_t i = 0; i < n; i++) body(i);
    return;
  }
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    for (;;) {
      int64_t b = next.fetch_add(chunk);
      if (b >= n) break;
      int64_t e = std::min(n, b + chunk);
      for (int64_t i = b; i < e; i++) body(i);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthreads; t++) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
}

namespace {

struct Mesh {
  std::vector<AsunaVertex> v;
  std::vector<uint32_t> idx;
  Bvh bvh;
  Box box;
};

struct Instance {
  float o2w[12];  // 3 rows x 4 cols
  float w2o[12];
  uint32_t mesh, material;
  int32_t light;
  Box box;  // world space
};

struct Hit {
  float t = 0, b1 = 0, b2 = 0;
  uint32_t inst = 0xFFFFFFFFu, prim = 0xFFFFFFFFu;
};

inline vec3 xf_point(const float* m, vec3 p) {
  return {m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
          m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}
inline vec3 xf_vector(const float* m, vec3 v) {
  return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
          m[8] * v.x + m[9] * v.y + m[10] * v.z};
}
// v * M (row vector times 3x4): the GLSL `n * gl_WorldToObjectEXT` normal transform.
inline vec3 xf_normal(const float* m, vec3 n) {
  return {m[0] * n.x + m[4] * n.y + m[8] * n.z, m[1] * n.x + m[5] * n.y + m[9] * n.z,
          m[2] * n.x + m[6] * n.y + m[10] * n.z};
}

struct HitState {
  int lightId;
  vec2 uv;
  vec3 pos, V, N, geoN, ffN, X, Y;
  AsunaMaterial mat;
};

}  // namespace

struct oracle_ctx {
  uint32_t W = 0, H = 0;
  std::vector<Texture> textures;
  Texture env[3];  // 0 env, 1 marginal, 2 conditional (binding order, rchit_layouts.glsl:28)
  std::vector<Mesh> meshes;
  std::vector<AsunaMaterial> materials;
  std::vector<AsunaLight> lights;
  std::vector<Instance> instances;
  Bvh tlas;
  bool built = false;
  AsunaCamera cam{};
  AsunaSunSky sunsky{};
  AsunaState pc{};
  uint32_t rank = 0, world = 1;
  bool have_accum = false;
  std::vector<float> images[ASUNA_NUM_OUTPUT_IMAGES];
  std::vector<float> partial;
  AsunaStats stats{};
  std::atomic<uint64_t> n_closest{0}, n_shadow{0}, n_shadow_nz{0}, n_incoherent{0}, n_node_visits{0}, n_tri_tests{0};
  std::string err;
  int threads = 0;

  // ------------------------------------------------------------------ traversal
  void intersect_mesh(const Mesh& m, vec3 o, vec3 d, float tmin, uint32_t inst, Hit& best, float& tmax,
                      bool any, bool& found) const {
    if (m.bvh.nodes.empty() || m.idx.empty()) return;
    vec3 inv(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    RayShear rs(d);
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    uint64_t nv = 0, nt = 0;
    while (sp) {
      const BvhNode& n = m.bvh.nodes[stack[--sp]];
      float tn;
      nv++;
      if (!hit_box(n.box, o, inv, tmin, tmax, tn)) continue;
      if (n.count) {
        for (uint32_t i = 0; i < n.count; i++) {
          uint32_t prim = m.bvh.prims[n.left + i];
          const uint32_t* id = &m.idx[3 * (size_t)prim];
          float t, b1, b2;
          nt++;
          if (!hit_triangle(o, rs, vec3(m.v[id[0]].pos), vec3(m.v[id[1]].pos), vec3(m.v[id[2]].pos), t, b1, b2))
            continue;
          if (!(t > tmin)) continue;
          bool closer = t < tmax || (t == tmax && found && (inst < best.inst || (inst == best.inst && prim < best.prim)));
          if (!closer) continue;
          best = Hit{t, b1, b2, inst, prim};
          tmax = t;
          found = true;
          if (any) return;
        }
      } else {
        const BvhNode& a = m.bvh.nodes[n.left];
        const BvhNode& b = m.bvh.nodes[n.left + 1];
        float ta, tb;
        bool ha = hit_box(a.box, o, inv, tmin, tmax, ta), hb = hit_box(b.box, o, inv, tmin, tmax, tb);
        if (ha && hb) {
          if (ta < tb) {
            stack[sp++] = n.left + 1;
            stack[sp++] = n.left;
          } else {
            stack[sp++] = n.left;
            stack[sp++] = n.left + 1;
          }
        } else if (ha)
          stack[sp++] = n.left;
        else if (hb)
          stack[sp++] = n.left + 1;
      }
    }
    const_cast<oracle_ctx*>(this)->n_node_visits.fetch_add(nv, std::memory_order_relaxed);
    const_cast<oracle_ctx*>(this)->n_tri_tests.fetch_add(nt, std::memory_order_relaxed);
  }

  // traceRayEXT semantics: rays are taken into object space per instance, t is shared.
  bool trace(vec3 o, vec3 d, float tmin, float tmax, bool any, Hit& best) const {
    bool found = false;
    if (tlas.nodes.empty() || instances.empty()) return false;
    vec3 inv(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
      const BvhNode& n = tlas.nodes[stack[--sp]];
      float tn;
      if (!hit_box(n.box, o, inv, tmin, tmax, tn)) continue;
      if (n.count) {
        for (uint32_t i = 0; i < n.count; i++) {
          uint32_t ii = tlas.prims[n.left + i];
          const Instance& in = instances[ii];
          float ti;
          if (!hit_box(in.box, o, inv, tmin, tmax, ti)) continue;
          vec3 oo = xf_point(in.w2o, o), od = xf_vector(in.w2o, d);
          intersect_mesh(meshes[in.mesh], oo, od, tmin, ii, best, tmax, any, found);
          if (any && found) return true;
        }
      } else {
        stack[sp++] = n.left;
        stack[sp++] = n.left + 1;
      }
    }
    return found;
  }

  // ------------------------------------------------------------------ rchit_layouts.glsl
  vec4 textureEval(int texId, vec2 uv) const { return textureBilinear(textures[(size_t)texId], uv); }

  // rchit_layouts.glsl:61-65
  void configureShadingFrame(HitState& s) const {
    if (pc.useFaceNormal == 1) s.N = s.geoN;
    basis(s.N, s.X, s.Y);
    s.ffN = dot(s.N, s.V) > 0 ? s.N : -s.N;
  }
  // rchit_layouts.glsl:67-95
  HitState getHitState(const Hit& h, vec3 worldRayDir) const {
    HitState s;
    const Instance& in = instances[h.inst];
    const Mesh& m = meshes[in.mesh];
    const uint32_t* id = &m.idx[3 * (size_t)h.prim];
    const AsunaVertex &v0 = m.v[id[0]], &v1 = m.v[id[1]], &v2 = m.v[id[2]];
    vec3 ba(1.0f - h.b1 - h.b2, h.b1, h.b2);
    s.lightId = in.light;
    s.uv = {v0.uv[0] * ba.x + v1.uv[0] * ba.y + v2.uv[0] * ba.z, v0.uv[1] * ba.x + v1.uv[1] * ba.y + v2.uv[1] * ba.z};
    vec3 p0(v0.pos), p1(v1.pos), p2(v2.pos);
    s.pos = xf_point(in.o2w, p0 * ba.x + p1 * ba.y + p2 * ba.z);
    s.N = vec3(v0.normal) * ba.x + vec3(v1.normal) * ba.y + vec3(v2.normal) * ba.z;
    s.N = makeNormal(xf_normal(in.w2o, s.N));
    s.ffN = cross(p1 - p0, p2 - p0);
    s.ffN = makeNormal(xf_normal(in.w2o, s.ffN));
    s.geoN = s.ffN;  // A.3-5
    s.V = makeNormal(-worldRayDir);
    if (s.lightId < 0) s.mat = materials[in.material];
    configureShadingFrame(s);
    return s;
  }

  // ------------------------------------------------------------------ sample_light.glsl
  // :84-95
  float pdfEnvmap(vec3 L) const {
    const mat4& E = *reinterpret_cast<const mat4*>(cam.envTransform);
    L = transformDirection(transpose(E), L);
    float theta = std::acos(clampf(L.y, -1.0f, 1.0f));
    vec2 uv = {(PI + std::atan2(L.z, L.x)) * INV_2PI, theta * INV_PI};
    float pdf = textureBilinear(env[2], uv).y * textureBilinear(env[1], vec2{0.f, uv.y}).y;
    float sinTheta = std::sin(theta);
    if (sinTheta == 0) return 0;
    return (pdf * pc.envMapResolution[0] * pc.envMapResolution[1]) / (TWO_PI * PI * sinTheta);
  }
  // :97-104
  vec3 evalEnvmap(vec3 L) const {
    const mat4& E = *reinterpret_cast<const mat4*>(cam.envTransform);
    L = transformDirection(transpose(E), L);
    float theta = std::acos(clampf(L.y, -1.0f, 1.0f));
    vec2 uv = {(PI + std::atan2(L.z, L.x)) * INV_2PI, theta * INV_PI};
    vec4 c = textureBilinear(env[0], uv);
    return pc.envMapIntensity * vec3(c.x, c.y, c.z);
  }
  // :106-129
  vec3 sampleEnvmap(vec2 r, vec3& L, float& pdf) const {
    const mat4& E = *reinterpret_cast<const mat4*>(cam.envTransform);
    vec2 uv;
    uv.y = textureBilinear(env[1], vec2{0.f, r.x}).x;
    uv.x = textureBilinear(env[2], vec2{r.y, uv.y}).x;
    pdf = textureBilinear(env[2], uv).y * textureBilinear(env[1], vec2{0.f, uv.y}).y;
    float phi = uv.x * TWO_PI, theta = uv.y * PI;
    if (std::sin(theta) == 0.0f) pdf = 0.0f;
    pdf = (pdf * pc.envMapResolution[0] * pc.envMapResolution[1]) / (TWO_PI * PI * std::sin(theta));
    L = vec3(-std::sin(theta) * std::cos(phi), std::cos(theta), -std::sin(theta) * std::sin(phi));
    L = transformDirection(E, L);
    vec4 c = textureBilinear(env[0], uv);
    return pc.envMapIntensity * vec3(c.x, c.y, c.z);
  }
  // :10-82
  vec3 sampleOneLight(vec2 r, const AsunaLight& light, vec3 scatterPos, LightSamplingRecord& lRec) const {
    vec3 lu(light.u), lv(light.v), lp(light.position);
    switch (light.type) {
      case ASUNA_LIGHT_RECT:
      case ASUNA_LIGHT_TRIANGLE: {
        float r1 = r.x;
        float r2 = light.type == ASUNA_LIGHT_TRIANGLE ? (1 - r1) * r.y : r.y;  // :12-13 (A.3-2)
        vec3 lightSurfacePos = lp + lu * r1 + lv * r2;
        lRec.d = lightSurfacePos - scatterPos;
        lRec.dist = length(lRec.d);
        float distSq = lRec.dist * lRec.dist;
        lRec.d /= lRec.dist;
        lRec.n = makeNormal(cross(lu, lv));
        lRec.pdf = distSq / (light.area * std::fabs(dot(lRec.n, lRec.d)) + EPS);
        lRec.flags = EArea;
        return vec3(light.radiance);
      }
      case ASUNA_LIGHT_DIRECTIONAL: {
        lRec.d = makeNormal(vec3(light.direction));
        lRec.n = -lRec.d;
        lRec.dist = INFINITY_;
        lRec.pdf = 1.0f;
        lRec.flags = EDelta;
        return vec3(light.radiance);
      }
      case ASUNA_LIGHT_POINT: {
        lRec.d = lp - scatterPos;
        lRec.n = -lRec.d;
        lRec.dist = length(lRec.d);
        float distSq = lRec.dist * lRec.dist;
        lRec.d /= lRec.dist + EPS;  // A.3-12
        lRec.pdf = 1.0f;
        lRec.flags = EDelta;
        return vec3(light.radiance) / (distSq + EPS);
      }
    }
    return vec3(0.0f);
  }
  // rchit_layouts.glsl:98-116
  vec3 sampleEnvironmentLight(RayPayload& p, vec3 worldRayDir, LightSamplingRecord& lRec) const {
    lRec.flags = EArea;
    lRec.dist = INFINITY_;
    lRec.n = -makeNormal(worldRayDir);
    vec2 u = rand2(p.pRec.seed);
    if (sunsky.in_use == 1) {
      lRec.d = uniformSampleSphere(u);
      lRec.pdf = uniformSpherePdf();
      return sun_and_sky(sunsky, lRec.d);
    } else if (pc.hasEnvMap == 1) {
      return sampleEnvmap(u, lRec.d, lRec.pdf);
    }
    lRec.d = uniformSampleSphere(u);
    lRec.pdf = uniformSpherePdf();
    return vec3(pc.bgColor);
  }
  // rchit_layouts.glsl:118-167
  vec3 sampleLights(RayPayload& p, vec3 worldRayDir, vec3 scatterPos, vec3 scatterNormal, bool& visible,
                    LightSamplingRecord& lRec) const {
    bool allowDoubleSide = false;
    vec3 radiance(0.0f);
    lRec = LightSamplingRecord();  // A.3-7
    bool hasEnv = (pc.hasEnvMap == 1 || sunsky.in_use == 1);
    bool hasLight = (pc.numLights > 0);
    float envSelectPdf, analyticSelectPdf;
    if (hasEnv && hasLight)
      envSelectPdf = analyticSelectPdf = 0.5f;
    else if (hasEnv)
      envSelectPdf = 1.f, analyticSelectPdf = 0.f;
    else if (hasLight)
      envSelectPdf = 0.f, analyticSelectPdf = 1.f;
    else
      envSelectPdf = analyticSelectPdf = 0.f;

    float sel = rand1(p.pRec.seed);
    if (sel < envSelectPdf) {
      radiance = sampleEnvironmentLight(p, worldRayDir, lRec) / envSelectPdf;
      allowDoubleSide = true;
    } else if (sel < envSelectPdf + analyticSelectPdf) {
      int lightIndex = std::min(1 + (int)(rand1(p.pRec.seed) * pc.numLights), pc.numLights);
      const AsunaLight& light = lights[(size_t)lightIndex];
      vec2 r = rand2(p.pRec.seed);
      radiance = sampleOneLight(r, light, scatterPos, lRec) * (float)pc.numLights / analyticSelectPdf;
      allowDoubleSide = (light.doubleSide == 1);
    }
    p.dRec.ray.o = offsetPositionAlongNormal(scatterPos, scatterNormal);
    p.dRec.ray.d = lRec.d;
    p.dRec.dist = lRec.dist;
    visible = (dot(lRec.d, scatterNormal) > 0.0f && lRec.pdf > 0.0f);
    visible = visible && (dot(lRec.n, lRec.d) < 0 || allowDoubleSide);
    return radiance;
  }

  // Shared tails of every closest-hit main(): the NEE record and the next-ray update.
  void storeDirect(RayPayload& p, bool visible, vec3 bsdfWeight, float bsdfPdf, vec3 radiance,
                   const LightSamplingRecord& lRec) const {
    vec3 Ld(0.0f);
    if (visible) {
      float misWeight = powerHeuristic(lRec.pdf, bsdfPdf);
      Ld = misWeight * bsdfWeight * radiance * p.pRec.throughput / (lRec.pdf + EPS);
    }
    p.dRec.radiance = Ld;
    p.dRec.skip = !visible;
  }

  void applyNormalMap(HitState& s) const {  // e.g. brdf_lambertian.rchit:83-90
    if (s.mat.normalTextureId >= 0) {
      vec4 c = textureEval(s.mat.normalTextureId, s.uv);
      vec3 n = 2.0f * vec3(c.x, c.y, c.z) - 1.0f;
      s.N = toWorld(s.X, s.Y, s.N, n);
      s.N = makeNormal(s.N);
      configureShadingFrame(s);
    }
  }
  void fetchDiffuse(HitState& s) const {
    if (s.mat.diffuseTextureId >= 0) {
      vec4 c = textureEval(s.mat.diffuseTextureId, s.uv);
      for (int i = 0; i < 3; i++) s.mat.diffuse[i] = (&c.x)[i];
    }
  }
  void writeChannel(RayPayload& p, int ch, vec3 v) const {
    if (ch >= 0 && ch < ASUNA_NUM_OUTPUT_IMAGES - 1) p.channel[ch] = v;
  }

  // ------------------------------------------------------------------ shared microfacet pieces
  // identical text in brdf_pbr_metalness_roughness.rchit:18-53, brdf_rough_plastic.rchit:45-80,
  // brdf_kang18.rchit:18-53
  static float sqr(float x) { return x * x; }
  static float dAnisoGGX(float HdotN, float HdotX, float HdotY, float ax, float ay) {
    return 1 / (PI * ax * ay * sqr(sqr(HdotX / ax) + sqr(HdotY / ay) + sqr(HdotN)) + EPS);
  }
  static vec3 importanceSampleAnisoGGX(vec2 u, vec3 wo, float ax, float ay) {
    float factor = safeSqrt(u.x / std::fmax(1 - u.x, EPS));
    float phi = TWO_PI * u.y;
    vec3 wh(0, 0, 1);
    wh.x = -ax * factor * std::cos(phi);
    wh.y = -ay * factor * std::sin(phi);
    wh = makeNormal(wh);
    return reflect(-wo, wh);
  }
  static float importanceAnisoGGXPdf(vec3 wh, vec3 wo, float ax, float ay) {
    float pdf = 0.0f, HdotV = dot(wh, wo);
    vec3 wi = reflect(-wo, wh);
    if (wi.z > 0.0f && wo.z > 0.0f && wh.z > 0.0f)
      pdf = dAnisoGGX(wh.z, wh.x, wh.y, ax, ay) * std::fabs(wh.z) / (4 * HdotV + EPS);
    return pdf;
  }
  static float g1SmithAnisoGGX(float NdotV, float VdotX, float VdotY, float ax, float ay) {
    if (NdotV <= 0.0f) return 0.0f;
    vec3 factor(ax * VdotX, ay * VdotY, NdotV);
    return 1 / (NdotV + length(factor));
  }
  // brdf_plastic.rchit:12-41 == brdf_rough_plastic.rchit:12-41
  static float fresnelDielectricExt(float cosThetaI_, float eta) {
    if (eta == 1) return 0.0f;
    float scale = (cosThetaI_ > 0) ? 1 / eta : eta, cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) return 1.0f;
    float cosThetaI = std::fabs(cosThetaI_);
    float cosThetaT = std::sqrt(cosThetaTSqr);
    float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    return 0.5f * (Rs * Rs + Rp * Rp);
  }

  // ------------------------------------------------------------------ brdf_lambertian.rchit
  // :44-68
  void hitLight(RayPayload& p, int lightId, vec3 hitPos) const {
    const AsunaLight& light = lights[(size_t)lightId];
    vec3 lightDirection = makeNormal(p.pRec.ray.d);
    vec3 lightNormal = makeNormal(cross(vec3(light.u), vec3(light.v)));
    float lightSideProjection = dot(lightNormal, lightDirection);
    p.pRec.stop = true;
    if (lightSideProjection > 0 && light.doubleSide == 0) return;
    float misWeight = 1.0f;
    if (isNonSpecular(p.bRec.flags) && p.pRec.depth != 1) {
      float lightDist = length(hitPos - p.pRec.ray.o);
      float distSquare = lightDist * lightDist;
      float lightPdf = distSquare / (light.area * std::fabs(lightSideProjection) + EPS);
      misWeight = powerHeuristic(p.bRec.pdf, lightPdf);
    }
    p.pRec.radiance += p.pRec.throughput * vec3(light.radiance) * misWeight;
  }
  // :70-151
  void chitLambertian(RayPayload& p, HitState& s) const {
    fetchDiffuse(s);
    applyNormalMap(s);
    vec3 kd(s.mat.diffuse);
    if (p.pRec.depth == 1) {
      writeChannel(p, pc.diffuseOutChannel, kd);
      writeChannel(p, pc.normalOutChannel, s.N);
      writeChannel(p, pc.specularOutChannel, vec3(0.0f));
      writeChannel(p, pc.tangentOutChannel, s.X);
      writeChannel(p, pc.roughnessOutChannel, vec3(1, 1, 0));
      writeChannel(p, pc.positionOutChannel, s.pos);
      writeChannel(p, pc.uvOutChannel, vec3(s.uv.x, s.uv.y, 1));
    }
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);
      float bsdfPdf = 0;
      if (visible) {
        // eval(:12-21) with EArea; pdf(:23-29) with lRec.flags
        float NdotL = dot(s.ffN, lRec.d), NdotV = dot(s.ffN, s.V);
        if (!(NdotL < 0 || NdotV < 0)) w = kd * INV_PI * NdotL;
        if ((lRec.flags & EArea) != 0) bsdfPdf = cosineHemispherePdf(dot(s.ffN, lRec.d));
      }
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    // sampleBsdf(:31-42)
    BsdfSamplingRecord bRec;
    vec2 u = rand2(p.pRec.seed);
    vec3 wi = cosineSampleHemisphere(u);
    bRec.pdf = cosineHemispherePdf(wi.z);
    bRec.d = toWorld(s.X, s.Y, s.ffN, wi);
    bRec.flags = EDiffuseReflection;
    vec3 bsdfWeight = kd * INV_PI * std::fabs(wi.z);
    if (bRec.pdf <= 0.0f || isBlack(bsdfWeight)) {
      p.pRec.stop = true;
      return;
    }
    p.bRec = bRec;
    p.pRec.ray = Ray{offsetPositionAlongNormal(s.pos, s.ffN), bRec.d};
    p.pRec.throughput *= bsdfWeight / bRec.pdf;
  }

  // ------------------------------------------------------------------ brdf_emissive.rchit:12-26
  void chitEmissive(RayPayload& p, HitState& s) const {
    p.pRec.stop = true;
    vec3 rad(s.mat.radiance);
    if (s.mat.radianceTextureId >= 0) {
      vec4 c = textureEval(s.mat.radianceTextureId, s.uv);
      rad = vec3(s.mat.radianceFactor) * vec3(c.x, c.y, c.z);
    }
    if (pc.ignoreEmissive == 0) p.pRec.radiance += rad * p.pRec.throughput;
  }

  // common tail: reject / next ray with the `/(pdf+EPS)` convention
  void nextRay(RayPayload& p, const HitState& s, const BsdfSamplingRecord& bRec, vec3 bsdfWeight, vec3 offsetN) const {
    if (bRec.pdf <= 0.0f || length(bsdfWeight) == 0.0f) {
      p.pRec.stop = true;
      return;
    }
    p.bRec = bRec;
    p.pRec.ray = Ray{offsetPositionAlongNormal(s.pos, offsetN), bRec.d};
    p.pRec.throughput *= bsdfWeight / (bRec.pdf + EPS);
  }

  // ------------------------------------------------------------------ bsdf_dielectric.rchit
  static float dielectricFresnel(float cosThetaI, float eta) {  // :12-24
    float sinThetaTSq = eta * eta * (1.0f - cosThetaI * cosThetaI);
    if (sinThetaTSq > 1.0f) return 1.0f;
    float cosThetaT = std::sqrt(std::fmax(1.0f - sinThetaTSq, 0.0f));
    float rs = (eta * cosThetaT - cosThetaI) / (eta * cosThetaT + cosThetaI);
    float rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    return 0.5f * (rs * rs + rp * rp);
  }
  void chitDielectric(RayPayload& p, HitState& s) const {  // :74-138
    applyNormalMap(s);
    float eta = dot(s.V, s.N) > 0.0f ? (1.0f / s.mat.ior) : s.mat.ior;
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);
      float bsdfPdf = 0.0f;
      if (visible) {
        // eval(:26-38) is called with EArea -> always 0 (A.3-4); pdf(:41-52) with lRec.flags
        if ((lRec.flags & EDelta) != 0) {
          float F = dielectricFresnel(std::fabs(dot(s.V, s.ffN)), eta);
          if (dot(lRec.d, s.ffN) > 0) {
            if (std::fabs(dot(reflect(-s.V, s.ffN), lRec.d) - 1) < EPS) bsdfPdf = F;
          } else {
            if (std::fabs(dot(refract(-s.V, s.ffN, eta), lRec.d) - 1) < EPS) bsdfPdf = 1 - F;
          }
        }
      }
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    // sampleBsdf(:54-72)
    BsdfSamplingRecord bRec;
    float u = rand1(p.pRec.seed);
    float F = dielectricFresnel(std::fabs(dot(s.V, s.ffN)), eta);
    vec3 weight;
    if (u < F) {
      bRec.d = makeNormal(reflect(-s.V, s.ffN));
      bRec.pdf = F;
      bRec.flags = ESpecularReflection;
      weight = F * vec3(1.0f);
    } else {
      bRec.d = makeNormal(refract(-s.V, s.ffN, eta));
      bRec.pdf = 1 - F;
      bRec.flags = ESpecularTransmission;
      weight = (1 - F) * vec3(1.0f) * eta * eta;
    }
    nextRay(p, s, bRec, weight, signf(dot(bRec.d, s.N)) * s.N);
  }

  // ------------------------------------------------------------------ brdf_conductor.rchit
  static float conductorReflectance(float eta, float k, float cosThetaI) {  // :14-32 (A.3-10)
    float cosThetaISq = cosThetaI * cosThetaI;
    float sinThetaISq = std::fmax(1.0f - cosThetaISq, 0.0f);
    float sinThetaIQu = sinThetaISq * sinThetaISq;
    float innerTerm = eta * eta - k * k - sinThetaISq;
    float aSqPlusBSq = std::sqrt(std::fmax(innerTerm * innerTerm + 4.0f * eta * eta * k * k, 0.0f));
    float a = std::sqrt(std::fmax((aSqPlusBSq + innerTerm) * 0.5f, 0.0f));
    float Rs = ((aSqPlusBSq + cosThetaISq) - (2.0f * a * cosThetaI)) / ((aSqPlusBSq + cosThetaISq) + (2.0f * a * cosThetaI));
    float Rp = ((cosThetaISq * aSqPlusBSq + sinThetaIQu) - (2.0f * a * cosThetaI * sinThetaISq)) /
               ((cosThetaISq * aSqPlusBSq + sinThetaIQu) + (2.0f * a * cosThetaI * sinThetaISq));
    return 0.5f * (Rs + Rs * Rp);
  }
  static vec3 conductorReflectance3(vec3 eta, vec3 k, float c) {
    return {conductorReflectance(eta.x, k.x, c), conductorReflectance(eta.y, k.y, c), conductorReflectance(eta.z, k.z, c)};
  }
  void chitConductor(RayPayload& p, HitState& s) const {  // :89-155
    fetchDiffuse(s);
    applyNormalMap(s);
    vec3 kd(s.mat.diffuse), eta(s.mat.radiance), k(s.mat.radianceFactor);
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);  // eval(:40-52) with EArea: the EDelta test fails -> 0
      float bsdfPdf = 0.0f;
      if (visible) {  // pdf(:54-66)
        float NdotL = dot(lRec.d, s.ffN), NdotV = dot(s.V, s.ffN);
        if (!(NdotL < 0 || NdotV < 0 || ((lRec.flags & EDelta) == 0)))
          if (std::fabs(dot(reflect(-s.V, s.ffN), lRec.d) - 1) < EPS) bsdfPdf = 1.0f;
      }
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    // sampleBsdf(:68-87); consumes rand2 although unused
    (void)rand2(p.pRec.seed);
    BsdfSamplingRecord bRec;
    vec3 weight(0.0f);
    float NdotV = dot(s.V, s.ffN);
    if (NdotV <= 0) {
      bRec.flags = EBsdfNull;
      bRec.pdf = 0;
      bRec.d = vec3(0.0f);
    } else {
      bRec.d = reflect(-s.V, s.ffN);
      bRec.pdf = 1.0f;
      bRec.flags = ESpecularReflection;
      weight = kd * conductorReflectance3(eta, k, dot(s.ffN, bRec.d));
    }
    nextRay(p, s, bRec, weight, s.ffN);
  }

  // ------------------------------------------------------------------ brdf_mirror.rchit:39-101
  void chitMirror(RayPayload& p, HitState& s) const {
    fetchDiffuse(s);
    applyNormalMap(s);
    vec3 kd(s.mat.diffuse);
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);  // eval(:12-19) is called with EArea: the EDelta test fails -> 0
      float bsdfPdf = 0.0f;
      if (visible && (lRec.flags & EDelta) != 0)  // pdf(:21-26) with lRec.flags
        if (std::fabs(dot(reflect(-s.V, s.ffN), lRec.d) - 1) < EPS) bsdfPdf = 1.0f;
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    BsdfSamplingRecord bRec;  // sampleBsdf(:28-37): no random numbers
    bRec.pdf = 1.0f;
    bRec.d = reflect(-s.V, s.ffN);
    bRec.flags = ESpecularReflection;
    vec3 bsdfWeight = kd;
    if (bRec.pdf <= 0.0f || isBlack(bsdfWeight)) {
      p.pRec.stop = true;
      return;
    }
    p.bRec = bRec;
    p.pRec.ray = Ray{offsetPositionAlongNormal(s.pos, s.ffN), bRec.d};
    p.pRec.throughput *= bsdfWeight / bRec.pdf;
  }

  // ------------------------------------------------------------------ brdf_rough_conductor.rchit
  vec3 roughConductorEval(vec3 L, vec3 V, vec3 N, vec3 X, vec3 Y, vec3 kd, vec3 eta, vec3 k, float ax, float ay) const {  // :79-98, EArea
    float NdotL = dot(L, N), NdotV = dot(V, N);
    if (NdotL < 0 || NdotV < 0) return vec3(0.0f);
    vec3 H = makeNormal(L + V);
    vec3 Fs = conductorReflectance3(eta, k, NdotL);
    float Gs = g1SmithAnisoGGX(NdotV, dot(V, X), dot(V, Y), ax, ay);
    Gs *= g1SmithAnisoGGX(NdotL, dot(L, X), dot(L, Y), ax, ay);
    float Ds = dAnisoGGX(dot(H, N), dot(H, X), dot(H, Y), ax, ay);
    return kd * Fs * Gs * Ds * NdotL;
  }
  void chitRoughConductor(RayPayload& p, HitState& s) const {  // :151-220
    fetchDiffuse(s);
    if (s.mat.roughnessTextureId >= 0) {
      vec4 c = textureEval(s.mat.roughnessTextureId, s.uv);
      s.mat.anisoAlpha[0] = c.x, s.mat.anisoAlpha[1] = c.y;
    }
    applyNormalMap(s);
    vec3 kd(s.mat.diffuse), eta(s.mat.radiance), k(s.mat.radianceFactor);
    float ax = std::fmax(s.mat.anisoAlpha[0], EPS), ay = std::fmax(s.mat.anisoAlpha[1], EPS);
    const vec3 N = s.ffN, V = s.V, X = s.X, Y = s.Y;
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);
      float bsdfPdf = 0.0f;
      if (visible) {
        w = roughConductorEval(lRec.d, V, N, X, Y, kd, eta, k, ax, ay);
        float NdotL = dot(lRec.d, N), NdotV = dot(V, N);  // pdf(:100-115)
        if (!(NdotL < 0 || NdotV < 0 || ((lRec.flags & EArea) == 0))) {
          vec3 wi = toLocal(X, Y, N, lRec.d), wo = toLocal(X, Y, N, V);
          bsdfPdf = importanceAnisoGGXPdf(makeNormal(wi + wo), wo, ax, ay);
        }
      }
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    vec2 u = rand2(p.pRec.seed);  // sampleBsdf(:117-149)
    BsdfSamplingRecord bRec;
    vec3 weight(0.0f);
    float NdotV = dot(V, N);
    if (NdotV <= 0) {
      bRec.flags = EBsdfNull;
      bRec.pdf = 0;
      bRec.d = vec3(0.0f);
    } else {
      vec3 wo = toLocal(X, Y, N, V);
      vec3 wi = importanceSampleAnisoGGX(u, wo, ax, ay);
      vec3 L = bRec.d = toWorld(X, Y, N, wi);
      vec3 H = makeNormal(V + L);
      vec3 wh = makeNormal(wi + wo);
      float NdotL = dot(N, L);
      vec3 Fs = conductorReflectance3(eta, k, NdotL);
      float Gs = g1SmithAnisoGGX(NdotV, dot(V, X), dot(V, Y), ax, ay);
      Gs *= g1SmithAnisoGGX(NdotL, dot(L, X), dot(L, Y), ax, ay);
      float Ds = dAnisoGGX(dot(H, N), dot(H, X), dot(H, Y), ax, ay);
      bRec.flags = EGlossyReflection;
      bRec.pdf = importanceAnisoGGXPdf(wh, wo, ax, ay);
      weight = kd * Fs * Gs * Ds * NdotL;
    }
    nextRay(p, s, bRec, weight, s.ffN);
  }

  // ------------------------------------------------------------------ brdf_phong.rchit
  static vec3 phongEval(vec3 L, vec3 V, vec3 N, vec3 diffuse, vec3 specular, float shininess) {  // :12-31, EArea
    float NdotL = dot(N, L), NdotV = dot(N, V);
    if (NdotL < 0 || NdotV < 0) return vec3(0.0f);
    vec3 H = normalize(L + V);
    vec3 diffuseLobe = diffuse * INV_PI;
    vec3 specularLobe = specular * std::pow(std::fmax(dot(H, N), 0.0f), shininess) * (shininess + 2) * INV_2PI;
    float db = luminance(diffuse), sb = luminance(specular);
    float diffuseWeight = db / (db + sb), specularWeight = 1 - diffuseWeight;
    return (diffuseWeight * diffuseLobe + specularWeight * specularLobe) * NdotL;
  }
  static float pdfPhong(vec3 L, vec3 V, vec3 N, float shininess) {  // :33-38
    vec3 H = normalize(L + V);
    float NdotH = std::fmax(dot(H, N), 0.0f), VdotH = std::fmax(dot(H, V), 0.0f);
    return (shininess + 1) * INV_2PI * std::pow(NdotH, shininess) * 0.25f / (VdotH + EPS);
  }
  void chitPhong(RayPayload& p, HitState& s) const {  // :113-193
    fetchDiffuse(s);
    applyNormalMap(s);
    vec3 diffuse(s.mat.diffuse), specular(s.mat.rhoSpec);
    float shininess = s.mat.specular;
    if (p.pRec.depth == 1) {
      writeChannel(p, pc.diffuseOutChannel, diffuse);
      writeChannel(p, pc.normalOutChannel, s.ffN);
      writeChannel(p, pc.specularOutChannel, specular);
      writeChannel(p, pc.tangentOutChannel, s.X);
      writeChannel(p, pc.roughnessOutChannel, vec3(1, 1, 0));
      writeChannel(p, pc.positionOutChannel, s.pos);
      writeChannel(p, pc.uvOutChannel, vec3(s.uv.x, s.uv.y, 1));
    }
    const vec3 N = s.ffN, V = s.V;
    float db = luminance(diffuse), sb = luminance(specular);
    float diffuseWeight = db / (db + sb), specularWeight = 1 - diffuseWeight;
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);
      float bsdfPdf = 0.0f;
      if (visible) {
        w = phongEval(lRec.d, V, N, diffuse, specular, shininess);
        if ((lRec.flags & EArea) != 0)  // pdf(:40-53)
          bsdfPdf = diffuseWeight * cosineHemispherePdf(dot(N, lRec.d)) + specularWeight * pdfPhong(lRec.d, V, N, shininess);
      }
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    vec2 u = rand2(p.pRec.seed);  // sampleBsdf(:55-86)
    BsdfSamplingRecord bRec;
    if (u.x < diffuseWeight) {
      u.x /= diffuseWeight;
      vec3 wi = cosineSampleHemisphere(u);
      bRec.pdf = cosineHemispherePdf(wi.z);
      bRec.d = toWorld(s.X, s.Y, N, wi);
      bRec.flags = EDiffuseReflection;
    } else {
      u.x = (u.x - diffuseWeight) / specularWeight;
      float cosTheta = std::pow(u.x, 1 / (shininess + 1));
      float phi = TWO_PI * u.y;
      float sinTheta = safeSqrt(1 - cosTheta * cosTheta);
      vec3 wh(sinTheta * std::sin(phi), sinTheta * std::cos(phi), cosTheta);
      vec3 H = toWorld(s.X, s.Y, N, wh);
      vec3 L = reflect(-V, H);
      bRec.pdf = pdfPhong(L, V, N, shininess);
      bRec.d = L;
      bRec.flags = EGlossyReflection;
    }
    vec3 bsdfWeight = phongEval(bRec.d, V, N, diffuse, specular, shininess);
    if (bRec.pdf <= 0.0f || isBlack(bsdfWeight)) {
      p.pRec.stop = true;
      return;
    }
    p.bRec = bRec;
    p.pRec.ray = Ray{offsetPositionAlongNormal(s.pos, s.ffN), bRec.d};
    p.pRec.throughput *= bsdfWeight / bRec.pdf;
  }

  // ------------------------------------------------------------------ brdf_disney.rchit
  struct DisneyMaterial {  // :31-49
    vec3 baseColor;
    float anisotropic, metallic, roughness, subsurface, specularTint, sheen, sheenTint, clearcoat, clearcoatRoughness, ior, opacity, ax, ay;
  };
  static float GTR1(float NDotH, float a) {  // :55-60
    if (a >= 1.0f) return INV_PI;
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return (a2 - 1.0f) / (PI * std::log(a2) * t);
  }
  static vec3 SampleGTR1(float rgh, float r1, float r2) {  // :62-74 (r2 is unused there)
    float a = std::fmax(0.001f, rgh);
    float a2 = a * a;
    float phi = r1 * TWO_PI;
    float cosTheta = std::sqrt((1.0f - std::pow(a2, 1.0f - r1)) / (1.0f - a2));
    float sinTheta = clampf(std::sqrt(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
    return vec3(sinTheta * std::cos(phi), sinTheta * std::sin(phi), cosTheta);
  }
  static vec3 SampleGGXVNDF(vec3 V, float ax, float ay, float r1, float r2) {  // :96-114
    vec3 Vh = normalize(vec3(ax * V.x, ay * V.y, V.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    vec3 T1 = lensq > 0 ? vec3(-Vh.y, Vh.x, 0) * (1.0f / std::sqrt(lensq)) : vec3(1, 0, 0);
    vec3 T2 = cross(Vh, T1);
    float r = std::sqrt(r1);
    float phi = 2.0f * PI * r2;
    float t1 = r * std::cos(phi), t2 = r * std::sin(phi);
    float sv = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - sv) * std::sqrt(1.0f - t1 * t1) + sv * t2;
    vec3 Nh = t1 * T1 + t2 * T2 + std::sqrt(std::fmax(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(vec3(ax * Nh.x, ay * Nh.y, std::fmax(0.0f, Nh.z)));
  }
  static float GTR2Aniso(float NDotH, float HDotX, float HDotY, float ax, float ay) {  // :116-121
    float a = HDotX / ax, b = HDotY / ay;
    float c = a * a + b * b + NDotH * NDotH;
    return 1.0f / (PI * ax * ay * c * c);
  }
  static float SmithG(float NDotV, float alphaG) {  // :134-138
    float a = alphaG * alphaG, b = NDotV * NDotV;
    return (2.0f * NDotV) / (NDotV + std::sqrt(a + b - a * b));
  }
  static float SmithGAniso(float NDotV, float VDotX, float VDotY, float ax, float ay) {  // :140-145
    float a = VDotX * ax, b = VDotY * ay, c = NDotV;
    return (2.0f * NDotV) / (NDotV + std::sqrt(a * a + b * b + c * c));
  }
  static float SchlickFresnel(float u) {  // :147-151
    float m = clampf(1.0f - u, 0.0f, 1.0f);
    float m2 = m * m;
    return m2 * m2 * m;
  }
  static float DisneyFresnel(float metallic, float eta, float LDotH, float VDotH) {  // :167-171 (DielectricFresnel :153-165 == dielectric's)
    return mixf(dielectricFresnel(std::fabs(VDotH), eta), SchlickFresnel(LDotH), metallic);
  }
  static vec3 EvalDiffuse(const DisneyMaterial& mat, vec3 Csheen, vec3 V, vec3 L, vec3 H, float& pdf) {  // :173-197
    pdf = 0.0f;
    if (L.z <= 0.0f) return vec3(0.0f);
    float FL = SchlickFresnel(L.z), FV = SchlickFresnel(V.z), FH = SchlickFresnel(dot(L, H));
    float Fd90 = 0.5f + 2.0f * dot(L, H) * dot(L, H) * mat.roughness;
    float Fd = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
    float Fss90 = dot(L, H) * dot(L, H) * mat.roughness;
    float Fss = mixf(1.0f, Fss90, FL) * mixf(1.0f, Fss90, FV);
    float ss = 1.25f * (Fss * (1.0f / (L.z + V.z) - 0.5f) + 0.5f);
    vec3 Fsheen = FH * mat.sheen * Csheen;
    pdf = L.z * INV_PI;
    return (1.0f - mat.metallic) * (INV_PI * mixf(Fd, ss, mat.subsurface) * mat.baseColor + Fsheen);
  }
  static vec3 EvalSpecReflection(const DisneyMaterial& mat, float eta, vec3 specCol, vec3 V, vec3 L, vec3 H, float& pdf) {  // :199-212
    pdf = 0.0f;
    if (L.z <= 0.0f) return vec3(0.0f);
    float FM = DisneyFresnel(mat.metallic, eta, dot(L, H), dot(V, H));
    vec3 F = mix3(specCol, vec3(1.0f), FM);
    float D = GTR2Aniso(H.z, H.x, H.y, mat.ax, mat.ay);
    float G1 = SmithGAniso(std::fabs(V.z), V.x, V.y, mat.ax, mat.ay);
    float G2 = G1 * SmithGAniso(std::fabs(L.z), L.x, L.y, mat.ax, mat.ay);
    pdf = G1 * D / (4.0f * V.z);
    return F * D * G2 / (4.0f * L.z * V.z);
  }
  static vec3 EvalClearcoat(const DisneyMaterial& mat, vec3 V, vec3 L, vec3 H, float& pdf) {  // :214-226
    pdf = 0.0f;
    if (L.z <= 0.0f) return vec3(0.0f);
    float FH = dielectricFresnel(dot(V, H), 1.0f / 1.5f);
    float F = mixf(0.04f, 1.0f, FH);
    float D = GTR1(H.z, mat.clearcoatRoughness);
    float G = SmithG(L.z, 0.25f) * SmithG(V.z, 0.25f);
    float jacobian = 1.0f / (4.0f * dot(V, H));
    pdf = D * H.z * jacobian;
    return vec3(0.25f) * mat.clearcoat * F * D * G / (4.0f * L.z * V.z);
  }
  static void GetSpecColor(const DisneyMaterial& mat, float eta, vec3& specCol, vec3& sheenCol) {  // :228-236
    float lum = Luminance709(mat.baseColor);
    vec3 ctint = lum > 0.0f ? mat.baseColor / lum : vec3(1.0f);
    float F0 = (1.0f - eta) / (1.0f + eta);
    specCol = mix3(F0 * F0 * mix3(vec3(1.0f), ctint, mat.specularTint), mat.baseColor, mat.metallic);
    sheenCol = mix3(vec3(1.0f), ctint, mat.sheenTint);
  }
  static void GetLobeProbabilities(const DisneyMaterial& mat, vec3 specCol, float approxFresnel, float& diffuseWt, float& specReflectWt,
                                   float& clearcoatWt) {  // :238-250
    diffuseWt = Luminance709(mat.baseColor) * (1.0f - mat.metallic);
    specReflectWt = Luminance709(mix3(specCol, vec3(1.0f), approxFresnel));
    clearcoatWt = 0.25f * mat.clearcoat * (1.0f - mat.metallic);
    float totalWt = diffuseWt + specReflectWt + clearcoatWt;
    diffuseWt /= totalWt, specReflectWt /= totalWt, clearcoatWt /= totalWt;
  }
  static vec3 disneyEval(vec3 L, vec3 V, vec3 N, vec3 X, vec3 Y, const DisneyMaterial& mat, float eta, float& bsdfPdf) {  // :252-300, EArea
    vec3 weight(0.0f);
    bsdfPdf = 0.0f;
    V = toLocal(X, Y, N, V);
    L = toLocal(X, Y, N, L);
    if (L.z <= 0 || V.z <= 0) return weight;
    vec3 H = normalize(L + V);
    if (H.z < 0.0f) H = -H;
    vec3 specCol, sheenCol;
    GetSpecColor(mat, eta, specCol, sheenCol);
    float diffuseWt, specReflectWt, clearcoatWt;
    float fresnel = DisneyFresnel(mat.metallic, eta, dot(L, H), dot(V, H));
    GetLobeProbabilities(mat, specCol, fresnel, diffuseWt, specReflectWt, clearcoatWt);
    float pdf = 0.0f;
    if (diffuseWt > 0.0f && L.z > 0.0f) {
      weight += EvalDiffuse(mat, sheenCol, V, L, H, pdf);
      bsdfPdf += pdf * diffuseWt;
    }
    if (specReflectWt > 0.0f && L.z > 0.0f && V.z > 0.0f) {
      weight += EvalSpecReflection(mat, eta, specCol, V, L, H, pdf);
      bsdfPdf += pdf * specReflectWt;
    }
    if (clearcoatWt > 0.0f && L.z > 0.0f && V.z > 0.0f) {
      weight += EvalClearcoat(mat, V, L, H, pdf);
      bsdfPdf += pdf * clearcoatWt;
    }
    return weight * L.z;
  }
  void chitDisney(RayPayload& p, HitState& s) const {  // :373-487
    fetchDiffuse(s);
    if (s.mat.metalnessTextureId >= 0) s.mat.metalness = textureEval(s.mat.metalnessTextureId, s.uv).x;
    if (s.mat.roughnessTextureId >= 0) s.mat.roughness = textureEval(s.mat.roughnessTextureId, s.uv).x;
    applyNormalMap(s);
    float opacity = s.mat.rhoSpec[0];
    if (s.mat.opacityTextureId >= 0) opacity = textureEval(s.mat.opacityTextureId, s.uv).x;
    if (passThrough(p, s, opacity)) return;
    DisneyMaterial mat;
    {
      float aspect = std::sqrt(1.0f - s.mat.anisotropic * 0.9f);
      mat.ax = std::fmax(0.001f, s.mat.roughness * s.mat.roughness / aspect);
      mat.ay = std::fmax(0.001f, s.mat.roughness * s.mat.roughness * aspect);
      mat.baseColor = vec3(s.mat.diffuse);
      mat.anisotropic = s.mat.anisotropic;
      mat.metallic = s.mat.metalness;
      mat.roughness = std::fmax(s.mat.roughness * s.mat.roughness, 0.001f);
      mat.subsurface = s.mat.subsurface;
      mat.specularTint = s.mat.specularTint;
      mat.sheen = s.mat.sheen;
      mat.sheenTint = s.mat.sheenTint;
      mat.clearcoat = s.mat.clearcoat;
      mat.clearcoatRoughness = mixf(0.1f, 0.001f, s.mat.clearcoatGloss);
      mat.ior = s.mat.ior;
      mat.opacity = opacity;
    }
    float eta = dot(s.V, s.N) > 0.0f ? (1.0f / mat.ior) : mat.ior;
    const vec3 N = s.ffN, X = s.X, Y = s.Y;
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);
      float bsdfPdf = 0.0f;
      if (visible) w = disneyEval(lRec.d, s.V, N, X, Y, mat, eta, bsdfPdf);  // always called with EArea (:449-450)
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    vec2 u = rand2(p.pRec.seed);  // sampleBsdf(:302-371)
    BsdfSamplingRecord bRec;
    float pdf = 0.0f;
    vec3 f(0.0f);
    float r1 = u.x, r2 = u.y;
    vec3 V = toLocal(X, Y, N, s.V), L;
    vec3 specCol, sheenCol;
    GetSpecColor(mat, eta, specCol, sheenCol);
    float diffuseWt, specReflectWt, clearcoatWt;
    float approxFresnel = DisneyFresnel(mat.metallic, eta, V.z, V.z);
    GetLobeProbabilities(mat, specCol, approxFresnel, diffuseWt, specReflectWt, clearcoatWt);
    float cdf0 = diffuseWt, cdf1 = cdf0 + clearcoatWt;
    if (r1 < cdf0) {
      r1 /= cdf0;
      L = cosineSampleHemisphere(vec2{r1, r2});
      vec3 H = normalize(L + V);
      f = EvalDiffuse(mat, sheenCol, V, L, H, pdf);
      pdf *= diffuseWt;
      bRec.flags = EDiffuseReflection;
    } else if (r1 < cdf1) {
      r1 = (r1 - cdf0) / (cdf1 - cdf0);
      vec3 H = SampleGTR1(mat.clearcoatRoughness, r1, r2);
      if (H.z < 0.0f) H = -H;
      L = normalize(reflect(-V, H));
      f = EvalClearcoat(mat, V, L, H, pdf);
      pdf *= clearcoatWt;
      bRec.flags = EGlossyReflection;
    } else {
      r1 = (r1 - cdf1) / (1.0f - cdf1);
      vec3 H = SampleGGXVNDF(V, mat.ax, mat.ay, r1, r2);
      if (H.z < 0.0f) H = -H;
      L = normalize(reflect(-V, H));
      f = EvalSpecReflection(mat, eta, specCol, V, L, H, pdf);
      pdf *= specReflectWt;
      bRec.flags = EGlossyReflection;
    }
    bRec.d = toWorld(X, Y, N, L);
    bRec.pdf = pdf;
    // `abs(dot(N, L))` at :370 mixes the world-space normal with the local-space direction, as written
    vec3 bsdfWeight = f * std::fabs(dot(N, L));
    if (bRec.pdf <= 0.0f || isBlack(bsdfWeight)) {
      p.pRec.stop = true;
      return;
    }
    p.bRec = bRec;
    p.pRec.ray = Ray{offsetPositionAlongNormal(s.pos, s.ffN), bRec.d};
    p.pRec.throughput *= bsdfWeight / bRec.pdf;
  }

  // ------------------------------------------------------------------ brdf_plastic.rchit
  void chitPlastic(RayPayload& p, HitState& s) const {  // :133-204
    fetchDiffuse(s);
    applyNormalMap(s);
    vec3 kd(s.mat.diffuse);
    float eta = s.mat.ior;
    float fdrInt = s.mat.radiance[0];
    float dAvg = luminance(kd), sAvg = luminance(vec3(1.0f));
    float specularSamplingWeight = sAvg / (dAvg + sAvg);
    float invEta2 = 1 / (eta * eta);
    const vec3 N = s.ffN, V = s.V;
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);
      float bsdfPdf = 0.0f;
      if (visible) {
        vec3 L = lRec.d;
        float NdotL = dot(L, N), NdotV = dot(V, N);
        if (!(NdotL < 0 || NdotV < 0)) {
          float Fo = fresnelDielectricExt(NdotV, eta);
          {  // eval(:43-66) with flags = EArea -> diffuse branch
            float Fi = fresnelDielectricExt(NdotL, eta);
            vec3 diff = kd;
            diff /= (1.0f - diff * fdrInt);
            w = (1 - Fi) * (1 - Fo) * diff * invEta2 * INV_PI * NdotL;
          }
          {  // pdf(:68-92) with lRec.flags
            bool hasSpecular = ((lRec.flags & EDelta) != 0), hasDiffuse = ((lRec.flags & EArea) != 0);
            float probSpecular = (Fo * specularSamplingWeight) /
                                 (Fo * specularSamplingWeight + (1 - Fo) * (1 - specularSamplingWeight));
            if (hasSpecular) {
              if (std::fabs(dot(reflect(-V, N), L) - 1) < EPS) bsdfPdf = probSpecular;
            } else if (hasDiffuse) {
              bsdfPdf = NdotL * (1 - probSpecular);
            }
          }
        }
      }
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    // sampleBsdf(:94-131)
    vec2 u = rand2(p.pRec.seed);
    BsdfSamplingRecord bRec;
    vec3 weight(0.0f);
    float NdotV = dot(V, N);
    if (NdotV <= 0) {
      bRec.flags = EBsdfNull;
      bRec.pdf = 0;
      bRec.d = vec3(0.0f);
    } else {
      float Fo = fresnelDielectricExt(NdotV, eta);
      float probSpecular =
          (Fo * specularSamplingWeight) / (Fo * specularSamplingWeight + (1 - Fo) * (1 - specularSamplingWeight));
      if (u.x < probSpecular) {
        bRec.d = makeNormal(reflect(-V, N));
        bRec.flags = ESpecularReflection;
        bRec.pdf = probSpecular;
        weight = vec3(1.0f) * Fo;
      } else {
        u.x = (u.x - probSpecular) / (1 + EPS - probSpecular);
        vec3 wi = cosineSampleHemisphere(u);
        float Fi = fresnelDielectricExt(wi.z, eta);
        vec3 diff = kd;
        diff /= (1.0f - diff * fdrInt);
        bRec.pdf = (1 - probSpecular) * cosineHemispherePdf(wi.z);
        bRec.d = toWorld(s.X, s.Y, N, wi);
        bRec.flags = EDiffuseReflection;
        weight = (1 - Fi) * (1 - Fo) * invEta2 * diff * INV_PI * std::fabs(wi.z);
      }
    }
    nextRay(p, s, bRec, weight, s.ffN);
  }

  // ------------------------------------------------------------------ brdf_rough_plastic.rchit
  // eval(:82-111)
  vec3 roughPlasticEval(vec3 L, vec3 V, vec3 N, vec3 X, vec3 Y, vec3 kd, vec3 ks, float eta, float fdrInt,
                        float invEta2, float ax, float ay, uint32_t flags) const {
    vec3 weight(0.0f);
    float NdotL = dot(L, N), NdotV = dot(V, N);
    if (((flags & EArea) == 0) || NdotL < 0 || NdotV < 0) return weight;
    {
      vec3 H = makeNormal(L + V);
      float Fs = fresnelDielectricExt(dot(H, V), eta);
      float Gs = g1SmithAnisoGGX(NdotV, dot(V, X), dot(V, Y), ax, ay);
      Gs *= g1SmithAnisoGGX(NdotL, dot(L, X), dot(L, Y), ax, ay);
      float Ds = dAnisoGGX(dot(H, N), dot(H, X), dot(H, Y), ax, ay);
      weight += ks * Fs * Gs * Ds * NdotL;
    }
    {
      float Fo = fresnelDielectricExt(NdotV, eta);
      float Fi = fresnelDielectricExt(NdotL, eta);
      vec3 diff = kd;
      diff /= (1.0f - diff * fdrInt);
      weight += (1 - Fi) * (1 - Fo) * diff * invEta2 * INV_PI * NdotL;
    }
    return weight;
  }
  // pdf(:113-136)
  float roughPlasticPdf(vec3 L, vec3 V, vec3 N, vec3 X, vec3 Y, float ax, float ay, float eta,
                        float substrateSamplingWeight, uint32_t flags) const {
    float pdf = 0.0f;
    float NdotL = dot(L, N), NdotV = dot(V, N);
    if (NdotL < 0 || NdotV < 0 || ((flags & EArea) == 0)) return pdf;
    float Fo = fresnelDielectricExt(NdotV, eta);
    float substrateWeight = substrateSamplingWeight * (1.0f - Fo);
    float specularWeight = Fo;
    float probSpecular = specularWeight / (specularWeight + substrateWeight);
    vec3 wi = toLocal(X, Y, N, L), wo = toLocal(X, Y, N, V);
    vec3 wh = makeNormal(wi + wo);
    pdf += probSpecular * importanceAnisoGGXPdf(wh, wo, ax, ay);
    pdf += (1 - probSpecular) * cosineHemispherePdf(wi.z);
    return pdf;
  }
  void chitRoughPlastic(RayPayload& p, HitState& s) const {  // :191-269
    fetchDiffuse(s);
    if (s.mat.roughnessTextureId >= 0) {
      vec4 c = textureEval(s.mat.roughnessTextureId, s.uv);
      s.mat.anisoAlpha[0] = c.x;
      s.mat.anisoAlpha[1] = c.y;
    }
    applyNormalMap(s);
    vec3 kd(s.mat.diffuse);
    float eta = s.mat.ior, fdrInt = s.mat.radiance[0], invEta2 = 1 / (eta * eta);
    float ax = std::fmax(EPS, s.mat.anisoAlpha[0]), ay = std::fmax(EPS, s.mat.anisoAlpha[1]);
    float substrateSamplingWeight = luminance(kd);
    const vec3 ks(1.0f);
    const vec3 N = s.ffN, V = s.V, X = s.X, Y = s.Y;
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);
      float bsdfPdf = 0.0f;
      if (visible) {
        // A.3-1: the call sites (:234-240) pass (.., eta, fdrInt, ax, ay, invEta2, ..) into
        // eval(.., eta, fdrInt, invEta2, ax, ay, ..) and (.., eta, ax, ay, w, ..) into
        // pdf(.., ax, ay, eta, w, ..).  Reproduced positionally.
        w = roughPlasticEval(lRec.d, V, N, X, Y, kd, ks, eta, fdrInt, /*invEta2=*/ax, /*ax=*/ay, /*ay=*/invEta2, EArea);
        bsdfPdf = roughPlasticPdf(lRec.d, V, N, X, Y, /*ax=*/eta, /*ay=*/ax, /*eta=*/ay, substrateSamplingWeight, lRec.flags);
      }
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    // sampleBsdf(:138-189)
    vec2 u = rand2(p.pRec.seed);
    BsdfSamplingRecord bRec;
    vec3 weight(0.0f);
    float NdotV = dot(V, N);
    if (NdotV <= 0) {
      bRec.flags = EBsdfNull;
      bRec.pdf = 0;
      bRec.d = vec3(0.0f);
    } else {
      float Fo = fresnelDielectricExt(NdotV, eta);
      float substrateWeight = substrateSamplingWeight * (1.0f - Fo);
      float specularWeight = Fo;
      float probSpecular = specularWeight / (specularWeight + substrateWeight);
      if (u.x < probSpecular) {
        u.x = u.x / probSpecular;
        vec3 wo = toLocal(X, Y, N, V);
        vec3 wi = importanceSampleAnisoGGX(u, wo, ax, ay);
        vec3 L = bRec.d = toWorld(X, Y, N, wi);
        vec3 H = makeNormal(V + L);
        vec3 wh = makeNormal(wi + wo);
        float NdotL = dot(N, L);
        float Fs = fresnelDielectricExt(dot(H, V), eta);
        float Gs = g1SmithAnisoGGX(NdotV, dot(V, X), dot(V, Y), ax, ay);
        Gs *= g1SmithAnisoGGX(NdotL, dot(L, X), dot(L, Y), ax, ay);
        float Ds = dAnisoGGX(dot(H, N), dot(H, X), dot(H, Y), ax, ay);
        bRec.flags = EGlossyReflection;
        bRec.pdf = importanceAnisoGGXPdf(wh, wo, ax, ay) * probSpecular;
        weight = ks * Fs * Gs * Ds * NdotL;
      } else {
        u.x = (u.x - probSpecular) / (1 + EPS - probSpecular);
        vec3 wi = cosineSampleHemisphere(u);
        float Fi = fresnelDielectricExt(wi.z, eta);
        vec3 diff = kd;
        diff /= (1.0f - diff * fdrInt);
        bRec.pdf = (1 - probSpecular) * cosineHemispherePdf(wi.z);
        bRec.d = toWorld(X, Y, N, wi);
        bRec.flags = EDiffuseReflection;
        weight = (1 - Fi) * (1 - Fo) * invEta2 * diff * INV_PI * std::fabs(wi.z);
      }
    }
    nextRay(p, s, bRec, weight, s.ffN);
  }

  // ------------------------------------------------------------------ brdf_pbr_metalness_roughness.rchit
  static vec3 fresnelSchlick3(float cosThetaI, vec3 f0) {  // :14-16
    return f0 + (1.0f - f0) * std::pow(clampf(1 - cosThetaI, 0, 1), 5.0f);
  }
  vec3 pbrEval(vec3 L, vec3 V, vec3 N, vec3 X, vec3 Y, vec3 albedo, float ax, float ay, float metalness,
               float eta) const {  // :57-86
    vec3 weight(0.0f);
    float NdotL = dot(N, L), NdotV = dot(N, V);
    if (NdotL <= 0.0f || NdotV <= 0.0f) return weight;
    vec3 H = makeNormal(L + V);
    float HdotN = dot(H, N), HdotX = dot(H, X), HdotY = dot(H, Y), HdotV = dot(H, V);
    float F0 = sqr((eta - 1) / (eta + 1));
    vec3 Fs = fresnelSchlick3(HdotV, mix3(vec3(F0), albedo, metalness));
    float Ds = dAnisoGGX(HdotN, HdotX, HdotY, ax, ay);
    float Gs = g1SmithAnisoGGX(NdotV, dot(V, X), dot(V, Y), ax, ay);
    Gs *= g1SmithAnisoGGX(NdotL, dot(L, X), dot(L, Y), ax, ay);
    vec3 diffuse = albedo * (1 - metalness) * INV_PI;
    vec3 specular = Fs * Ds * Gs;
    weight += diffuse;
    weight += specular;
    weight *= NdotL;
    return weight;
  }
  float lobePdf(vec3 L, vec3 V, vec3 N, vec3 X, vec3 Y, float ax, float ay, float pDiffuse,
                uint32_t flags) const {  // pbr :88-105, kang18 :92-114
    float pdf = 0.0f;
    if ((flags & EArea) == 0) return pdf;
    float NdotL = dot(N, L), NdotV = dot(N, V);
    if (NdotL <= 0.0f || NdotV <= 0.0f) return pdf;
    vec3 H = makeNormal(L + V);
    vec3 wh = makeNormal(toLocal(X, Y, N, H));
    vec3 wo = makeNormal(toLocal(X, Y, N, V));
    vec3 wi = makeNormal(toLocal(X, Y, N, L));
    pdf += pDiffuse * cosineHemispherePdf(wi.z);
    pdf += (1 - pDiffuse) * importanceAnisoGGXPdf(wh, wo, ax, ay);
    return pdf;
  }
  // Returns true when the surface was passed through (opacity), :151-160 / kang18 :176-180.
  bool passThrough(RayPayload& p, const HitState& s, float opacity) const {
    if (rand1(p.pRec.seed) < opacity) {
      p.pRec.ray.o = offsetPositionAlongNormal(s.pos, -s.ffN);
      p.pRec.depth--;
      return true;
    }
    return false;
  }
  void chitPbr(RayPayload& p, HitState& s) const {  // :131-219
    fetchDiffuse(s);
    if (s.mat.metalnessTextureId >= 0) s.mat.metalness = textureEval(s.mat.metalnessTextureId, s.uv).x;
    if (s.mat.roughnessTextureId >= 0) s.mat.roughness = textureEval(s.mat.roughnessTextureId, s.uv).x;
    applyNormalMap(s);
    float opacity = s.mat.specular;
    if (s.mat.opacityTextureId >= 0) opacity = textureEval(s.mat.opacityTextureId, s.uv).x;
    if (passThrough(p, s, opacity)) return;
    float ax = std::fmax(sqr(s.mat.roughness), 0.001f), ay = ax;
    float eta = s.mat.ior, metalness = s.mat.metalness;
    vec3 albedo(s.mat.diffuse);
    if (p.pRec.depth == 1) {
      writeChannel(p, pc.diffuseOutChannel, albedo);
      writeChannel(p, pc.normalOutChannel, s.ffN);
    }
    const vec3 N = s.ffN, V = s.V, X = s.X, Y = s.Y;
    const float DIFFUSE_LOBE_PROBABILITY = 0.2f;
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);
      float bsdfPdf = 0.0f;
      if (visible) {
        w = pbrEval(lRec.d, V, N, X, Y, albedo, ax, ay, metalness, eta);
        bsdfPdf = lobePdf(lRec.d, V, N, X, Y, ax, ay, DIFFUSE_LOBE_PROBABILITY, lRec.flags);
      }
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    // sampleBsdf(:107-129): the rand2 argument is drawn before the lobe-select rand in the body
    vec2 u = rand2(p.pRec.seed);
    BsdfSamplingRecord bRec;
    vec3 wo = makeNormal(toLocal(X, Y, N, V));
    vec3 wi;
    if (rand1(p.pRec.seed) < DIFFUSE_LOBE_PROBABILITY) {
      wi = cosineSampleHemisphere(u);
      bRec.flags = EDiffuseReflection;
      bRec.pdf = cosineHemispherePdf(wi.z);
    } else {
      wi = importanceSampleAnisoGGX(u, wo, ax, ay);
      vec3 wh = makeNormal(wi + wo);
      bRec.pdf = importanceAnisoGGXPdf(wh, wo, ax, ay);
      bRec.flags = EGlossyReflection;
    }
    bRec.d = toWorld(X, Y, N, wi);
    vec3 weight = pbrEval(bRec.d, V, N, X, Y, albedo, ax, ay, metalness, eta) * std::fabs(wi.z);
    nextRay(p, s, bRec, weight, s.ffN);
  }

  // ------------------------------------------------------------------ brdf_kang18.rchit
  static float Luminance709(vec3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }  // :57-59
  vec3 kangEval(vec3 L, vec3 V, vec3 N, vec3 X, vec3 Y, vec3 kd, vec3 ks, float ax, float ay, float eta) const {  // :61-90
    vec3 weight(0.0f);
    float NdotL = dot(N, L), NdotV = dot(N, V);
    if (NdotL <= 0.0f || NdotV <= 0.0f) return weight;
    vec3 H = makeNormal(L + V);
    float HdotN = dot(H, N), HdotX = dot(H, X), HdotY = dot(H, Y), HdotV = dot(H, V);
    float F0 = sqr((eta - 1) / (eta + 1));
    float Fs = F0 + (1 - F0) * std::pow(clampf(1 - HdotV, 0, 1), 5.0f);
    float Ds = dAnisoGGX(HdotN, HdotX, HdotY, ax, ay);
    float Gs = g1SmithAnisoGGX(NdotV, dot(V, X), dot(V, Y), ax, ay);
    Gs *= g1SmithAnisoGGX(NdotL, dot(L, X), dot(L, Y), ax, ay);
    vec3 diffuse = kd * INV_PI;
    vec3 specular = ks * Fs * Ds * Gs;
    weight += diffuse;
    weight += specular;
    weight *= NdotL;
    return weight;
  }
  void chitKang18(RayPayload& p, HitState& s, const Instance& in) const {  // :141-249
    fetchDiffuse(s);
    vec3 rhoSpec(s.mat.rhoSpec);
    if (s.mat.metalnessTextureId >= 0) {
      vec4 c = textureEval(s.mat.metalnessTextureId, s.uv);
      rhoSpec = vec3(c.x, c.y, c.z);
    }
    if (s.mat.roughnessTextureId >= 0) {
      vec4 c = textureEval(s.mat.roughnessTextureId, s.uv);
      s.mat.anisoAlpha[0] = c.x;
      s.mat.anisoAlpha[1] = c.y;
    }
    float opacity = s.mat.metalness;
    if (s.mat.opacityTextureId >= 0) opacity = textureEval(s.mat.opacityTextureId, s.uv).x;
    if (s.mat.normalTextureId >= 0) {  // object-space normal map, :157-163
      vec4 c = textureEval(s.mat.normalTextureId, s.uv);
      vec3 n = 2.0f * vec3(c.x, c.y, c.z) - 1.0f;
      s.N = makeNormal(xf_normal(in.w2o, n));
      configureShadingFrame(s);
    }
    if (s.mat.tangentTextureId >= 0) {  // :165-169
      vec4 c = textureEval(s.mat.tangentTextureId, s.uv);
      vec3 t = 2.0f * vec3(c.x, c.y, c.z) - 1.0f;
      s.X = makeNormal(xf_normal(in.w2o, t));
    }
    s.Y = makeNormal(cross(s.N, s.X));
    s.X = makeNormal(cross(s.Y, s.N));
    s.ffN = dot(s.N, s.V) > 0 ? s.N : -s.N;
    if (passThrough(p, s, opacity)) return;
    float ax = std::fmax(s.mat.anisoAlpha[0], EPS), ay = std::fmax(s.mat.anisoAlpha[1], EPS);
    float eta = s.mat.ior;
    vec3 kd(s.mat.diffuse);
    if (p.pRec.depth == 1) {
      writeChannel(p, pc.diffuseOutChannel, kd);
      writeChannel(p, pc.normalOutChannel, s.ffN);
      writeChannel(p, pc.specularOutChannel, rhoSpec);
      writeChannel(p, pc.tangentOutChannel, s.X);
      writeChannel(p, pc.roughnessOutChannel, vec3(ax, ay, 0));
      writeChannel(p, pc.positionOutChannel, s.pos);
      writeChannel(p, pc.uvOutChannel, vec3(s.uv.x, s.uv.y, 1));
    }
    const vec3 N = s.ffN, V = s.V, X = s.X, Y = s.Y;
    float pd_ = Luminance709(kd), ps_ = Luminance709(rhoSpec);
    float pDiffuse = pd_ / (pd_ + ps_ + EPS);
    {
      bool visible;
      LightSamplingRecord lRec;
      vec3 radiance = sampleLights(p, p.pRec.ray.d, s.pos, s.ffN, visible, lRec);
      vec3 w(0.0f);
      float bsdfPdf = 0.0f;
      if (visible) {
        w = kangEval(lRec.d, V, N, X, Y, kd, rhoSpec, ax, ay, eta);
        bsdfPdf = lobePdf(lRec.d, V, N, X, Y, ax, ay, pDiffuse, lRec.flags);
      }
      storeDirect(p, visible, w, bsdfPdf, radiance, lRec);
    }
    // sampleBsdf(:116-139): rand2 argument first, then the lobe-select rand; no |wi.z| factor
    vec2 u = rand2(p.pRec.seed);
    BsdfSamplingRecord bRec;
    vec3 wo = makeNormal(toLocal(X, Y, N, V));
    vec3 wi;
    if (rand1(p.pRec.seed) < pDiffuse) {
      wi = cosineSampleHemisphere(u);
      bRec.flags = EDiffuseReflection;
      bRec.pdf = cosineHemispherePdf(wi.z);
    } else {
      wi = importanceSampleAnisoGGX(u, wo, ax, ay);
      vec3 wh = makeNormal(wi + wo);
      bRec.pdf = importanceAnisoGGXPdf(wh, wo, ax, ay);
      bRec.flags = EGlossyReflection;
    }
    bRec.d = toWorld(X, Y, N, wi);
    vec3 weight = kangEval(bRec.d, V, N, X, Y, kd, rhoSpec, ax, ay, eta);
    nextRay(p, s, bRec, weight, s.ffN);
  }

  // ------------------------------------------------------------------ raytrace.default.rmiss:24-55
  void miss(RayPayload& p) const {
    p.pRec.stop = true;
    vec3 env(0.0f), d = p.pRec.ray.d;
    if (sunsky.in_use == 1)
      env = sun_and_sky(sunsky, d);
    else if (pc.hasEnvMap == 1)
      env = evalEnvmap(d);
    else
      env = vec3(pc.bgColor);
    float misWeight = 1.0f;
    float envPdf = 0.0f;
    if (p.pRec.depth != 1 && isNonSpecular(p.bRec.flags)) {
      if (sunsky.in_use == 1)
        envPdf = uniformSpherePdf();
      else if (pc.hasEnvMap == 1)
        envPdf = pdfEnvmap(d);
      else
        envPdf = uniformSpherePdf();
      misWeight = powerHeuristic(p.bRec.pdf, envPdf);
    }
    p.pRec.radiance += p.pRec.throughput * env * misWeight;
  }

  // SBT dispatch: instanceShaderBindingTableRecordOffset = material type, emitters ->
  // lambertian hit group (pipeline_raytrace.cpp:134-140).
  int closestHit(RayPayload& p, const Hit& h) const {
    const Instance& in = instances[h.inst];
    HitState s = getHitState(h, p.pRec.ray.d);
    if (s.lightId >= 0) {  // brdf_lambertian.rchit:74-78
      hitLight(p, s.lightId, s.pos);
      return 0;
    }
    switch (s.mat.type) {
      case ASUNA_MAT_LAMBERTIAN: chitLambertian(p, s); break;
      case ASUNA_MAT_EMISSIVE: chitEmissive(p, s); break;
      case ASUNA_MAT_DIELECTRIC: chitDielectric(p, s); break;
      case ASUNA_MAT_CONDUCTOR: chitConductor(p, s); break;
      case ASUNA_MAT_PLASTIC: chitPlastic(p, s); break;
      case ASUNA_MAT_ROUGH_PLASTIC: chitRoughPlastic(p, s); break;
      case ASUNA_MAT_PBR_METALNESS_ROUGHNESS: chitPbr(p, s); break;
      case ASUNA_MAT_KANG18: chitKang18(p, s, in); break;
      case ASUNA_MAT_MIRROR: chitMirror(p, s); break;
      case ASUNA_MAT_ROUGH_CONDUCTOR: chitRoughConductor(p, s); break;
      case ASUNA_MAT_PHONG: chitPhong(p, s); break;
      case ASUNA_MAT_DISNEY: chitDisney(p, s); break;
      default: p.pRec.stop = true; return -1;
    }
    return 0;
  }

  // ------------------------------------------------------------------ raytrace.projective.rgen:26-179
  void cameraRay(uint32_t x, uint32_t y, vec2 jitter, uint32_t& seed, vec3& rayOrigin, vec3& rayDir) const {
    const mat4& c2w = *reinterpret_cast<const mat4*>(cam.cameraToWorld);
    const mat4& r2c = *reinterpret_cast<const mat4*>(cam.rasterToCamera);
    vec3 origin = transformPoint(c2w, vec3(0.0f));
    vec2 pixel = {(float)x + jitter.x, (float)y + jitter.y};
    rayOrigin = origin;
    rayDir = vec3(0.0f);
    if (cam.type == ASUNA_CAMERA_PERSPECTIVE) {
      vec3 pCamera = transformPoint(r2c, vec3(pixel.x, pixel.y, 0.f));
      vec3 r = makeNormal(pCamera);
      if (cam.aperture > 0.f) {
        vec2 uLens = rand2(seed);
        vec2 pLens = cam.aperture * concentricSampleDisk(uLens);
        float ft = cam.focalDistance / r.z;
        vec3 pFocus = ft * r;
        vec3 o(pLens.x, pLens.y, 0.f);
        rayOrigin = transformPoint(c2w, o);
        r = pFocus - o;
      }
      rayDir = transformDirection(c2w, r);
    } else if (cam.type == ASUNA_CAMERA_OPENCV) {
      vec3 r((pixel.x - cam.fxfycxcy[2]) / cam.fxfycxcy[0], (pixel.y - cam.fxfycxcy[3]) / cam.fxfycxcy[1], 1.f);
      rayDir = transformDirection(c2w, r);
    }
  }

  void renderPixel(uint32_t x, uint32_t y, int curFrame, bool first) {
    RayPayload p;
    p.pRec.seed = xxhash32Seed(x, y, (uint32_t)curFrame);
    vec2 jitter = curFrame == 0 ? vec2{0.5f, 0.5f} : rand2(p.pRec.seed);
    vec3 ro, rd;
    cameraRay(x, y, jitter, p.pRec.seed, ro, rd);
    p.pRec.ray = Ray{ro, rd};
    p.pRec.stop = false;
    p.pRec.radiance = vec3(0.0f);
    p.pRec.throughput = vec3(1.0f);
    p.bRec.flags = EBsdfNull;
    uint64_t nc = 0, ns = 0, ni = 0, nz = 0;
    for (p.pRec.depth = 1; p.pRec.depth <= pc.maxPathDepth;) {
      p.dRec.skip = true;
      p.dRec.radiance = vec3(0.0f);
      Hit h;
      nc++;
      if (p.pRec.depth >= 2) ni++;
      if (trace(p.pRec.ray.o, p.pRec.ray.d, MINIMUM, INFINITY_, false, h))
        closestHit(p, h);
      else
        miss(p);
      if (!p.dRec.skip) {
        float maxDist = p.dRec.dist - 2 * EPS;
        Hit sh;
        ns++;
        if (!isBlack(p.dRec.radiance)) nz++;
        if (!trace(p.dRec.ray.o, p.dRec.ray.d, 0.0f, maxDist, true, sh)) p.pRec.radiance += p.dRec.radiance;
      }
      if (p.pRec.stop) break;
      p.pRec.depth++;
    }
    n_closest.fetch_add(nc, std::memory_order_relaxed);
    n_shadow.fetch_add(ns, std::memory_order_relaxed);
    n_shadow_nz.fetch_add(nz, std::memory_order_relaxed);
    n_incoherent.fetch_add(ni, std::memory_order_relaxed);

    // rgen:134-149
    const float stddev = 0.5f, radius = 4 * stddev, alpha = -1.0f / (2.0f * stddev * stddev);
    const float expXY = std::exp(alpha * radius * radius);
    vec2 off = {jitter.x - 0.5f, jitter.y - 0.5f};
    float filterWeight = std::fmax(0.0f, std::exp(alpha * off.x * off.x) - expXY) *
                         std::fmax(0.0f, std::exp(alpha * off.y * off.y) - expXY);
    vec3 L = clamp3(p.pRec.radiance, 0.0f, 10.0f);
    vec3 radianceWeightSum = filterWeight * L;
    float filterWeightSum = filterWeight;

    size_t px = 4 * ((size_t)y * W + x);
    float* I0 = &images[0][px];
    float* I8 = &images[8][px];
    if (curFrame == 0) {  // rgen:155-162
      for (uint32_t cid = 0; cid < pc.nMultiChannel && cid < ASUNA_NUM_OUTPUT_IMAGES - 1; cid++) {
        float* Ic = &images[cid + 1][px];
        Ic[0] = p.channel[cid].x, Ic[1] = p.channel[cid].y, Ic[2] = p.channel[cid].z, Ic[3] = 1.f;
      }
    }
    if (first) {
      vec3 r = radianceWeightSum / filterWeightSum;
      I0[0] = r.x, I0[1] = r.y, I0[2] = r.z, I0[3] = 1.f;
      I8[0] = I8[1] = I8[2] = I8[3] = filterWeightSum;
    } else {  // rgen:171-178
      vec3 oldRadiance(I0[0], I0[1], I0[2]);
      float oldW = I8[0];
      vec3 oldSum = oldRadiance * oldW;
      float newW = oldW + filterWeightSum;
      vec3 newSum = oldSum + radianceWeightSum;
      vec3 r = newSum / newW;
      I0[0] = r.x, I0[1] = r.y, I0[2] = r.z, I0[3] = 1.f;
      I8[0] = I8[1] = I8[2] = I8[3] = newW;
    }
  }
};

// ============================================================================ C ABI
extern "C" {

void oracle_abi_sizes(uint32_t out[6]) {
  out[0] = sizeof(AsunaVertex), out[1] = sizeof(AsunaMaterial), out[2] = sizeof(AsunaLight);
  out[3] = sizeof(AsunaCamera), out[4] = sizeof(AsunaState), out[5] = sizeof(AsunaSunSky);
}
int oracle_create(oracle_ctx** out, int) {
  *out = new oracle_ctx();
  (*out)->pc.curFrame = -1;
  (*out)->pc.spp = 1;
  (*out)->pc.maxPathDepth = 3;
  (*out)->pc.envMapIntensity = 1.f;
  float id[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  std::memcpy((*out)->cam.envTransform, id, sizeof id);
  return 0;
}
void oracle_destroy(oracle_ctx* c) { delete c; }
const char* oracle_last_error(oracle_ctx* c) { return c->err.c_str(); }
int oracle_set_threads(oracle_ctx* c, int n) {
  c->threads = n;
  return 0;
}
int oracle_set_film(oracle_ctx* c, uint32_t w, uint32_t h) {
  c->W = w, c->H = h;
  for (auto& im : c->images) im.assign((size_t)w * h * 4, 0.0f);
  return 0;
}
static void fill_tex(Texture& t, const float* p, uint32_t w, uint32_t h) {
  t.w = w, t.h = h;
  t.rgba.assign(p, p + (size_t)w * h * 4);
}
int oracle_add_texture(oracle_ctx* c, const float* rgba, uint32_t w, uint32_t h) {
  c->textures.emplace_back();
  fill_tex(c->textures.back(), rgba, w, h);
  return (int)c->textures.size() - 1;
}
int oracle_set_envmap(oracle_ctx* c, const float* rgba, const float* marg, const float* cond, uint32_t w, uint32_t h) {
  fill_tex(c->env[0], rgba, w, h);
  fill_tex(c->env[1], marg, w, h);
  fill_tex(c->env[2], cond, w, h);
  return 0;
}
int oracle_add_mesh(oracle_ctx* c, const AsunaVertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni) {
  c->meshes.emplace_back();
  Mesh& m = c->meshes.back();
  m.v.assign(v, v + nv);
  m.idx.assign(idx, idx + (ni / 3) * 3);
  for (uint32_t i : m.idx)
    if (i >= nv) {
      c->err = "index out of range";
      c->meshes.pop_back();
      return ASUNA_E_INVALID;
    }
  c->built = false;
  return (int)c->meshes.size() - 1;
}
int oracle_add_material(oracle_ctx* c, const AsunaMaterial* m) {
  c->materials.push_back(*m);
  return (int)c->materials.size() - 1;
}
int oracle_set_lights(oracle_ctx* c, const AsunaLight* l, uint32_t n) {
  c->lights.assign(l, l + n);
  return 0;
}
int oracle_add_instance(oracle_ctx* c, const float x[16], uint32_t mesh, uint32_t material, int32_t light) {
  if (mesh >= c->meshes.size() || (light < 0 && material >= c->materials.size())) {
    c->err = "instance refers to unknown mesh/material";
    return ASUNA_E_INVALID;
  }
  Instance in{};
  // column-major 4x4 -> 3x4 rows (VkTransformMatrixKHR, nvvk::toTransformMatrixKHR)
  for (int r = 0; r < 3; r++)
    for (int col = 0; col < 4; col++) in.o2w[r * 4 + col] = x[col * 4 + r];
  // world-to-object: inverse of the affine part, computed in double
  double a[9], inv[9];
  for (int r = 0; r < 3; r++)
    for (int col = 0; col < 3; col++) a[r * 3 + col] = in.o2w[r * 4 + col];
  double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  double id = 1.0 / det;
  inv[0] = (a[4] * a[8] - a[5] * a[7]) * id, inv[1] = (a[2] * a[7] - a[1] * a[8]) * id, inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  inv[3] = (a[5] * a[6] - a[3] * a[8]) * id, inv[4] = (a[0] * a[8] - a[2] * a[6]) * id, inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  inv[6] = (a[3] * a[7] - a[4] * a[6]) * id, inv[7] = (a[1] * a[6] - a[0] * a[7]) * id, inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  for (int r = 0; r < 3; r++) {
    for (int col = 0; col < 3; col++) in.w2o[r * 4 + col] = (float)inv[r * 3 + col];
    in.w2o[r * 4 + 3] = (float)-(inv[r * 3 + 0] * in.o2w[3] + inv[r * 3 + 1] * in.o2w[7] + inv[r * 3 + 2] * in.o2w[11]);
  }
  in.mesh = mesh, in.material = material, in.light = light;
  c->instances.push_back(in);
  c->built = false;
  return (int)c->instances.size() - 1;
}
int oracle_build_accel(oracle_ctx* c, float* ms) {
  auto t0 = std::chrono::steady_clock::now();
  for (Mesh& m : c->meshes) {
    std::vector<Box> boxes(m.idx.size() / 3);
    m.box = Box();
    for (size_t t = 0; t < boxes.size(); t++) {
      for (int k = 0; k < 3; k++) boxes[t].grow(vec3(m.v[m.idx[3 * t + k]].pos));
      m.box.grow(boxes[t]);
    }
    m.bvh.build(boxes);
  }
  std::vector<Box> ib(c->instances.size());
  for (size_t i = 0; i < ib.size(); i++) {
    Instance& in = c->instances[i];
    const Box& mb = c->meshes[in.mesh].box;
    in.box = Box();
    for (int k = 0; k < 8; k++) {
      vec3 p((k & 1) ? mb.hi.x : mb.lo.x, (k & 2) ? mb.hi.y : mb.lo.y, (k & 4) ? mb.hi.z : mb.lo.z);
      in.box.grow(xf_point(in.o2w, p));
    }
    ib[i] = in.box;
  }
  c->tlas.build(ib);
  c->built = true;
  float t = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  c->stats.build_ms = t;
  if (ms) *ms = t;
  return 0;
}
int oracle_set_camera(oracle_ctx* c, const AsunaCamera* cam) {
  c->cam = *cam;
  return 0;
}
int oracle_set_sunsky(oracle_ctx* c, const AsunaSunSky* s) {
  c->sunsky = *s;
  return 0;
}
int oracle_set_state(oracle_ctx* c, const AsunaState* s) {
  if (s->spp != 1) {
    c->err = "spp must be 1 per frame (reference tracer.cpp:211)";
    return ASUNA_E_INVALID;
  }
  c->pc = *s;
  return 0;
}
int oracle_reset_frame(oracle_ctx* c) {
  c->pc.curFrame = -1;
  c->have_accum = false;
#ifdef ASUNA_REF_SHADERS
  std::fill(c->images[0].begin(), c->images[0].end(), 0.0f);
  std::fill(c->images[8].begin(), c->images[8].end(), 0.0f);
#endif
  return 0;
}
int oracle_set_partition(oracle_ctx* c, uint32_t rank, uint32_t world) {
  if (world == 0 || rank >= world) return ASUNA_E_INVALID;
  c->rank = rank, c->world = world;
  return 0;
}
#ifdef ASUNA_REF_SHADERS
// traceRayEXT's geometric query, answered by the oracle BVH for the generated reference shaders.
static int ref_trace_cb(void* ctx, const float o[3], const float d[3], float tmin, float tmax, int any, RefHit* out) {
  oracle_ctx* c = static_cast<oracle_ctx*>(ctx);
  Hit h;
  if (any)
    c->n_shadow.fetch_add(1, std::memory_order_relaxed);
  else
    c->n_closest.fetch_add(1, std::memory_order_relaxed);
  if (!c->trace(vec3(o), vec3(d), tmin, tmax, any != 0, h)) return 0;
  out->t = h.t, out->b1 = h.b1, out->b2 = h.b2, out->inst = h.inst, out->prim = h.prim;
  std::memcpy(out->o2w, c->instances[h.inst].o2w, sizeof out->o2w);
  std::memcpy(out->w2o, c->instances[h.inst].w2o, sizeof out->w2o);
  return 1;
}
static void ref_bind(oracle_ctx* c) {
  static std::vector<RefTex> tex;
  static std::vector<RefInst> inst;
  tex.resize(c->textures.size());
  for (size_t i = 0; i < tex.size(); i++) tex[i] = RefTex{c->textures[i].rgba.data(), (int32_t)c->textures[i].w, (int32_t)c->textures[i].h};
  inst.resize(c->instances.size());
  for (size_t i = 0; i < inst.size(); i++) {
    const Instance& in = c->instances[i];
    const Mesh& m = c->meshes[in.mesh];
    const AsunaMaterial* mat = in.light >= 0 ? nullptr : &c->materials[in.material];
    inst[i] = RefInst{m.v.data(), m.idx.data(), mat, in.light, mat ? mat->type : (uint32_t)ASUNA_MAT_LAMBERTIAN};
  }
  RefBind b{};
  b.ctx = c, b.trace = ref_trace_cb;
  b.camera = &c->cam, b.sunsky = &c->sunsky, b.pc = &c->pc, b.lights = c->lights.data();
  b.n_textures = (uint32_t)tex.size(), b.textures = tex.data();
  for (int i = 0; i < 3; i++) b.env[i] = RefTex{c->env[i].rgba.data(), (int32_t)c->env[i].w, (int32_t)c->env[i].h};
  b.n_instances = (uint32_t)inst.size(), b.instances = inst.data();
  for (int i = 0; i < ASUNA_NUM_OUTPUT_IMAGES; i++) b.images[i] = c->images[i].data();
  b.width = c->W, b.height = c->H;
  refglsl_bind(&b);
}
int oracle_is_reference_glsl(void) { return 1; }
#else
int oracle_is_reference_glsl(void) { return 0; }
#endif

int oracle_render_frames(oracle_ctx* c, uint32_t n) {
  if (!c->built || c->W == 0) {
    c->err = "render before build_accel/set_film";
    return ASUNA_E_INVALID;
  }
  auto t0 = std::chrono::steady_clock::now();
  int nthreads = c->threads;
  for (uint32_t f = 0; f < n; f++) {
    c->pc.curFrame++;
    int cf = c->pc.curFrame;
    if ((uint32_t)cf % c->world != c->rank) continue;
    bool first = !c->have_accum;
#ifdef ASUNA_REF_SHADERS
    // rgen:155-178 decides replace-vs-accumulate on curFrame == 0; a partition whose first frame is not 0
    // accumulates onto the planes zeroed by reset_frame, which is the same arithmetic as `first` below.
    (void)first;
    ref_bind(c);
    parallel_for((int64_t)c->H, 2, nthreads, [&](int64_t y) {
      for (uint32_t x = 0; x < c->W; x++) refglsl_render_pixel(x, (uint32_t)y);
    });
#else
    parallel_for((int64_t)c->H, 2, nthreads, [&](int64_t y) {
      for (uint32_t x = 0; x < c->W; x++) c->renderPixel(x, (uint32_t)y, cf, first);
    });
#endif
    c->have_accum = true;
    c->stats.paths += (uint64_t)c->W * c->H;
  }
  c->stats.total_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}
int oracle_sync(oracle_ctx*) { return 0; }
int oracle_read_channel(oracle_ctx* c, int ch, float* out) {
  if (ch < 0 || ch >= ASUNA_NUM_OUTPUT_IMAGES) return ASUNA_E_INVALID;
  std::memcpy(out, c->images[ch].data(), c->images[ch].size() * sizeof(float));
  return 0;
}
int oracle_read_channel_async(oracle_ctx* c, int ch, float* out) { return oracle_read_channel(c, ch, out); }  // nothing to overlap on the CPU
int oracle_wait_reads(oracle_ctx*) { return 0; }
// post.idle.frag:71-133 + utils/tonemapping.glsl over image 0 (≙ asuna_post_process).  In libref.so the fragment shader
// itself runs (refglsl_post_process).
int oracle_post_process(oracle_ctx* c, const AsunaPost* tm, float* out) {
  if (!tm || !out || c->W == 0 || tm->tmType >= ASUNA_TM_NUM) return ASUNA_E_INVALID;
  if (tm->tmType == ASUNA_TM_CUSTOM && (tm->autoExposure & 2)) return ASUNA_E_UNSUPPORTED;
#ifdef ASUNA_REF_SHADERS
  refglsl_post_process(c->images[0].data(), c->W, c->H, tm, out);
#else
  const size_t n = (size_t)c->W * c->H;
  auto srgb = [](vec3 v) { return pow3(v, 1.0f / 2.2f); };
  auto lin = [](vec3 v) { return pow3(v, 2.2f); };
  auto unch = [](vec3 v) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((v * (A * v + C * B) + D * E) / (v * (A * v + B) + D * F)) - E / F;
  };
  auto hejl = [](vec3 v) {
    v = vec3(std::fmax(0.0f, v.x - 0.004f), std::fmax(0.0f, v.y - 0.004f), std::fmax(0.0f, v.z - 0.004f));
    return (v * (6.2f * v + 0.5f)) / (v * (6.2f * v + 1.7f) + 0.06f);
  };
  auto pbrt = [](float x) { return x < 0.0031308f ? 12.92f * x : 1.055f * std::pow(x, 1.0f / 2.4f) - 0.055f; };
  float avg_lum = 1.0f;
  if (tm->tmType == ASUNA_TM_CUSTOM && (tm->autoExposure & 1)) {
    double s[3] = {0, 0, 0};
    for (size_t i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) s[k] += c->images[0][4 * i + k];
    avg_lum = 0.2126f * (float)(s[0] / n) + 0.7152f * (float)(s[1] / n) + 0.0722f * (float)(s[2] / n);
  }
  for (size_t i = 0; i < n; i++) {
    const float* in = &c->images[0][4 * i];
    const uint32_t x = (uint32_t)(i % c->W), y = (uint32_t)(i / c->W);
    vec3 hdr(in[0], in[1], in[2]), o;
    switch (tm->tmType) {
      case ASUNA_TM_NONE: o = hdr; break;
      case ASUNA_TM_GAMMA: o = srgb(hdr / (hdr / 1.5f + 1.0f)); break;
      case ASUNA_TM_REINHARD:
      case ASUNA_TM_FILMIC: o = hejl(hdr); break;
      case ASUNA_TM_ACES: {
        vec3 v = (hdr * (2.51f * hdr + 0.03f)) / (hdr * (2.43f * hdr + 0.59f) + 0.14f);
        o = srgb(clamp3(v, 0.0f, 1.0f));
        break;
      }
      case ASUNA_TM_PBRT: o = vec3(pbrt(hdr.x), pbrt(hdr.y), pbrt(hdr.z)); break;
      default: {
        vec3 v = hdr;
        if (tm->autoExposure & 1) {
          const float Yxyz = 0.3575761f * v.x + 0.7151522f * v.y + 0.1191920f * v.z;
          const float Y = (tm->key / avg_lum) * Yxyz;
          const float Yd = (Y * (1.0f + Y / (tm->Ywhite * tm->Ywhite))) / (1.0f + Y);
          v = v / Yxyz * Yd;
        }
        v = unch(v * tm->avgLum * 2.0f);
        v = srgb(v * (vec3(1.0f) / unch(vec3(11.2f))));
        uint32_t vx = x * 1664525u + 1013904223u, vy = y * 1664525u + 1013904223u, vz = 1013904223u;
        vx += vy * vz, vy += vz * vx, vz += vx * vy;
        vx ^= vx >> 16, vy ^= vy >> 16, vz ^= vz >> 16;
        vx += vy * vz, vy += vz * vx, vz += vx * vy;
        auto u2f = [](uint32_t r) { return intBitsToFloat((int32_t)(0x3f800000u | (r >> 9))) - 1.0f; };
        const vec3 noise(u2f(vx), u2f(vy), u2f(vz));
        const float q = 1.0f / 255.0f;
        const vec3 l = lin(v), sg = srgb(l);
        const vec3 c0(std::floor(sg.x / q) * q, std::floor(sg.y / q) * q, std::floor(sg.z / q) * q), c1 = c0 + q;
        const vec3 l0 = lin(c0), l1 = lin(c1);
        const vec3 d(mixf(l0.x, l1.x, noise.x), mixf(l0.y, l1.y, noise.y), mixf(l0.z, l1.z, noise.z));
        v = vec3(d.x < l.x ? c1.x : c0.x, d.y < l.y ? c1.y : c0.y, d.z < l.z ? c1.z : c0.z);
        v = clamp3(vec3(mixf(0.5f, v.x, tm->contrast), mixf(0.5f, v.y, tm->contrast), mixf(0.5f, v.z, tm->contrast)), 0.0f, 1.0f);
        v = pow3(v, 1.0f / tm->brightness);
        const float g = 0.299f * v.x + 0.587f * v.y + 0.114f * v.z;
        v = vec3(mixf(g, v.x, tm->saturation), mixf(g, v.y, tm->saturation), mixf(g, v.z, tm->saturation));
        const float ux = ((((float)x + 0.5f) / (float)c->W) * tm->renderingRatio[0] - 0.5f) * 2.0f;
        const float uy = ((((float)y + 0.5f) / (float)c->H) * tm->renderingRatio[1] - 0.5f) * 2.0f;
        o = v * (1.0f - (ux * ux + uy * uy) * tm->vignette);
      }
    }
    out[4 * i] = o.x, out[4 * i + 1] = o.y, out[4 * i + 2] = o.z, out[4 * i + 3] = in[3];
  }
#endif
  return 0;
}
int oracle_export_partial(oracle_ctx* c, void** out) {
  size_t n = (size_t)c->W * c->H;
  c->partial.resize(n * 4);
  for (size_t i = 0; i < n; i++) {
    float w = c->have_accum ? c->images[8][4 * i] : 0.0f;
    for (int k = 0; k < 3; k++) c->partial[4 * i + k] = c->have_accum ? c->images[0][4 * i + k] * w : 0.0f;
    c->partial[4 * i + 3] = w;
  }
  *out = c->partial.data();
  return 0;
}
int oracle_import_partial(oracle_ctx* c) {
  size_t n = (size_t)c->W * c->H;
  for (size_t i = 0; i < n; i++) {
    float w = c->partial[4 * i + 3];
    for (int k = 0; k < 3; k++) c->images[0][4 * i + k] = c->partial[4 * i + k] / w;
    c->images[0][4 * i + 3] = 1.f;
    for (int k = 0; k < 4; k++) c->images[8][4 * i + k] = w;
  }
  c->have_accum = true;
  return 0;
}
int oracle_get_stats(oracle_ctx* c, AsunaStats* s) {
  c->stats.closest_rays = c->n_closest, c->stats.shadow_rays = c->n_shadow;
  c->stats.incoherent_closest_rays = c->n_incoherent;
  c->stats.node_visits = c->n_node_visits, c->stats.tri_tests = c->n_tri_tests;
  *s = c->stats;
  return 0;
}
int oracle_reset_stats(oracle_ctx* c) {
  float b = c->stats.build_ms;
  c->stats = AsunaStats{};
  c->stats.build_ms = b;
  c->n_closest = c->n_shadow = c->n_shadow_nz = c->n_incoherent = 0;
  c->n_node_visits = c->n_tri_tests = 0;
  return 0;
}
// BVH2 node visits, triangle tests (instrumented traversal) and the number of shadow rays
// whose NEE contribution was non-zero (the ones a renderer actually needs), since the last reset.
int oracle_traversal_counters(oracle_ctx* c, uint64_t out[3]) {
  out[0] = c->n_node_visits, out[1] = c->n_tri_tests, out[2] = c->n_shadow_nz;
  return 0;
}
int oracle_trace_rays(oracle_ctx* c, const float* rays, uint32_t n, float* tuv, uint32_t* ip) {
  if (!c->built) return ASUNA_E_INVALID;
  parallel_for((int64_t)n, 1024, c->threads, [&](int64_t i) {
    const float* r = rays + 8 * i;
    Hit h;
    bool f = c->trace(vec3(r[0], r[1], r[2]), vec3(r[4], r[5], r[6]), r[3], r[7], false, h);
    tuv[3 * i + 0] = f ? h.t : 0.f, tuv[3 * i + 1] = h.b1, tuv[3 * i + 2] = h.b2;
    ip[2 * i + 0] = h.inst, ip[2 * i + 1] = h.prim;
  });
  return 0;
}
int oracle_occlusion_rays(oracle_ctx* c, const float* rays, uint32_t n, uint8_t* occ) {
  if (!c->built) return ASUNA_E_INVALID;
  parallel_for((int64_t)n, 1024, c->threads, [&](int64_t i) {
    const float* r = rays + 8 * i;
    Hit h;
    occ[i] = c->trace(vec3(r[0], r[1], r[2]), vec3(r[4], r[5], r[6]), r[3], r[7], true, h) ? 1 : 0;
  });
  return 0;
}
int oracle_trace_primary(oracle_ctx* c, uint32_t* ip, float* t) {
  if (!c->built) return ASUNA_E_INVALID;
  parallel_for((int64_t)c->H, 2, c->threads, [&](int64_t y) {
    for (uint32_t x = 0; x < c->W; x++) {
      uint32_t seed = xxhash32Seed(x, (uint32_t)y, 0);
      vec3 o, d;
      c->cameraRay(x, (uint32_t)y, vec2{0.5f, 0.5f}, seed, o, d);
      Hit h;
      bool f = c->trace(o, d, MINIMUM, INFINITY_, false, h);
      size_t i = (size_t)y * c->W + x;
      ip[2 * i] = h.inst, ip[2 * i + 1] = h.prim;
      t[i] = f ? h.t : 0.f;
    }
  });
  return 0;
}

// ---- unit-test hooks: expose individual restated functions so tests can pin them ----
// SAH cost of mesh `mesh`'s binned-SAH BVH2 after the optimal 8-wide collapse (bvh.h wide_sah_cost): the quality
// yardstick for the GPU builder (tests/test_gpu_kernels.py).
double oracle_wide_sah_cost(oracle_ctx* c, uint32_t mesh, double c_node, double c_prim, uint32_t pmax) {
  if (mesh >= c->meshes.size()) return -1.0;
  return c->meshes[mesh].bvh.wide_sah_cost(c_node, c_prim, pmax);
}
// n independent closest-hit (inst/prim/b1/b2 given) or miss (inst = 0xFFFFFFFF) shader invocations on caller-made payloads.
static int shade_probe_one(oracle_ctx* c, uint32_t inst, uint32_t prim, float b1, float b2, ShadeProbe* q) {
  const bool is_hit = inst != 0xFFFFFFFFu;
  if (is_hit && (inst >= c->instances.size() || 3 * (size_t)prim >= c->meshes[c->instances[inst].mesh].idx.size())) return ASUNA_E_INVALID;
#ifdef ASUNA_REF_SHADERS
  RefHit h{};
  if (is_hit) {
    h.b1 = b1, h.b2 = b2, h.inst = inst, h.prim = prim;
    std::memcpy(h.o2w, c->instances[inst].o2w, sizeof h.o2w);
    std::memcpy(h.w2o, c->instances[inst].w2o, sizeof h.w2o);
  }
  refglsl_shade_probe(is_hit ? &h : nullptr, q);
#else
  auto st = [](float* f, vec3 v) { f[0] = v.x, f[1] = v.y, f[2] = v.z; };
  RayPayload p;
  p.pRec.ray = Ray{vec3(q->ray_o), vec3(q->ray_d)};
  p.pRec.radiance = vec3(q->radiance), p.pRec.throughput = vec3(q->throughput);
  p.pRec.depth = (int)q->depth, p.pRec.seed = q->seed, p.pRec.stop = q->stop != 0;
  p.bRec.d = vec3(q->brec_d), p.bRec.pdf = q->brec_pdf, p.bRec.flags = q->brec_flags;
  p.dRec.skip = true;
  p.dRec.radiance = vec3(0.0f);
  p.dRec.dist = 0, p.dRec.ray = Ray{vec3(0.0f), vec3(0.0f)};
  for (auto& ch : p.channel) ch = vec3(0.0f);
  if (is_hit) {
    Hit h;
    h.b1 = b1, h.b2 = b2, h.inst = inst, h.prim = prim;
    c->closestHit(p, h);
  } else {
    c->miss(p);
  }
  st(q->ray_o, p.pRec.ray.o), st(q->ray_d, p.pRec.ray.d), st(q->radiance, p.pRec.radiance), st(q->throughput, p.pRec.throughput);
  q->depth = (uint32_t)p.pRec.depth, q->seed = p.pRec.seed, q->stop = p.pRec.stop ? 1u : 0u;
  st(q->brec_d, p.bRec.d), q->brec_pdf = p.bRec.pdf, q->brec_flags = p.bRec.flags;
  st(q->drec_radiance, p.dRec.radiance), q->drec_dist = p.dRec.dist;
  st(q->drec_o, p.dRec.ray.o), st(q->drec_d, p.dRec.ray.d), q->drec_skip = p.dRec.skip ? 1u : 0u;
  for (int i = 0; i < 8; i++) st(q->channel[i], p.channel[i]);
#endif
  return 0;
}
int oracle_shade_probes(oracle_ctx* c, uint32_t n, const uint32_t* inst, const uint32_t* prim, const float* b1,
                        const float* b2, ShadeProbe* q) {
#ifdef ASUNA_REF_SHADERS
  ref_bind(c);
#endif
  for (uint32_t i = 0; i < n; i++) {
    int rc = shade_probe_one(c, inst[i], prim[i], b1[i], b2[i], q + i);
    if (rc) return rc;
  }
  return 0;
}
uint32_t oracle_xxhash32(uint32_t x, uint32_t y, uint32_t z) { return xxhash32Seed(x, y, z); }
uint32_t oracle_pcg(uint32_t* state) { return pcg(*state); }
float oracle_rand(uint32_t* state) { return rand1(*state); }
void oracle_sun_and_sky(const AsunaSunSky* ss, const float d[3], float out[3]) {
  vec3 r = sun_and_sky(*ss, vec3(d));
  out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void oracle_offset_position(const float p[3], const float n[3], float out[3]) {
  vec3 r = offsetPositionAlongNormal(vec3(p), vec3(n));
  out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void oracle_basis(const float n[3], float f[3], float r[3]) {
  vec3 F, R;
  basis(vec3(n), F, R);
  f[0] = F.x, f[1] = F.y, f[2] = F.z, r[0] = R.x, r[1] = R.y, r[2] = R.z;
}
void oracle_concentric_disk(const float u[2], float out[2]) {
  vec2 r = concentricSampleDisk(vec2{u[0], u[1]});
  out[0] = r.x, out[1] = r.y;
}
void oracle_cosine_hemisphere(const float u[2], float out[3]) {
  vec3 r = cosineSampleHemisphere(vec2{u[0], u[1]});
  out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void oracle_uniform_sphere(const float u[2], float out[3]) {
  vec3 r = uniformSampleSphere(vec2{u[0], u[1]});
  out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
float oracle_power_heuristic(float a, float b) { return powerHeuristic(a, b); }
// out = radiance(3) d(3) n(3) dist pdf flags, like refglsl_sample_one_light
void oracle_sample_one_light(const AsunaLight* light, const float r[2], const float pos[3], float out[12]) {
  oracle_ctx c;
  LightSamplingRecord rec;
  vec3 L = c.sampleOneLight(vec2{r[0], r[1]}, *light, vec3(pos), rec);
  out[0] = L.x, out[1] = L.y, out[2] = L.z, out[3] = rec.d.x, out[4] = rec.d.y, out[5] = rec.d.z;
  out[6] = rec.n.x, out[7] = rec.n.y, out[8] = rec.n.z, out[9] = rec.dist, out[10] = rec.pdf, out[11] = (float)rec.flags;
}
// env-map sampling / lookup through the context's tables (set_envmap, set_camera for envTransform, set_state)
void oracle_envmap_sample(oracle_ctx* c, const float r[2], float out[7]) {
  vec3 L;
  float pdf;
  vec3 col = c->sampleEnvmap(vec2{r[0], r[1]}, L, pdf);
  out[0] = col.x, out[1] = col.y, out[2] = col.z, out[3] = L.x, out[4] = L.y, out[5] = L.z, out[6] = pdf;
}
void oracle_envmap_eval_pdf(oracle_ctx* c, const float d[3], float out[4]) {
  vec3 col = c->evalEnvmap(vec3(d));
  out[0] = col.x, out[1] = col.y, out[2] = col.z, out[3] = c->pdfEnvmap(vec3(d));
}
void oracle_texture_bilinear(const float* rgba, uint32_t w, uint32_t h, float u, float v, float out[4]) {
  Texture t;
  fill_tex(t, rgba, w, h);
  vec4 r = textureBilinear(t, vec2{u, v});
  out[0] = r.x, out[1] = r.y, out[2] = r.z, out[3] = r.w;
}

}  // extern "C"
