// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// Restatement of the reference's GLSL utility code in scalar C++:
//   reference src/shaders/utils/math.glsl         -> RNG, sampling helpers, basis, offsetPositionAlongNormal
//   reference src/shaders/utils/structs.glsl      -> flags and records
//   reference src/shaders/utils/sample_light.glsl -> light samplers, env-map eval/pdf/sample
//   reference src/shaders/utils/sun_and_sky.glsl  -> sun & sky model
// Each function names the lines it follows.  fp32 throughout, like the shaders.
#pragma once
#include <vector>

#include "../include/asuna_b200.h"
#include "vecmath.h"

namespace orc {

constexpr float PI = 3.14159265358979323846f;
constexpr float TWO_PI = 6.28318530717958647692f;
constexpr float INV_PI = 0.31830988618379067154f;
constexpr float INV_2PI = 0.15915494309189533577f;
constexpr float INV_4PI = 0.07957747154594766788f;
constexpr float PI_OVER_2 = 1.57079632679489661923f;
constexpr float PI_OVER_4 = 0.78539816339744830961f;
constexpr float EPS = 0.001f;          // math.glsl:13
constexpr float INFINITY_ = 1e10f;     // math.glsl:14
constexpr float MINIMUM = 0.00001f;    // math.glsl:15

// ---- RNG: math.glsl:20-44 ----
inline uint32_t xxhash32Seed(uint32_t px, uint32_t py, uint32_t pz) {
  const uint32_t P0 = 2246822519U, P1 = 3266489917U, P2 = 668265263U, P3 = 374761393U;
  uint32_t h32 = pz + P3 + px * P1;
  h32 = P2 * ((h32 << 17) | (h32 >> (32 - 17)));
  h32 += py * P1;
  h32 = P2 * ((h32 << 17) | (h32 >> (32 - 17)));
  h32 = P0 * (h32 ^ (h32 >> 15));
  h32 = P1 * (h32 ^ (h32 >> 13));
  return h32 ^ (h32 >> 16);
}
inline uint32_t pcg(uint32_t& state) {
  uint32_t prev = state * 747796405u + 2891336453u;
  uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  state = prev;
  return (word >> 22u) ^ word;
}
// 1.0/float(0xffffffffu): float(0xffffffff) rounds to 2^32, so the factor is exactly 2^-32
// and the result can be exactly 1.0f.
inline float rand1(uint32_t& seed) { return (float)pcg(seed) * (1.0f / 4294967296.0f); }
inline vec2 rand2(uint32_t& seed) {
  float a = rand1(seed);
  float b = rand1(seed);
  return {a, b};
}

// ---- flags: structs.glsl:16-59 ----
enum : uint32_t {
  EBsdfNull = 0,
  EDiffuseReflection = 1u << 0,
  EDiffuseTransmission = 1u << 1,
  EGlossyReflection = 1u << 2,
  EGlossyTransmission = 1u << 3,
  ESpecularReflection = 1u << 4,
  ESpecularTransmission = 1u << 5,
  ESmooth = EDiffuseReflection | EDiffuseTransmission | EGlossyReflection | EGlossyTransmission,
  ELightNull = 0,
  EDelta = 1u << 0,
  EArea = 1u << 1,
};
inline bool isNonSpecular(uint32_t flags) { return (flags & ESmooth) != 0; }
inline bool isBlack(vec3 v) { return length(v) == 0.0f; }

struct Ray {
  vec3 o, d;
};
struct BsdfSamplingRecord {
  vec3 d;
  float pdf = 0;
  uint32_t flags = 0;
};
struct LightSamplingRecord {
  vec3 d, n;
  float dist = 0, pdf = 0;
  uint32_t flags = 0;
};
struct DirectLightRecord {
  vec3 radiance;
  float dist = 0;
  Ray ray;
  bool skip = true;
};
struct PathRecord {
  Ray ray;
  vec3 radiance, throughput;
  int depth = 0;  // uint in GLSL; depth-- on a pass-through never goes below 0
  uint32_t seed = 0;
  bool stop = false;
};
struct RayPayload {
  PathRecord pRec;
  BsdfSamplingRecord bRec;
  vec3 channel[ASUNA_NUM_OUTPUT_IMAGES - 1];
  DirectLightRecord dRec;
};

// ---- math.glsl:46-66 ----
inline vec3 transformPoint(const mat4& M, vec3 p) {
  vec4 h = mul(M, vec4{p.x, p.y, p.z, 1.0f});
  return vec3(h.x, h.y, h.z) / h.w;
}
inline vec3 transformVector(const mat4& M, vec3 v) {
  vec4 h = mul(M, vec4{v.x, v.y, v.z, 0.0f});
  return {h.x, h.y, h.z};
}
inline vec3 makeNormal(vec3 n) {
  if (length(n) == 0.0f) return n;
  return normalize(n);
}
inline vec3 transformDirection(const mat4& M, vec3 d) { return makeNormal(transformVector(M, d)); }
inline float safeSqrt(float v) { return std::sqrt(std::fmax(0.0f, v)); }
inline vec3 toWorld(vec3 X, vec3 Y, vec3 Z, vec3 V) { return V.x * X + V.y * Y + V.z * Z; }
inline vec3 toLocal(vec3 X, vec3 Y, vec3 Z, vec3 V) { return {dot(V, X), dot(V, Y), dot(V, Z)}; }

// ---- math.glsl:133-189 ----
inline vec3 uniformSampleSphere(vec2 u) {
  float z = 1.0f - 2.0f * u.x;
  float r = std::sqrt(std::fmax(0.0f, 1.0f - z * z));
  float phi = TWO_PI * u.y;
  return {r * std::cos(phi), r * std::sin(phi), z};
}
inline float uniformSpherePdf() { return INV_4PI; }
inline vec2 concentricSampleDisk(vec2 u) {
  vec2 uo = {2.0f * u.x - 1.0f, 2.0f * u.y - 1.0f};
  if (uo.x == 0.0f && uo.y == 0.0f) return {0.0f, 0.0f};
  float theta, r;
  if (std::fabs(uo.x) > std::fabs(uo.y)) {
    r = uo.x;
    theta = PI_OVER_4 * (uo.y / uo.x);
  } else {
    r = uo.y;
    theta = PI_OVER_2 - PI_OVER_4 * (uo.x / uo.y);
  }
  return {r * std::cos(theta), r * std::sin(theta)};
}
inline vec3 cosineSampleHemisphere(vec2 u) {
  vec2 d = concentricSampleDisk(u);
  float z = std::sqrt(std::fmax(0.0f, 1.0f - d.x * d.x - d.y * d.y));
  return {d.x, d.y, z};
}
inline float cosineHemispherePdf(float cosTheta) {
  if (cosTheta <= 0.0f) return 0.0f;
  return cosTheta * INV_PI;
}
inline float powerHeuristic(float a, float b) {
  a = a * a;
  b = b * b + a;
  if (b == 0.0f) return 0.0f;
  return a / b;
}

// ---- math.glsl:199-216 (HANDLE_SINGULARITY defined) ----
inline void basis(vec3 n, vec3& f, vec3& r) {
  if (n.z < -0.999999f) {
    f = vec3(0, -1, 0);
    r = vec3(-1, 0, 0);
  } else {
    float a = 1.0f / (1.0f + n.z);
    float b = -n.x * n.y * a;
    f = vec3(1.0f - n.x * n.x * a, b, -n.x);
    r = vec3(b, 1.0f - n.y * n.y * a, -n.y);
  }
}

// ---- math.glsl:241-266 ----
inline vec3 offsetPositionAlongNormal(vec3 p, vec3 n) {
  const float int_scale = 256.0f;
  int32_t ox = (int32_t)(int_scale * n.x), oy = (int32_t)(int_scale * n.y), oz = (int32_t)(int_scale * n.z);
  vec3 pi(intBitsToFloat(floatBitsToInt(p.x) + ((p.x < 0) ? -ox : ox)),
          intBitsToFloat(floatBitsToInt(p.y) + ((p.y < 0) ? -oy : oy)),
          intBitsToFloat(floatBitsToInt(p.z) + ((p.z < 0) ? -oz : oz)));
  const float origin = 1.0f / 32.0f;
  const float floatScale = 1.0f / 65536.0f;
  return {std::fabs(p.x) < origin ? p.x + floatScale * n.x : pi.x,
          std::fabs(p.y) < origin ? p.y + floatScale * n.y : pi.y,
          std::fabs(p.z) < origin ? p.z + floatScale * n.z : pi.z};
}

// sun_and_sky.glsl:29-31 -- also the `luminance` the plastic shaders call.
inline float luminance(vec3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }

// ---- textures: sampler LINEAR/LINEAR, REPEAT, LOD 0 (reference src/core/texture.cpp:99-107).
// fp32 bilinear footprint per the Vulkan texel-addressing rules: unnormalised coordinate
// minus 0.5, floor -> i0, i1 = i0+1, both wrapped modulo the size.
struct Texture {
  uint32_t w = 0, h = 0;
  std::vector<float> rgba;
};
inline int wrapi(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}
inline vec4 textureBilinear(const Texture& t, vec2 uv) {
  float x = uv.x * (float)t.w - 0.5f, y = uv.y * (float)t.h - 0.5f;
  if (!std::isfinite(x) || !std::isfinite(y)) return {0, 0, 0, 0};
  float fx0 = std::floor(x), fy0 = std::floor(y);
  float fx = x - fx0, fy = y - fy0;
  // keep the integer conversion in range for absurd uvs
  int x0 = wrapi((int)std::fmod(fx0, (float)t.w), (int)t.w), y0 = wrapi((int)std::fmod(fy0, (float)t.h), (int)t.h);
  int x1 = wrapi(x0 + 1, (int)t.w), y1 = wrapi(y0 + 1, (int)t.h);
  const float* p00 = &t.rgba[4 * ((size_t)y0 * t.w + x0)];
  const float* p10 = &t.rgba[4 * ((size_t)y0 * t.w + x1)];
  const float* p01 = &t.rgba[4 * ((size_t)y1 * t.w + x0)];
  const float* p11 = &t.rgba[4 * ((size_t)y1 * t.w + x1)];
  float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
  vec4 r;
  r.x = w00 * p00[0] + w10 * p10[0] + w01 * p01[0] + w11 * p11[0];
  r.y = w00 * p00[1] + w10 * p10[1] + w01 * p01[1] + w11 * p11[1];
  r.z = w00 * p00[2] + w10 * p10[2] + w01 * p01[2] + w11 * p11[2];
  r.w = w00 * p00[3] + w10 * p10[3] + w01 * p01[3] + w11 * p11[3];
  return r;
}

// ---- sun & sky: sun_and_sky.glsl.  Same arithmetic, written as one Perez evaluator that the
// luminance and the two chromaticity channels share (the reference spells it out three times,
// :174-177, :185-188, :220-223).
inline float perez(float A, float B, float C, float D, float E, float cos_theta, float gamma, float cos_gamma,
                   float theta_sun, float cos_theta_sun) {
  return ((1.0f + A * std::exp(B / cos_theta)) * (1.0f + C * std::exp(D * gamma) + E * cos_gamma * cos_gamma)) /
         ((1.0f + A * std::exp(B / 1.0f)) * (1.0f + C * std::exp(D * theta_sun) + E * cos_theta_sun * cos_theta_sun));
}
// :201-226
inline float sky_luminance(vec3 dir, vec3 sun, float T) {
  float cg = dot(sun, dir);
  if (cg < 0.0f) cg = 0.0f;
  if (cg > 1.0f) cg = 2.0f - cg;
  float gamma = std::acos(cg);
  float ts = std::acos(sun.z);
  return perez(0.178721f * T - 1.463037f, -0.355402f * T + 0.427494f, -0.022669f * T + 5.325056f,
               0.120647f * T - 2.577052f, -0.066967f * T + 0.370275f, dir.z, gamma, cg, ts, sun.z);
}
// :134-199
inline vec3 sky_color_xyz(vec3 dir, vec3 sun, float T, float lum) {
  float cg = dot(sun, dir);
  if (cg > 1.0f) cg = 2.0f - cg;
  float gamma = std::acos(cg);
  float cts = sun.z;
  float ts = std::acos(cts);
  float t2 = T * T, ts2 = ts * ts, ts3 = ts2 * ts;
  float zx = ((+0.001650f * ts3 - 0.003742f * ts2 + 0.002088f * ts + 0) * t2 +
              (-0.029028f * ts3 + 0.063773f * ts2 - 0.032020f * ts + 0.003948f) * T +
              (+0.116936f * ts3 - 0.211960f * ts2 + 0.060523f * ts + 0.258852f));
  float zy = ((+0.002759f * ts3 - 0.006105f * ts2 + 0.003162f * ts + 0) * t2 +
              (-0.042149f * ts3 + 0.089701f * ts2 - 0.041536f * ts + 0.005158f) * T +
              (+0.153467f * ts3 - 0.267568f * ts2 + 0.066698f * ts + 0.266881f));
  float x = perez(-0.019257f * T - (0.29f - std::pow(cts, 0.5f) * 0.09f), -0.066513f * T + 0.000818f,
                  -0.000417f * T + 0.212479f, -0.064097f * T - 0.898875f, -0.003251f * T + 0.045178f, dir.z, gamma,
                  cg, ts, cts);
  float y = perez(-0.016698f * T - 0.260787f, -0.094958f * T + 0.009213f, -0.007928f * T + 0.210230f,
                  -0.044050f * T - 1.653694f, -0.010922f * T + 0.052919f, dir.z, gamma, cg, ts, cts);
  const float sat = 1.0f;
  x = zx * ((x * sat) + (1.0f - sat));
  y = zy * ((y * sat) + (1.0f - sat));
  vec3 xyz;
  xyz.y = lum;
  xyz.x = (x / y) * xyz.y;
  xyz.z = ((1.0f - x - y) / y) * xyz.y;
  return xyz;
}
constexpr float SS_PI = 3.1415926535f;  // sun_and_sky.glsl:24
// :228-243
inline vec3 calc_env_color(vec3 sun, vec3 dir, float T) {
  float ts = std::acos(sun.z);
  float chi = (4.0f / 9.0f - T / 120.0f) * (SS_PI - 2.0f * ts);
  float lum = 1000.0f * ((4.0453f * T - 4.9710f) * std::tan(chi) - 0.2155f * T + 2.4192f);
  lum *= sky_luminance(dir, sun, T);
  vec3 XYZ = sky_color_xyz(dir, sun, T, lum);
  vec3 c(3.241f * XYZ.x - 1.537f * XYZ.y - 0.499f * XYZ.z, -0.969f * XYZ.x + 1.876f * XYZ.y + 0.042f * XYZ.z,
         0.056f * XYZ.x - 0.204f * XYZ.y + 1.057f * XYZ.z);
  return c * SS_PI;
}
// :108-132
inline vec3 calc_sun_color(vec3 sun, float T) {
  if (!(sun.z > 0.0f)) return vec3(0.0f);
  const vec3 ko(12.0f, 8.5f, 0.9f), wl(0.610f, 0.550f, 0.470f);
  const vec3 solRad(1.0f * 127500 / 0.9878f, 0.992f * 127500 / 0.9878f, 0.911f * 127500 / 0.9878f);
  float m = 1.0f / (sun.z + 0.15f * std::pow(93.885f - std::acos(sun.z) * 180 / SS_PI, -1.253f));
  float beta = 0.04608f * T - 0.04586f;
  vec3 ta = exp3(-m * beta * pow3(wl, -1.3f));
  vec3 to = exp3(-m * ko * 0.0035f);
  vec3 tr = exp3(-m * 0.008735f * pow3(wl, -4.08f));
  return tr * ta * to * solRad;
}
// :33-58, :60-106 -- cosine-weighted direction around +z for the 5x5 ground-irradiance stencil
inline vec3 xyz2dir(vec3 main, float x, float y, float z) {
  vec3 u;
  if (std::fabs(main.x) < std::fabs(main.y))
    u = vec3(0.0f, -main.z, main.y);
  else
    u = vec3(main.z, 0.0f, -main.x);
  u = normalize(u);  // the degenerate re-derivation at :48-54 recomputes the same vector
  vec3 v = cross(main, u);
  return x * u + y * v + z * main;
}
inline vec3 diffuse_dir_about(vec3 normal, float sx, float sy) {
  float lx = 2 * sx - 1, ly = 2 * sy - 1, r = 0.0f, phi = 0.0f;
  if (!(lx == 0.0f && ly == 0.0f)) {
    if (lx > -ly) {
      if (lx > ly) {
        r = lx;
        phi = (SS_PI / 4.0f) * (1.0f + ly / lx);
      } else {
        r = ly;
        phi = (SS_PI / 4.0f) * (3.0f - lx / ly);
      }
    } else {
      if (lx < ly) {
        r = -lx;
        phi = (SS_PI / 4.0f) * (5.0f + ly / lx);
      } else {
        r = -ly;
        phi = (SS_PI / 4.0f) * (7.0f - lx / ly);
      }
    }
  }
  float x = r * std::cos(phi), y = r * std::sin(phi);
  float z2 = 1.0f - x * x - y * y;
  float z = z2 > 0.0f ? std::sqrt(z2) : 0.0f;
  return xyz2dir(normal, x, y, z);
}
// :245-262 -- float loop counters exactly as written (u,v = 0.1, 0.3, ... while < 1)
inline vec3 calc_irrad(vec3 sun, float T) {
  vec3 acc(0.0f);
  for (float u = 1.f / 10.f; u < 1.f; u += 1.f / 5.f)
    for (float v = 1.f / 10.f; v < 1.f; v += 1.f / 5.f)
      acc += calc_env_color(sun, diffuse_dir_about(vec3(0, 0, 1), u, v), T);
  return acc / 25.0f;
}
// :264-276
inline float tweak_saturation(float s, float haze) {
  float lowsat = std::pow(s, 3.0f);
  if (s <= 1.0f) {
    float h = (haze - 2.0f) / 15.0f;
    h = clampf(h, 0.0f, 1.0f);
    h = std::pow(h, 3.0f);
    return (s * (1.0f - h)) + lowsat * h;
  }
  return 1.0f;
}
// :278-288
inline vec3 arch_vectortweak(vec3 d, int y_is_up, float horiz_height) {
  vec3 o = d;
  if (y_is_up == 1) o = vec3(d.x, d.z, d.y);
  if (horiz_height != 0) {
    o.z -= horiz_height;
    o = normalize(o);
  }
  return o;
}
// :290-310 (the saturation>1 clamp there writes to a dead copy, so it has no effect)
inline vec3 arch_colortweak(vec3 tint, float saturation, float redness) {
  float intensity = luminance(tint);
  vec3 o = (saturation <= 0.0f) ? vec3(intensity) : tint * saturation + vec3(intensity * (1.0f - saturation));
  return o * vec3(1.0f + redness, 1.0f, 1.0f - redness);
}
// :312-394
inline vec2 calc_physical_scale(float disk_scale, float glow_int, float disk_int) {
  float disk_r = 0.00465f * disk_scale;
  float glow_r = disk_r * 10.0f;
  float glow_integral = glow_int * ((4.f * SS_PI) - (24.f * SS_PI) / (glow_r * glow_r) +
                                    (24.f * SS_PI) * std::sin(glow_r) / (glow_r * glow_r * glow_r));
  float target = disk_int * SS_PI;
  float glow_scale = 1.0f;
  float max_glow = 0.5f * target;
  if (glow_integral > max_glow) {
    glow_scale *= max_glow / glow_integral;
    target -= max_glow;
  } else {
    target -= glow_integral;
  }
  float area = 2 * SS_PI * (1 - std::cos(disk_r));
  float target_intensity = target / area;
  float actual_integral = 1.0f * area;
  float actual_intensity = disk_int * 100.0f * actual_integral / area;
  return {(target_intensity == 0.0f) ? 0.0f : target_intensity / actual_intensity, glow_scale};
}
// :396-403
inline float night_brightness_adjustment(vec3 sun) {
  const float lmt = 0.30901699437494742410229341718282f;
  if (sun.z <= -lmt) return 0.0f;
  float f = (sun.z + lmt) / lmt;
  f *= f;
  f *= f;
  return f;
}
// :405-533
inline vec3 sun_and_sky(const AsunaSunSky& ss, vec3 in_direction) {
  float factor = 1.0f, night_factor = 1.0f;
  vec3 rgb_scale(ss.rgb_unit_conversion);
  float horiz_height = ss.horizon_height / 10.0f;
  vec3 dir = arch_vectortweak(in_direction, ss.y_is_up, horiz_height);
  float haze = 2.0f + ss.haze;
  if (haze < 2.0f) haze = 2.0f;
  float saturation = tweak_saturation(ss.saturation, haze);
  if (luminance(rgb_scale) < 0.0f) rgb_scale = vec3(1.0f / 80000.0f);
  rgb_scale *= ss.multiplier;
  if (ss.multiplier <= 0.0f) return vec3(0.0f);

  float downness = dir.z;
  vec3 real_dir = dir;
  if (dir.z < 0.001f) {
    dir.z = 0.001f;
    dir = normalize(dir);
  }
  vec3 sun = normalize(vec3(ss.sun_direction));
  sun = arch_vectortweak(sun, ss.y_is_up, horiz_height);
  vec3 real_sun = sun;
  if (sun.z < 0.001f) {
    if (sun.z < 0.0f) factor = night_brightness_adjustment(sun);
    sun.z = 0.001f;
    sun = normalize(sun);
  }
  vec3 tint(0.0f);
  if (factor > 0.0f) {
    tint = calc_env_color(sun, dir, haze);
    if (factor < 1.0f) tint *= factor;
  }
  vec3 sun_color = calc_sun_color(sun, downness > 0 ? haze : 2.0f);
  if (ss.sun_disk_intensity > 0.0f && ss.sun_disk_scale > 0.0f) {
    float sun_angle = std::acos(dot(real_dir, real_sun));
    float sun_radius = 0.00465f * ss.sun_disk_scale * 10.0f;
    if (sun_angle < sun_radius) {
      float disk_scale = 1.0f, glow_scale = 1.0f;
      if (ss.physically_scaled_sun == 1) {
        vec2 s = calc_physical_scale(ss.sun_disk_scale, ss.sun_glow_intensity, ss.sun_disk_intensity);
        disk_scale = s.x;
        glow_scale = s.y;
      }
      float f = (1.0f - sun_angle / sun_radius) * 10.0f;
      f = std::pow(f / 10.0f, 3.0f) * 2.0f * ss.sun_glow_intensity * glow_scale +
          smoothstepf(8.5f, 9.5f + (haze / 50.0f), f) * 100.0f * ss.sun_disk_intensity * disk_scale;
      tint += sun_color * f;
    }
  }
  vec3 out = tint * rgb_scale;
  if (downness <= 0.0f) {
    vec3 down(ss.ground_color);
    vec3 irrad = calc_irrad(sun, 2.0f);
    down *= (irrad + sun_color * sun.z) * rgb_scale;
    if (factor < 1) down *= factor;
    float blur = ss.horizon_blur / 10.0f;
    if (blur > 0.0f) {
      float d = -downness / blur;
      if (d > 1.0f) d = 1.0f;
      d = smoothstepf(0.0f, 1.0f, d);
      out = out * (1.0f - d) + down * d;
      night_factor = 1.0f - d;
    } else {
      out = down;
      night_factor = 0.0f;
    }
  }
  vec3 result = arch_colortweak(out, saturation, ss.redblueshift);
  if (night_factor > 0.0f) {
    vec3 night = vec3(ss.night_color) * night_factor;
    if (result.x < night.x) result.x = night.x;
    if (result.y < night.y) result.y = night.y;
    if (result.z < night.z) result.z = night.z;
  }
  return result * SS_PI;
}

}  // namespace orc
