// Device-side shading library of the wavefront integrator: RNG, samplers, texture / env-map
// fetches, the sun & sky model, light sampling and the eight in-scope BSDFs.
//
// Semantics follow the reference GLSL (cited per function); the structure does not: there is no
// ray payload and no shader binding table -- k_shade (integrator.cu) keeps the path in registers,
// calls one of the `shade_*` functions below and scatters the results into the SoA path state.
#pragma once
#include "device_types.cuh"
#include "vec.cuh"

namespace asuna {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInv2Pi = 0.15915494309189533577f;
constexpr float kInv4Pi = 0.07957747154594766788f;
constexpr float kPiOver2 = 1.57079632679489661923f;
constexpr float kPiOver4 = 0.78539816339744830961f;
constexpr float kEps = 0.001f;       // reference src/shaders/utils/math.glsl:13
constexpr float kInfinity = 1e10f;   // math.glsl:14
constexpr float kMinimum = 0.00001f; // math.glsl:15

// BSDF / light flags, reference src/shaders/utils/structs.glsl:16-50
enum : uint32_t {
  kBsdfNull = 0,
  kDiffuseReflection = 1u << 0,
  kGlossyReflection = 1u << 2,
  kSpecularReflection = 1u << 4,
  kSpecularTransmission = 1u << 5,
  kSmooth = (1u << 0) | (1u << 1) | (1u << 2) | (1u << 3),
  kLightDelta = 1u << 0,
  kLightArea = 1u << 1,
};

// ---- RNG: math.glsl:20-44 ------------------------------------------------------------------
ADEV uint32_t xxhash32_seed(uint32_t px, uint32_t py, uint32_t pz) {
  const uint32_t P0 = 2246822519U, P1 = 3266489917U, P2 = 668265263U, P3 = 374761393U;
  uint32_t h = pz + P3 + px * P1;
  h = P2 * __funnelshift_l(h, h, 17);
  h += py * P1;
  h = P2 * __funnelshift_l(h, h, 17);
  h = P0 * (h ^ (h >> 15));
  h = P1 * (h ^ (h >> 13));
  return h ^ (h >> 16);
}
ADEV uint32_t pcg_next(uint32_t& state) {
  uint32_t prev = state * 747796405u + 2891336453u;
  uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  state = prev;
  return (word >> 22u) ^ word;
}
// pcg * (1.0/float(0xffffffffu)): the divisor rounds to 2^32, result may be exactly 1.0f
ADEV float rnd(uint32_t& seed) { return __uint2float_rn(pcg_next(seed)) * 2.3283064365386963e-10f; }
ADEV float2 rnd2(uint32_t& seed) {
  float a = rnd(seed);
  float b = rnd(seed);
  return make_float2(a, b);
}

// ---- small helpers: math.glsl:58-216 -------------------------------------------------------
ADEV float3 make_normal(float3 n) {
  float l = length(n);
  return l == 0.0f ? n : n / l;
}
ADEV float safe_sqrt(float v) { return sqrtf(fmaxf(0.0f, v)); }
ADEV float3 to_world(float3 X, float3 Y, float3 Z, float3 v) { return v.x * X + v.y * Y + v.z * Z; }
ADEV float3 to_local(float3 X, float3 Y, float3 Z, float3 v) { return f3(dot(v, X), dot(v, Y), dot(v, Z)); }
ADEV float3 uniform_sample_sphere(float2 u) {
  float z = 1.0f - 2.0f * u.x;
  float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
  float phi = kTwoPi * u.y;
  return f3(r * cosf(phi), r * sinf(phi), z);
}
ADEV float2 concentric_sample_disk(float2 u) {
  float ox = 2.0f * u.x - 1.0f, oy = 2.0f * u.y - 1.0f;
  if (ox == 0.0f && oy == 0.0f) return make_float2(0.0f, 0.0f);
  float theta, r;
  if (fabsf(ox) > fabsf(oy)) {
    r = ox;
    theta = kPiOver4 * (oy / ox);
  } else {
    r = oy;
    theta = kPiOver2 - kPiOver4 * (ox / oy);
  }
  return make_float2(r * cosf(theta), r * sinf(theta));
}
ADEV float3 cosine_sample_hemisphere(float2 u) {
  float2 d = concentric_sample_disk(u);
  return f3(d.x, d.y, sqrtf(fmaxf(0.0f, 1.0f - d.x * d.x - d.y * d.y)));
}
ADEV float cosine_hemisphere_pdf(float c) { return c <= 0.0f ? 0.0f : c * kInvPi; }
ADEV float power_heuristic(float a, float b) {
  a = a * a;
  b = b * b + a;
  return b == 0.0f ? 0.0f : a / b;
}
ADEV void basis(float3 n, float3& f, float3& r) {
  if (n.z < -0.999999f) {
    f = f3(0, -1, 0);
    r = f3(-1, 0, 0);
  } else {
    float a = 1.0f / (1.0f + n.z);
    float b = -n.x * n.y * a;
    f = f3(1.0f - n.x * n.x * a, b, -n.x);
    r = f3(b, 1.0f - n.y * n.y * a, -n.y);
  }
}
// math.glsl:241-266 (Waechter & Binder self-intersection offset)
ADEV float offset_component(float p, float n) {
  int of_i = (int)(256.0f * n);
  float p_i = __int_as_float(__float_as_int(p) + ((p < 0) ? -of_i : of_i));
  return fabsf(p) < (1.0f / 32.0f) ? p + (1.0f / 65536.0f) * n : p_i;
}
ADEV float3 offset_position_along_normal(float3 p, float3 n) {
  return f3(offset_component(p.x, n.x), offset_component(p.y, n.y), offset_component(p.z, n.z));
}
ADEV float luminance(float3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }  // sun_and_sky.glsl:29-31

// ---- textures: fp32 bilinear, REPEAT, LOD 0 (sampler of reference src/core/texture.cpp:99-107).
// Texels are float4 in linear memory, so each of the four taps is one 128-bit load.
ADEV int wrap_index(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}
// Not inlined: the textured shaders call it from up to nine places, and inlined copies (float and integer
// remainders) made up 28 % of the 10 k-instruction PBR kernel, which was stalling on instruction fetch.
ADEV int wrap_coord(float f0, int n) {  // f0 = floorf(coordinate): its remainder modulo n in [0, n)
  // |f0| < 2^30 converts to int exactly, so the float remainder of the general path is only needed beyond that
  return fabsf(f0) < 1073741824.0f ? wrap_index((int)f0, n) : wrap_index((int)fmodf(f0, (float)n), n);
}
__device__ __noinline__ float4 tex_bilinear(const DTexture& t, float u, float v) {
  float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
  if (!isfinite(x) || !isfinite(y)) return make_float4(0, 0, 0, 0);
  float fx0 = floorf(x), fy0 = floorf(y);
  float fx = x - fx0, fy = y - fy0;
  int x0 = wrap_coord(fx0, t.w), y0 = wrap_coord(fy0, t.h);
  int x1 = wrap_index(x0 + 1, t.w), y1 = wrap_index(y0 + 1, t.h);
  float4 p00 = __ldg(&t.texels[(size_t)y0 * t.w + x0]), p10 = __ldg(&t.texels[(size_t)y0 * t.w + x1]);
  float4 p01 = __ldg(&t.texels[(size_t)y1 * t.w + x0]), p11 = __ldg(&t.texels[(size_t)y1 * t.w + x1]);
  float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
  return make_float4(w00 * p00.x + w10 * p10.x + w01 * p01.x + w11 * p11.x,
                     w00 * p00.y + w10 * p10.y + w01 * p01.y + w11 * p11.y,
                     w00 * p00.z + w10 * p10.z + w01 * p01.z + w11 * p11.z,
                     w00 * p00.w + w10 * p10.w + w01 * p01.w + w11 * p11.w);
}

// ---- sun & sky: reference src/shaders/utils/sun_and_sky.glsl --------------------------------
constexpr float kSsPi = 3.1415926535f;  // :24
struct PerezCoeffs {
  float A, B, C, D, E;
};
// one evaluator for the Perez form the reference spells out at :174-177, :185-188 and :220-223
ADEV float perez_num(PerezCoeffs k, float cos_theta, float gamma, float cos_gamma) {
  return (1.0f + k.A * expf(k.B / cos_theta)) * (1.0f + k.C * expf(k.D * gamma) + k.E * cos_gamma * cos_gamma);
}
ADEV float perez_den(PerezCoeffs k, float theta_s, float cos_theta_s) {  // depends on the sun only
  return (1.0f + k.A * expf(k.B / 1.0f)) * (1.0f + k.C * expf(k.D * theta_s) + k.E * cos_theta_s * cos_theta_s);
}
ADEV float perez_ratio(PerezCoeffs k, float cos_theta, float gamma, float cos_gamma, float theta_s, float cos_theta_s) {
  return perez_num(k, cos_theta, gamma, cos_gamma) / perez_den(k, theta_s, cos_theta_s);
}
// :228-243 with :201-226 and :134-199 folded in
ADEV float3 sky_env_color(float3 sun, float3 dir, float T) {
  float theta_s = acosf(sun.z);
  float chi = (4.0f / 9.0f - T / 120.0f) * (kSsPi - 2.0f * theta_s);
  float lum = 1000.0f * ((4.0453f * T - 4.9710f) * tanf(chi) - 0.2155f * T + 2.4192f);
  float cg_raw = dot(sun, dir);
  {  // sky_luminance: cos_gamma clamped at 0 and mirrored above 1
    float cg = cg_raw < 0.0f ? 0.0f : cg_raw;
    if (cg > 1.0f) cg = 2.0f - cg;
    PerezCoeffs kY = {0.178721f * T - 1.463037f, -0.355402f * T + 0.427494f, -0.022669f * T + 5.325056f,
                      0.120647f * T - 2.577052f, -0.066967f * T + 0.370275f};
    lum *= perez_ratio(kY, dir.z, acosf(cg), cg, theta_s, sun.z);
  }
  // sky_color_xyz: cos_gamma only mirrored
  float cg = cg_raw > 1.0f ? 2.0f - cg_raw : cg_raw;
  float gamma = acosf(cg);
  float t2 = T * T, ts2 = theta_s * theta_s, ts3 = ts2 * theta_s;
  float zx = ((+0.001650f * ts3 - 0.003742f * ts2 + 0.002088f * theta_s + 0) * t2 +
              (-0.029028f * ts3 + 0.063773f * ts2 - 0.032020f * theta_s + 0.003948f) * T +
              (+0.116936f * ts3 - 0.211960f * ts2 + 0.060523f * theta_s + 0.258852f));
  float zy = ((+0.002759f * ts3 - 0.006105f * ts2 + 0.003162f * theta_s + 0) * t2 +
              (-0.042149f * ts3 + 0.089701f * ts2 - 0.041536f * theta_s + 0.005158f) * T +
              (+0.153467f * ts3 - 0.267568f * ts2 + 0.066698f * theta_s + 0.266881f));
  PerezCoeffs kx = {-0.019257f * T - (0.29f - powf(sun.z, 0.5f) * 0.09f), -0.066513f * T + 0.000818f,
                    -0.000417f * T + 0.212479f, -0.064097f * T - 0.898875f, -0.003251f * T + 0.045178f};
  PerezCoeffs ky = {-0.016698f * T - 0.260787f, -0.094958f * T + 0.009213f, -0.007928f * T + 0.210230f,
                    -0.044050f * T - 1.653694f, -0.010922f * T + 0.052919f};
  float x = zx * perez_ratio(kx, dir.z, gamma, cg, theta_s, sun.z);
  float y = zy * perez_ratio(ky, dir.z, gamma, cg, theta_s, sun.z);
  float Y = lum, X = (x / y) * Y, Z = ((1.0f - x - y) / y) * Y;
  return f3(3.241f * X - 1.537f * Y - 0.499f * Z, -0.969f * X + 1.876f * Y + 0.042f * Z,
            0.056f * X - 0.204f * Y + 1.057f * Z) * kSsPi;
}
// :108-132
ADEV float3 sun_disk_color(float3 sun, float T) {
  if (!(sun.z > 0.0f)) return f3(0.0f);
  const float3 ko = f3(12.0f, 8.5f, 0.9f), wl = f3(0.610f, 0.550f, 0.470f);
  const float3 sol = f3(1.0f * 127500 / 0.9878f, 0.992f * 127500 / 0.9878f, 0.911f * 127500 / 0.9878f);
  float m = 1.0f / (sun.z + 0.15f * powf(93.885f - acosf(sun.z) * 180 / kSsPi, -1.253f));
  float beta = 0.04608f * T - 0.04586f;
  float3 ta = exp3(-m * beta * pow3(wl, -1.3f));
  float3 to = exp3(-m * ko * 0.0035f);
  float3 tr = exp3(-m * 0.008735f * pow3(wl, -4.08f));
  return tr * ta * to * sol;
}
// :33-106 specialised for the only caller (normal = +z): stencil direction for sample (sx, sy)
ADEV float3 ground_stencil_dir(float sx, float sy) {
  float lx = 2 * sx - 1, ly = 2 * sy - 1, r = 0.0f, phi = 0.0f;
  if (!(lx == 0.0f && ly == 0.0f)) {
    if (lx > -ly) {
      if (lx > ly) { r = lx; phi = (kSsPi / 4.0f) * (1.0f + ly / lx); }
      else { r = ly; phi = (kSsPi / 4.0f) * (3.0f - lx / ly); }
    } else {
      if (lx < ly) { r = -lx; phi = (kSsPi / 4.0f) * (5.0f + ly / lx); }
      else { r = -ly; phi = (kSsPi / 4.0f) * (7.0f - lx / ly); }
    }
  }
  float x = r * cosf(phi), y = r * sinf(phi);
  float z2 = 1.0f - x * x - y * y;
  float z = z2 > 0.0f ? sqrtf(z2) : 0.0f;
  // xyz2dir with main = (0,0,1): u = (1,0,0), v = main x u = (0,1,0)
  return f3(x, y, z);
}
ADEV float3 tweak_vector(float3 d, int y_is_up, float horiz_height) {  // :278-288
  float3 o = (y_is_up == 1) ? f3(d.x, d.z, d.y) : d;
  if (horiz_height != 0) {
    o.z -= horiz_height;
    o = normalize(o);
  }
  return o;
}
// Sun vector as the model uses it (:436-447): tweaked, lifted to z >= 0.001; `factor` = night brightness adjustment.
ADEV float3 sun_vector(const AsunaSunSky& ss, float horiz, float& factor, float3& real_sun) {
  float3 sun = tweak_vector(normalize(f3(ss.sun_direction)), ss.y_is_up, horiz);
  real_sun = sun;
  if (sun.z < 0.001f) {
    if (sun.z < 0.0f) {  // night_brightness_adjustment :396-403
      const float lmt = 0.30901699437494742410229341718282f;
      if (sun.z <= -lmt) factor = 0.0f;
      else {
        float f = (sun.z + lmt) / lmt;
        f *= f;
        factor = f * f;
      }
    }
    sun.z = 0.001f;
    sun = normalize(sun);
  }
  return sun;
}
// calc_irrad :245-262 (float loop counters as written): mean sky colour over a 5x5 cosine stencil around +z.  It
// depends on the sun parameters only, not on the ray, so it is evaluated once per sun/sky setting
// (k_sky_ground_irradiance, integrator.cu) and handed to every below-horizon lookup through FrameParams.
ADEV float3 sky_ground_irradiance(const AsunaSunSky& ss) {
  float factor = 1.0f;
  float3 real_sun;
  float3 sun = sun_vector(ss, ss.horizon_height / 10.0f, factor, real_sun);
  float3 irrad = f3(0.0f);
  for (float u = 1.f / 10.f; u < 1.f; u += 1.f / 5.f)
    for (float v = 1.f / 10.f; v < 1.f; v += 1.f / 5.f) irrad += sky_env_color(sun, ground_stencil_dir(u, v), 2.0f);
  irrad /= 25.0f;
  return irrad;
}
// Everything of the model that depends on the AsunaSunSky setting only (:134-243 sky coefficients for T = haze,
// :108-132 sun colours, :264-276 saturation, :312-394 disk / glow scales, :245-262 ground irradiance): filled once per
// setting by k_sky_prepare (integrator.cu) with the same expressions sun_and_sky evaluates, then read by every lookup.
ADEV void sky_prepare(const AsunaSunSky& ss, SkyPre& p) {
  p.horiz = ss.horizon_height / 10.0f;
  p.haze = fmaxf(2.0f + ss.haze, 2.0f);
  {  // tweak_saturation :264-276
    float s = ss.saturation;
    if (s <= 1.0f) {
      float h = clampf((p.haze - 2.0f) / 15.0f, 0.0f, 1.0f);
      h = powf(h, 3.0f);
      p.sat = (s * (1.0f - h)) + powf(s, 3.0f) * h;
    } else
      p.sat = 1.0f;
  }
  float3 rgb_scale = f3(ss.rgb_unit_conversion);
  if (luminance(rgb_scale) < 0.0f) rgb_scale = f3(1.0f / 80000.0f);
  rgb_scale *= ss.multiplier;
  p.rgb_scale[0] = rgb_scale.x, p.rgb_scale[1] = rgb_scale.y, p.rgb_scale[2] = rgb_scale.z;
  float factor = 1.0f;
  float3 real_sun;
  const float3 sun = sun_vector(ss, p.horiz, factor, real_sun);
  p.factor = factor;
  p.sun[0] = sun.x, p.sun[1] = sun.y, p.sun[2] = sun.z;
  p.real_sun[0] = real_sun.x, p.real_sun[1] = real_sun.y, p.real_sun[2] = real_sun.z;
  const float3 up = sun_disk_color(sun, p.haze), down = sun_disk_color(sun, 2.0f);
  p.sun_color_up[0] = up.x, p.sun_color_up[1] = up.y, p.sun_color_up[2] = up.z;
  p.sun_color_down[0] = down.x, p.sun_color_down[1] = down.y, p.sun_color_down[2] = down.z;
  p.sun_radius = 0.00465f * ss.sun_disk_scale * 10.0f;
  p.disk_scale = 1.0f, p.glow_scale = 1.0f;
  if (ss.physically_scaled_sun == 1) {  // calc_physical_scale :312-394
    float disk_r = 0.00465f * ss.sun_disk_scale, glow_r = disk_r * 10.0f;
    float glow_integral = ss.sun_glow_intensity * ((4.f * kSsPi) - (24.f * kSsPi) / (glow_r * glow_r) +
                                                    (24.f * kSsPi) * sinf(glow_r) / (glow_r * glow_r * glow_r));
    float target = ss.sun_disk_intensity * kSsPi;
    float max_glow = 0.5f * target;
    if (glow_integral > max_glow) {
      p.glow_scale *= max_glow / glow_integral;
      target -= max_glow;
    } else
      target -= glow_integral;
    float area = 2 * kSsPi * (1 - cosf(disk_r));
    float target_intensity = target / area;
    float actual_intensity = ss.sun_disk_intensity * 100.0f * (1.0f * area) / area;
    p.disk_scale = (target_intensity == 0.0f) ? 0.0f : target_intensity / actual_intensity;
  }
  {  // sky_env_color's sun-only part for T = haze
    const float T = p.haze;
    const float theta_s = acosf(sun.z);
    const float chi = (4.0f / 9.0f - T / 120.0f) * (kSsPi - 2.0f * theta_s);
    p.lum0 = 1000.0f * ((4.0453f * T - 4.9710f) * tanf(chi) - 0.2155f * T + 2.4192f);
    const float t2 = T * T, ts2 = theta_s * theta_s, ts3 = ts2 * theta_s;
    p.zx = ((+0.001650f * ts3 - 0.003742f * ts2 + 0.002088f * theta_s + 0) * t2 +
            (-0.029028f * ts3 + 0.063773f * ts2 - 0.032020f * theta_s + 0.003948f) * T +
            (+0.116936f * ts3 - 0.211960f * ts2 + 0.060523f * theta_s + 0.258852f));
    p.zy = ((+0.002759f * ts3 - 0.006105f * ts2 + 0.003162f * theta_s + 0) * t2 +
            (-0.042149f * ts3 + 0.089701f * ts2 - 0.041536f * theta_s + 0.005158f) * T +
            (+0.153467f * ts3 - 0.267568f * ts2 + 0.066698f * theta_s + 0.266881f));
    const PerezCoeffs kY = {0.178721f * T - 1.463037f, -0.355402f * T + 0.427494f, -0.022669f * T + 5.325056f,
                            0.120647f * T - 2.577052f, -0.066967f * T + 0.370275f};
    const PerezCoeffs kx = {-0.019257f * T - (0.29f - powf(sun.z, 0.5f) * 0.09f), -0.066513f * T + 0.000818f,
                            -0.000417f * T + 0.212479f, -0.064097f * T - 0.898875f, -0.003251f * T + 0.045178f};
    const PerezCoeffs ky = {-0.016698f * T - 0.260787f, -0.094958f * T + 0.009213f, -0.007928f * T + 0.210230f,
                            -0.044050f * T - 1.653694f, -0.010922f * T + 0.052919f};
    const PerezCoeffs* src[3] = {&kY, &kx, &ky};
    for (int c = 0; c < 3; c++) {
      p.perez[c][0] = src[c]->A, p.perez[c][1] = src[c]->B, p.perez[c][2] = src[c]->C, p.perez[c][3] = src[c]->D,
      p.perez[c][4] = src[c]->E;
      p.perez_den[c] = perez_den(*src[c], theta_s, sun.z);
    }
  }
  const float3 irrad = sky_ground_irradiance(ss);
  p.ground_irrad[0] = irrad.x, p.ground_irrad[1] = irrad.y, p.ground_irrad[2] = irrad.z;
}
// sky_env_color(sun, dir, haze) from the prepared coefficients: only the direction-dependent factors remain
ADEV float3 sky_env_color_pre(const SkyPre& p, float3 sun, float3 dir) {
  const PerezCoeffs kY = {p.perez[0][0], p.perez[0][1], p.perez[0][2], p.perez[0][3], p.perez[0][4]};
  const PerezCoeffs kx = {p.perez[1][0], p.perez[1][1], p.perez[1][2], p.perez[1][3], p.perez[1][4]};
  const PerezCoeffs ky = {p.perez[2][0], p.perez[2][1], p.perez[2][2], p.perez[2][3], p.perez[2][4]};
  float lum = p.lum0;
  const float cg_raw = dot(sun, dir);
  {
    float cg = cg_raw < 0.0f ? 0.0f : cg_raw;
    if (cg > 1.0f) cg = 2.0f - cg;
    lum *= perez_num(kY, dir.z, acosf(cg), cg) / p.perez_den[0];
  }
  const float cg = cg_raw > 1.0f ? 2.0f - cg_raw : cg_raw;
  const float gamma = acosf(cg);
  const float x = p.zx * (perez_num(kx, dir.z, gamma, cg) / p.perez_den[1]);
  const float y = p.zy * (perez_num(ky, dir.z, gamma, cg) / p.perez_den[2]);
  const float Y = lum, X = (x / y) * Y, Z = ((1.0f - x - y) / y) * Y;
  return f3(3.241f * X - 1.537f * Y - 0.499f * Z, -0.969f * X + 1.876f * Y + 0.042f * Z,
            0.056f * X - 0.204f * Y + 1.057f * Z) * kSsPi;
}
// :405-533 with the per-setting values taken from `p`
__device__ __noinline__ float3 sun_and_sky(const AsunaSunSky& ss, const SkyPre& p, float3 in_dir) {
  if (ss.multiplier <= 0.0f) return f3(0.0f);
  float night_factor = 1.0f;
  const float factor = p.factor, haze = p.haze, sat = p.sat;
  const float3 rgb_scale = f3(p.rgb_scale), sun = f3(p.sun), real_sun = f3(p.real_sun);
  float3 dir = tweak_vector(in_dir, ss.y_is_up, p.horiz);
  const float downness = dir.z;
  const float3 real_dir = dir;
  if (dir.z < 0.001f) {
    dir.z = 0.001f;
    dir = normalize(dir);
  }
  float3 tint = f3(0.0f);
  if (factor > 0.0f) {
    tint = sky_env_color_pre(p, sun, dir);
    if (factor < 1.0f) tint *= factor;
  }
  const float3 sun_color = downness > 0 ? f3(p.sun_color_up) : f3(p.sun_color_down);
  if (ss.sun_disk_intensity > 0.0f && ss.sun_disk_scale > 0.0f) {
    float sun_angle = acosf(dot(real_dir, real_sun));
    if (sun_angle < p.sun_radius) {
      float f = (1.0f - sun_angle / p.sun_radius) * 10.0f;
      f = powf(f / 10.0f, 3.0f) * 2.0f * ss.sun_glow_intensity * p.glow_scale +
          smoothstepf(8.5f, 9.5f + (haze / 50.0f), f) * 100.0f * ss.sun_disk_intensity * p.disk_scale;
      tint += sun_color * f;
    }
  }
  float3 out = tint * rgb_scale;
  if (downness <= 0.0f) {
    float3 down = f3(ss.ground_color) * ((f3(p.ground_irrad) + sun_color * sun.z) * rgb_scale);
    if (factor < 1) down *= factor;
    float blur = ss.horizon_blur / 10.0f;
    if (blur > 0.0f) {
      float d = fminf(-downness / blur, 1.0f);
      d = smoothstepf(0.0f, 1.0f, d);
      out = out * (1.0f - d) + down * d;
      night_factor = 1.0f - d;
    } else {
      out = down;
      night_factor = 0.0f;
    }
  }
  // arch_colortweak :290-310
  float inten = luminance(out);
  float3 res = (sat <= 0.0f) ? f3(inten) : out * sat + f3(inten * (1.0f - sat));
  res = res * f3(1.0f + ss.redblueshift, 1.0f, 1.0f - ss.redblueshift);
  if (night_factor > 0.0f) {
    float3 night = f3(ss.night_color) * night_factor;
    res = f3(fmaxf(res.x, night.x), fmaxf(res.y, night.y), fmaxf(res.z, night.z));
  }
  return res * kSsPi;
}

// ---- environment map: reference src/shaders/utils/sample_light.glsl:84-129 ------------------
struct EnvCtx {
  const DTexture* env;  // [3]
  const float* env_transform;
  float res_x, res_y, intensity;
};
ADEV float2 env_dir_to_uv(float3 L, float& theta) {
  theta = acosf(clampf(L.y, -1.0f, 1.0f));
  return make_float2((kPi + atan2f(L.z, L.x)) * kInv2Pi, theta * kInvPi);
}
ADEV float env_pdf(const EnvCtx& e, float3 L) {
  L = make_normal(mat4_vector_transposed(e.env_transform, L));
  float theta;
  float2 uv = env_dir_to_uv(L, theta);
  float pdf = tex_bilinear(e.env[2], uv.x, uv.y).y * tex_bilinear(e.env[1], 0.f, uv.y).y;
  float st = sinf(theta);
  if (st == 0) return 0;
  return (pdf * e.res_x * e.res_y) / (kTwoPi * kPi * st);
}
ADEV float3 env_eval(const EnvCtx& e, float3 L) {
  L = make_normal(mat4_vector_transposed(e.env_transform, L));
  float theta;
  float2 uv = env_dir_to_uv(L, theta);
  return e.intensity * f3(tex_bilinear(e.env[0], uv.x, uv.y));
}
ADEV float3 env_sample(const EnvCtx& e, float2 r, float3& L, float& pdf) {
  float v = tex_bilinear(e.env[1], 0.f, r.x).x;  // marginal (fetched at u = 0: A.3-15)
  float u = tex_bilinear(e.env[2], r.y, v).x;    // conditional
  pdf = tex_bilinear(e.env[2], u, v).y * tex_bilinear(e.env[1], 0.f, v).y;
  float phi = u * kTwoPi, theta = v * kPi;
  float st = sinf(theta), ct = cosf(theta);
  if (st == 0.0f) pdf = 0.0f;
  pdf = (pdf * e.res_x * e.res_y) / (kTwoPi * kPi * st);
  L = f3(-st * cosf(phi), ct, -st * sinf(phi));
  L = make_normal(mat4_vector(e.env_transform, L));
  return e.intensity * f3(tex_bilinear(e.env[0], u, v));
}

// ---- per-thread path / surface records -------------------------------------------------------
struct Surface {  // HitState of rchit_layouts.glsl:37-58 minus the material copy
  float2 uv;
  float3 pos, V, N, geoN, ffN, X, Y;
};
struct LightSample {  // LightSamplingRecord, structs.glsl:52-63
  float3 d, n;
  float dist, pdf;
  uint32_t flags;
};
struct PathRegs {
  float3 ray_o, ray_d, throughput, radiance;
  uint32_t seed;
  int depth;
  uint32_t bsdf_flags;
  float bsdf_pdf;
  bool stop;
  // next-event estimation result (DirectLightRecord)
  bool nee;
  float3 nee_o, nee_d, nee_L;
  float nee_dist;
};
struct ShadeEnv {  // read-only inputs of one shade call
  const SceneView* scene;
  const FrameParams* fp;
  EnvCtx env;
  const OutputImages* out;  // kernel parameter (constant bank): channel cid -> image cid + 1, written on frame 0 only
  bool frame0;
};

ADEV void configure_frame(const AsunaState& pc, Surface& s) {  // rchit_layouts.glsl:61-65
  if (pc.useFaceNormal == 1) s.N = s.geoN;
  basis(s.N, s.X, s.Y);
  s.ffN = dot(s.N, s.V) > 0 ? s.N : -s.N;
}

// sample_light.glsl:10-82
ADEV float3 sample_one_light(float2 r, const AsunaLight& light, float3 pos, LightSample& ls) {
  float3 lu = f3(light.u), lv = f3(light.v), lp = f3(light.position);
  if (light.type == ASUNA_LIGHT_RECT || light.type == ASUNA_LIGHT_TRIANGLE) {
    float r1 = r.x;
    float r2 = light.type == ASUNA_LIGHT_TRIANGLE ? (1 - r1) * r.y : r.y;  // A.3-2
    ls.d = lp + lu * r1 + lv * r2 - pos;
    ls.dist = length(ls.d);
    float dist_sq = ls.dist * ls.dist;
    ls.d /= ls.dist;
    ls.n = make_normal(cross(lu, lv));
    ls.pdf = dist_sq / (light.area * fabsf(dot(ls.n, ls.d)) + kEps);
    ls.flags = kLightArea;
    return f3(light.radiance);
  } else if (light.type == ASUNA_LIGHT_DIRECTIONAL) {
    ls.d = make_normal(f3(light.direction));
    ls.n = -ls.d;
    ls.dist = kInfinity;
    ls.pdf = 1.0f;
    ls.flags = kLightDelta;
    return f3(light.radiance);
  } else if (light.type == ASUNA_LIGHT_POINT) {
    ls.d = lp - pos;
    ls.n = -ls.d;
    ls.dist = length(ls.d);
    float dist_sq = ls.dist * ls.dist;
    ls.d /= ls.dist + kEps;  // A.3-12
    ls.pdf = 1.0f;
    ls.flags = kLightDelta;
    return f3(light.radiance) / (dist_sq + kEps);
  }
  return f3(0.0f);
}

// rchit_layouts.glsl:98-167.  Fills the NEE ray of `p` and returns the light radiance estimate.
ADEV float3 sample_lights(const ShadeEnv& se, PathRegs& p, float3 pos, float3 normal, bool& visible, LightSample& ls) {
  const AsunaState& pc = se.fp->pc;
  const AsunaSunSky& sk = se.fp->sunsky;
  bool allow_double = false;
  float3 radiance = f3(0.0f);
  ls.d = ls.n = f3(0.0f);
  ls.dist = ls.pdf = 0.0f;
  ls.flags = 0;  // A.3-7
  bool has_env = (pc.hasEnvMap == 1 || sk.in_use == 1);
  bool has_light = (pc.numLights > 0);
  float env_sel = has_env ? (has_light ? 0.5f : 1.0f) : 0.0f;
  float ana_sel = has_light ? (has_env ? 0.5f : 1.0f) : 0.0f;
  float sel = rnd(p.seed);
  if (sel < env_sel) {
    ls.flags = kLightArea;
    ls.dist = kInfinity;
    ls.n = -make_normal(p.ray_d);
    float2 u = rnd2(p.seed);
    if (sk.in_use == 1) {
      ls.d = uniform_sample_sphere(u);
      ls.pdf = kInv4Pi;
      radiance = sun_and_sky(sk, se.fp->sky, ls.d);
    } else if (pc.hasEnvMap == 1) {
      radiance = env_sample(se.env, u, ls.d, ls.pdf);
    } else {
      ls.d = uniform_sample_sphere(u);
      ls.pdf = kInv4Pi;
      radiance = f3(pc.bgColor);
    }
    radiance = radiance / env_sel;
    allow_double = true;
  } else if (sel < env_sel + ana_sel) {
    int li = min(1 + (int)(rnd(p.seed) * pc.numLights), pc.numLights);
    const AsunaLight light = se.scene->lights[li];
    float2 r = rnd2(p.seed);
    radiance = sample_one_light(r, light, pos, ls) * (float)pc.numLights / ana_sel;
    allow_double = (light.doubleSide == 1);
  }
  p.nee_o = offset_position_along_normal(pos, normal);
  p.nee_d = ls.d;
  p.nee_dist = ls.dist;
  visible = (dot(ls.d, normal) > 0.0f && ls.pdf > 0.0f);
  visible = visible && (dot(ls.n, ls.d) < 0 || allow_double);
  return radiance;
}

// For the delta BSDFs (dielectric, conductor) eval(..., EArea) is identically zero (A.3-4), so the direct-light
// record of sampleLights never contributes: only its random-number consumption (A.1) has to be reproduced --
// not the env-map / sun-sky / light evaluation behind it.
ADEV void sample_lights_rng_only(const ShadeEnv& se, PathRegs& p) {
  const AsunaState& pc = se.fp->pc;
  bool has_env = (pc.hasEnvMap == 1 || se.fp->sunsky.in_use == 1);
  bool has_light = (pc.numLights > 0);
  float env_sel = has_env ? (has_light ? 0.5f : 1.0f) : 0.0f;
  float ana_sel = has_light ? (has_env ? 0.5f : 1.0f) : 0.0f;
  float sel = rnd(p.seed);
  if (sel < env_sel) {
    (void)rnd2(p.seed);
  } else if (sel < env_sel + ana_sel) {
    (void)rnd(p.seed);
    (void)rnd2(p.seed);
  }
  p.nee = false;
  p.nee_L = f3(0.0f);
}

ADEV void store_direct(PathRegs& p, bool visible, float3 w, float bsdf_pdf, float3 radiance, const LightSample& ls) {
  float3 Ld = f3(0.0f);
  if (visible) Ld = power_heuristic(ls.pdf, bsdf_pdf) * w * radiance * p.throughput / (ls.pdf + kEps);
  p.nee_L = Ld;
  p.nee = visible;
}
// tail of every closest-hit main() that divides by (pdf + EPS)
ADEV void next_ray(PathRegs& p, const Surface& s, float3 d, float pdf, uint32_t flags, float3 w, float3 offset_n) {
  if (pdf <= 0.0f || length(w) == 0.0f) {
    p.stop = true;
    return;
  }
  p.bsdf_pdf = pdf;
  p.bsdf_flags = flags;
  p.ray_o = offset_position_along_normal(s.pos, offset_n);
  p.ray_d = d;
  p.throughput *= w / (pdf + kEps);
}

ADEV float4 tex(const ShadeEnv& se, int id, float2 uv) { return tex_bilinear(se.scene->textures[id], uv.x, uv.y); }
ADEV void apply_normal_map(const ShadeEnv& se, const AsunaMaterial& m, Surface& s) {
  if (m.normalTextureId >= 0) {
    float3 n = 2.0f * f3(tex(se, m.normalTextureId, s.uv)) - 1.0f;
    s.N = make_normal(to_world(s.X, s.Y, s.N, n));
    configure_frame(se.fp->pc, s);
  }
}
ADEV float3 diffuse_of(const ShadeEnv& se, const AsunaMaterial& m, const Surface& s) {
  return m.diffuseTextureId >= 0 ? f3(tex(se, m.diffuseTextureId, s.uv)) : f3(m.diffuse);
}
ADEV void write_aov(const ShadeEnv& se, uint32_t pixel, int ch, float3 v) {
  if (se.frame0 && ch >= 0 && ch < ASUNA_NUM_OUTPUT_IMAGES - 1) se.out->img[ch + 1][pixel] = make_float4(v.x, v.y, v.z, 1.0f);
}

// ---- microfacet pieces shared by pbr / rough_plastic / kang18 (identical text in the three shaders)
ADEV float sqr(float x) { return x * x; }
ADEV float ggx_d(float HdotN, float HdotX, float HdotY, float ax, float ay) {
  return 1 / (kPi * ax * ay * sqr(sqr(HdotX / ax) + sqr(HdotY / ay) + sqr(HdotN)) + kEps);
}
ADEV float3 ggx_sample(float2 u, float3 wo, float ax, float ay) {
  float factor = safe_sqrt(u.x / fmaxf(1 - u.x, kEps));
  float phi = kTwoPi * u.y;
  float3 wh = make_normal(f3(-ax * factor * cosf(phi), -ay * factor * sinf(phi), 1.0f));
  return reflect(-wo, wh);
}
ADEV float ggx_pdf(float3 wh, float3 wo, float ax, float ay) {
  float HdotV = dot(wh, wo);
  float3 wi = reflect(-wo, wh);
  if (wi.z > 0.0f && wo.z > 0.0f && wh.z > 0.0f) return ggx_d(wh.z, wh.x, wh.y, ax, ay) * fabsf(wh.z) / (4 * HdotV + kEps);
  return 0.0f;
}
ADEV float ggx_g1(float NdotV, float VdotX, float VdotY, float ax, float ay) {
  if (NdotV <= 0.0f) return 0.0f;
  return 1 / (NdotV + length(f3(ax * VdotX, ay * VdotY, NdotV)));
}
// brdf_plastic.rchit:12-41 == brdf_rough_plastic.rchit:12-41 (cosThetaT output unused by callers)
ADEV float fresnel_dielectric_ext(float cos_i_, float eta) {
  if (eta == 1) return 0.0f;
  float scale = (cos_i_ > 0) ? 1 / eta : eta, cos_t2 = 1 - (1 - cos_i_ * cos_i_) * (scale * scale);
  if (cos_t2 <= 0.0f) return 1.0f;
  float ci = fabsf(cos_i_), ct = sqrtf(cos_t2);
  float Rs = (ci - eta * ct) / (ci + eta * ct);
  float Rp = (eta * ci - ct) / (eta * ci + ct);
  return 0.5f * (Rs * Rs + Rp * Rp);
}

// ---- emitter hit: brdf_lambertian.rchit:44-68 -------------------------------------------------
ADEV void shade_light_hit(const ShadeEnv& se, PathRegs& p, int light_id, float3 hit_pos) {
  const AsunaLight light = se.scene->lights[light_id];
  float3 dir = make_normal(p.ray_d);
  float3 ln = make_normal(cross(f3(light.u), f3(light.v)));
  float side = dot(ln, dir);
  p.stop = true;
  if (side > 0 && light.doubleSide == 0) return;
  float mis = 1.0f;
  if ((p.bsdf_flags & kSmooth) != 0 && p.depth != 1) {
    float dist = length(hit_pos - p.ray_o);
    float light_pdf = dist * dist / (light.area * fabsf(side) + kEps);
    mis = power_heuristic(p.bsdf_pdf, light_pdf);
  }
  p.radiance += p.throughput * f3(light.radiance) * mis;
}

// ---- brdf_lambertian.rchit:70-151 -------------------------------------------------------------
ADEV void shade_lambertian(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m, uint32_t pixel) {
  const AsunaState& pc = se.fp->pc;
  float3 kd = diffuse_of(se, m, s);
  apply_normal_map(se, m, s);
  if (p.depth == 1) {
    write_aov(se, pixel, pc.diffuseOutChannel, kd);
    write_aov(se, pixel, pc.normalOutChannel, s.N);
    write_aov(se, pixel, pc.specularOutChannel, f3(0.0f));
    write_aov(se, pixel, pc.tangentOutChannel, s.X);
    write_aov(se, pixel, pc.roughnessOutChannel, f3(1, 1, 0));
    write_aov(se, pixel, pc.positionOutChannel, s.pos);
    write_aov(se, pixel, pc.uvOutChannel, f3(s.uv.x, s.uv.y, 1));
  }
  {
    bool visible;
    LightSample ls;
    float3 radiance = sample_lights(se, p, s.pos, s.ffN, visible, ls);
    float3 w = f3(0.0f);
    float bpdf = 0.0f;
    if (visible) {
      float NdotL = dot(s.ffN, ls.d), NdotV = dot(s.ffN, s.V);
      if (!(NdotL < 0 || NdotV < 0)) w = kd * kInvPi * NdotL;
      if (ls.flags & kLightArea) bpdf = cosine_hemisphere_pdf(NdotL);
    }
    store_direct(p, visible, w, bpdf, radiance, ls);
  }
  float3 wi = cosine_sample_hemisphere(rnd2(p.seed));
  float pdf = cosine_hemisphere_pdf(wi.z);
  float3 w = kd * kInvPi * fabsf(wi.z);
  if (pdf <= 0.0f || length(w) == 0.0f) {
    p.stop = true;
    return;
  }
  p.bsdf_pdf = pdf;
  p.bsdf_flags = kDiffuseReflection;
  p.ray_o = offset_position_along_normal(s.pos, s.ffN);
  p.ray_d = to_world(s.X, s.Y, s.ffN, wi);
  p.throughput *= w / pdf;  // lambertian divides by pdf, not pdf + EPS
}

// ---- brdf_emissive.rchit:12-26 ----------------------------------------------------------------
ADEV void shade_emissive(const ShadeEnv& se, PathRegs& p, const Surface& s, const AsunaMaterial& m) {
  p.stop = true;
  float3 rad = f3(m.radiance);
  if (m.radianceTextureId >= 0) rad = f3(m.radianceFactor) * f3(tex(se, m.radianceTextureId, s.uv));
  if (se.fp->pc.ignoreEmissive == 0) p.radiance += rad * p.throughput;
}

// ---- bsdf_dielectric.rchit --------------------------------------------------------------------
ADEV float dielectric_fresnel(float cos_i, float eta) {  // :12-24
  float sin_t2 = eta * eta * (1.0f - cos_i * cos_i);
  if (sin_t2 > 1.0f) return 1.0f;
  float cos_t = sqrtf(fmaxf(1.0f - sin_t2, 0.0f));
  float rs = (eta * cos_t - cos_i) / (eta * cos_t + cos_i);
  float rp = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
  return 0.5f * (rs * rs + rp * rp);
}
ADEV void shade_dielectric(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m) {  // :74-138
  apply_normal_map(se, m, s);
  float eta = dot(s.V, s.N) > 0.0f ? (1.0f / m.ior) : m.ior;
  float F = dielectric_fresnel(fabsf(dot(s.V, s.ffN)), eta);
  sample_lights_rng_only(se, p);  // :96-118: eval(...,EArea) = 0, the shadow ray it fires adds nothing
  float u = rnd(p.seed);
  float3 d, w;
  float pdf;
  uint32_t flags;
  if (u < F) {
    d = make_normal(reflect(-s.V, s.ffN));
    pdf = F;
    flags = kSpecularReflection;
    w = f3(F);
  } else {
    d = make_normal(refract(-s.V, s.ffN, eta));
    pdf = 1 - F;
    flags = kSpecularTransmission;
    w = f3((1 - F) * eta * eta);
  }
  next_ray(p, s, d, pdf, flags, w, signf(dot(d, s.N)) * s.N);
}

// ---- brdf_conductor.rchit ---------------------------------------------------------------------
ADEV float conductor_reflectance(float eta, float k, float ci) {  // :14-32 (returns 0.5(Rs + Rs Rp), A.3-10)
  float ci2 = ci * ci;
  float si2 = fmaxf(1.0f - ci2, 0.0f);
  float si4 = si2 * si2;
  float inner = eta * eta - k * k - si2;
  float a2b2 = sqrtf(fmaxf(inner * inner + 4.0f * eta * eta * k * k, 0.0f));
  float a = sqrtf(fmaxf((a2b2 + inner) * 0.5f, 0.0f));
  float Rs = ((a2b2 + ci2) - (2.0f * a * ci)) / ((a2b2 + ci2) + (2.0f * a * ci));
  float Rp = ((ci2 * a2b2 + si4) - (2.0f * a * ci * si2)) / ((ci2 * a2b2 + si4) + (2.0f * a * ci * si2));
  return 0.5f * (Rs + Rs * Rp);
}
ADEV void shade_conductor(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m) {  // :89-155
  float3 kd = diffuse_of(se, m, s);
  apply_normal_map(se, m, s);
  float3 eta = f3(m.radiance), k = f3(m.radianceFactor);
  sample_lights_rng_only(se, p);  // :107-133: eval(...,EArea) = 0
  (void)rnd2(p.seed);  // sampleBsdf takes a vec2 it never uses (:68)
  float NdotV = dot(s.V, s.ffN);
  if (NdotV <= 0) {
    p.stop = true;
    return;
  }
  float3 d = reflect(-s.V, s.ffN);
  float c = dot(s.ffN, d);
  float3 w = kd * f3(conductor_reflectance(eta.x, k.x, c), conductor_reflectance(eta.y, k.y, c),
                     conductor_reflectance(eta.z, k.z, c));
  next_ray(p, s, d, 1.0f, kSpecularReflection, w, s.ffN);
}

// ---- brdf_plastic.rchit:133-204 ---------------------------------------------------------------
ADEV void shade_plastic(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m) {
  float3 kd = diffuse_of(se, m, s);
  apply_normal_map(se, m, s);
  float eta = m.ior, fdr = m.radiance[0];
  float d_avg = luminance(kd), s_avg = luminance(f3(1.0f));
  float spec_w = s_avg / (d_avg + s_avg);
  float inv_eta2 = 1 / (eta * eta);
  const float3 N = s.ffN, V = s.V;
  float NdotV = dot(V, N);
  float Fo = fresnel_dielectric_ext(NdotV, eta);
  float prob_spec = (Fo * spec_w) / (Fo * spec_w + (1 - Fo) * (1 - spec_w));
  float3 diff = kd;
  diff /= (1.0f - diff * fdr);
  {
    bool visible;
    LightSample ls;
    float3 radiance = sample_lights(se, p, s.pos, s.ffN, visible, ls);
    float3 w = f3(0.0f);
    float bpdf = 0.0f;
    if (visible) {
      float NdotL = dot(ls.d, N);
      if (!(NdotL < 0 || NdotV < 0)) {
        float Fi = fresnel_dielectric_ext(NdotL, eta);
        w = (1 - Fi) * (1 - Fo) * diff * inv_eta2 * kInvPi * NdotL;  // eval(:43-66) with EArea
        if (ls.flags & kLightDelta) {                                // pdf(:68-92) with lRec.flags
          if (fabsf(dot(reflect(-V, N), ls.d) - 1) < kEps) bpdf = prob_spec;
        } else if (ls.flags & kLightArea)
          bpdf = NdotL * (1 - prob_spec);
      }
    }
    store_direct(p, visible, w, bpdf, radiance, ls);
  }
  float2 u = rnd2(p.seed);
  if (NdotV <= 0) {
    p.stop = true;
    return;
  }
  if (u.x < prob_spec) {
    next_ray(p, s, make_normal(reflect(-V, N)), prob_spec, kSpecularReflection, f3(Fo), s.ffN);
  } else {
    u.x = (u.x - prob_spec) / (1 + kEps - prob_spec);
    float3 wi = cosine_sample_hemisphere(u);
    float Fi = fresnel_dielectric_ext(wi.z, eta);
    float pdf = (1 - prob_spec) * cosine_hemisphere_pdf(wi.z);
    float3 w = (1 - Fi) * (1 - Fo) * inv_eta2 * diff * kInvPi * fabsf(wi.z);
    next_ray(p, s, to_world(s.X, s.Y, N, wi), pdf, kDiffuseReflection, w, s.ffN);
  }
}

// ---- brdf_rough_plastic.rchit -----------------------------------------------------------------
struct RoughPlasticArgs {
  float3 kd, ks;
  float eta, fdr, inv_eta2, ax, ay;
};
ADEV float3 rough_plastic_eval(float3 L, const Surface& s, RoughPlasticArgs a) {  // :82-111 (flags = EArea)
  const float3 N = s.ffN, V = s.V;
  float NdotL = dot(L, N), NdotV = dot(V, N);
  if (NdotL < 0 || NdotV < 0) return f3(0.0f);
  float3 H = make_normal(L + V);
  float Fs = fresnel_dielectric_ext(dot(H, V), a.eta);
  float Gs = ggx_g1(NdotV, dot(V, s.X), dot(V, s.Y), a.ax, a.ay) * ggx_g1(NdotL, dot(L, s.X), dot(L, s.Y), a.ax, a.ay);
  float Ds = ggx_d(dot(H, N), dot(H, s.X), dot(H, s.Y), a.ax, a.ay);
  float3 w = a.ks * Fs * Gs * Ds * NdotL;
  float Fo = fresnel_dielectric_ext(NdotV, a.eta), Fi = fresnel_dielectric_ext(NdotL, a.eta);
  float3 diff = a.kd;
  diff /= (1.0f - diff * a.fdr);
  return w + (1 - Fi) * (1 - Fo) * diff * a.inv_eta2 * kInvPi * NdotL;
}
ADEV float rough_plastic_pdf(float3 L, const Surface& s, float ax, float ay, float eta, float substrate_w,
                             uint32_t flags) {  // :113-136
  const float3 N = s.ffN, V = s.V;
  float NdotL = dot(L, N), NdotV = dot(V, N);
  if (NdotL < 0 || NdotV < 0 || (flags & kLightArea) == 0) return 0.0f;
  float Fo = fresnel_dielectric_ext(NdotV, eta);
  float prob_spec = Fo / (Fo + substrate_w * (1.0f - Fo));
  float3 wi = to_local(s.X, s.Y, N, L), wo = to_local(s.X, s.Y, N, V);
  float3 wh = make_normal(wi + wo);
  return prob_spec * ggx_pdf(wh, wo, ax, ay) + (1 - prob_spec) * cosine_hemisphere_pdf(wi.z);
}
ADEV void shade_rough_plastic(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m) {  // :191-269
  RoughPlasticArgs a;
  a.kd = diffuse_of(se, m, s);
  float2 alpha = make_float2(m.anisoAlpha[0], m.anisoAlpha[1]);
  if (m.roughnessTextureId >= 0) {
    float4 c = tex(se, m.roughnessTextureId, s.uv);
    alpha = make_float2(c.x, c.y);
  }
  apply_normal_map(se, m, s);
  a.ks = f3(1.0f);
  a.eta = m.ior;
  a.fdr = m.radiance[0];
  a.inv_eta2 = 1 / (a.eta * a.eta);
  a.ax = fmaxf(kEps, alpha.x);
  a.ay = fmaxf(kEps, alpha.y);
  float substrate_w = luminance(a.kd);
  const float3 N = s.ffN, V = s.V;
  {
    bool visible;
    LightSample ls;
    float3 radiance = sample_lights(se, p, s.pos, s.ffN, visible, ls);
    float3 w = f3(0.0f);
    float bpdf = 0.0f;
    if (visible) {
      // A.3-1: the NEE call sites (:234-240) pass the arguments in a different order than the
      // signatures: eval receives (invEta2<-ax, ax<-ay, ay<-invEta2), pdf receives
      // (ax<-eta, ay<-ax, eta<-ay).  Reproduced.
      RoughPlasticArgs b = a;
      b.inv_eta2 = a.ax;
      b.ax = a.ay;
      b.ay = a.inv_eta2;
      w = rough_plastic_eval(ls.d, s, b);
      bpdf = rough_plastic_pdf(ls.d, s, /*ax=*/a.eta, /*ay=*/a.ax, /*eta=*/a.ay, substrate_w, ls.flags);
    }
    store_direct(p, visible, w, bpdf, radiance, ls);
  }
  float2 u = rnd2(p.seed);
  float NdotV = dot(V, N);
  if (NdotV <= 0) {
    p.stop = true;
    return;
  }
  float Fo = fresnel_dielectric_ext(NdotV, a.eta);
  float prob_spec = Fo / (Fo + substrate_w * (1.0f - Fo));
  if (u.x < prob_spec) {
    u.x = u.x / prob_spec;
    float3 wo = to_local(s.X, s.Y, N, V);
    float3 wi = ggx_sample(u, wo, a.ax, a.ay);
    float3 L = to_world(s.X, s.Y, N, wi);
    float3 H = make_normal(V + L);
    float3 wh = make_normal(wi + wo);
    float NdotL = dot(N, L);
    float Fs = fresnel_dielectric_ext(dot(H, V), a.eta);
    float Gs = ggx_g1(NdotV, dot(V, s.X), dot(V, s.Y), a.ax, a.ay) * ggx_g1(NdotL, dot(L, s.X), dot(L, s.Y), a.ax, a.ay);
    float Ds = ggx_d(dot(H, N), dot(H, s.X), dot(H, s.Y), a.ax, a.ay);
    next_ray(p, s, L, ggx_pdf(wh, wo, a.ax, a.ay) * prob_spec, kGlossyReflection, a.ks * Fs * Gs * Ds * NdotL, s.ffN);
  } else {
    u.x = (u.x - prob_spec) / (1 + kEps - prob_spec);
    float3 wi = cosine_sample_hemisphere(u);
    float Fi = fresnel_dielectric_ext(wi.z, a.eta);
    float3 diff = a.kd;
    diff /= (1.0f - diff * a.fdr);
    float pdf = (1 - prob_spec) * cosine_hemisphere_pdf(wi.z);
    float3 w = (1 - Fi) * (1 - Fo) * a.inv_eta2 * diff * kInvPi * fabsf(wi.z);
    next_ray(p, s, to_world(s.X, s.Y, N, wi), pdf, kDiffuseReflection, w, s.ffN);
  }
}

// ---- GGX + Lambert two-lobe BRDFs: pbr_metalness_roughness and kang18 ---------------------------
// eval: pbr :57-86 (Schlick on mix(F0, albedo, metalness), diffuse scaled by 1-metalness),
//       kang18 :61-90 (scalar Schlick times rhoSpec).  `spec_tint`/`f0` carry the difference.
ADEV float3 two_lobe_eval(float3 L, const Surface& s, float3 diffuse_term, float3 f0, float3 spec_scale, float ax, float ay) {
  const float3 N = s.ffN, V = s.V;
  float NdotL = dot(N, L), NdotV = dot(N, V);
  if (NdotL <= 0.0f || NdotV <= 0.0f) return f3(0.0f);
  float3 H = make_normal(L + V);
  float HdotV = dot(H, V);
  float3 Fs = f0 + (1.0f - f0) * powf(clampf(1 - HdotV, 0, 1), 5.0f);
  float Ds = ggx_d(dot(H, N), dot(H, s.X), dot(H, s.Y), ax, ay);
  float Gs = ggx_g1(NdotV, dot(V, s.X), dot(V, s.Y), ax, ay) * ggx_g1(NdotL, dot(L, s.X), dot(L, s.Y), ax, ay);
  return (diffuse_term + spec_scale * Fs * Ds * Gs) * NdotL;
}
ADEV float two_lobe_pdf(float3 L, const Surface& s, float ax, float ay, float p_diffuse, uint32_t flags) {  // pbr :88-105
  if ((flags & kLightArea) == 0) return 0.0f;
  const float3 N = s.ffN, V = s.V;
  float NdotL = dot(N, L), NdotV = dot(N, V);
  if (NdotL <= 0.0f || NdotV <= 0.0f) return 0.0f;
  float3 H = make_normal(L + V);
  float3 wh = make_normal(to_local(s.X, s.Y, N, H));
  float3 wo = make_normal(to_local(s.X, s.Y, N, V));
  float3 wi = make_normal(to_local(s.X, s.Y, N, L));
  return p_diffuse * cosine_hemisphere_pdf(wi.z) + (1 - p_diffuse) * ggx_pdf(wh, wo, ax, ay);
}
// opacity pass-through, pbr :151-160 / kang18 :176-180: continue the same ray from behind the surface
ADEV bool pass_through(PathRegs& p, const Surface& s, float opacity) {
  if (rnd(p.seed) < opacity) {
    p.ray_o = offset_position_along_normal(s.pos, -s.ffN);
    p.depth--;
    return true;
  }
  return false;
}
ADEV void two_lobe_sample(const ShadeEnv& se, PathRegs& p, const Surface& s, float3 diffuse_term, float3 f0,
                          float3 spec_scale, float ax, float ay, float p_diffuse, bool times_cos) {
  float2 u = rnd2(p.seed);  // argument evaluated before the lobe-select rand of the body (A.1)
  const float3 N = s.ffN;
  float3 wo = make_normal(to_local(s.X, s.Y, N, s.V));
  float3 wi;
  float pdf;
  uint32_t flags;
  if (rnd(p.seed) < p_diffuse) {
    wi = cosine_sample_hemisphere(u);
    flags = kDiffuseReflection;
    pdf = cosine_hemisphere_pdf(wi.z);
  } else {
    wi = ggx_sample(u, wo, ax, ay);
    pdf = ggx_pdf(make_normal(wi + wo), wo, ax, ay);
    flags = kGlossyReflection;
  }
  float3 d = to_world(s.X, s.Y, N, wi);
  float3 w = two_lobe_eval(d, s, diffuse_term, f0, spec_scale, ax, ay);
  if (times_cos) w = w * fabsf(wi.z);  // pbr multiplies eval (which already holds NdotL) by |wi.z| again (:125-127)
  next_ray(p, s, d, pdf, flags, w, s.ffN);
}

ADEV void shade_pbr(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m, uint32_t pixel) {  // :131-219
  const AsunaState& pc = se.fp->pc;
  float3 albedo = diffuse_of(se, m, s);
  float metalness = m.metalnessTextureId >= 0 ? tex(se, m.metalnessTextureId, s.uv).x : m.metalness;
  float roughness = m.roughnessTextureId >= 0 ? tex(se, m.roughnessTextureId, s.uv).x : m.roughness;
  apply_normal_map(se, m, s);
  float opacity = m.opacityTextureId >= 0 ? tex(se, m.opacityTextureId, s.uv).x : m.specular;
  if (pass_through(p, s, opacity)) return;
  float ax = fmaxf(sqr(roughness), 0.001f), ay = ax;
  float eta = m.ior;
  if (p.depth == 1) {
    write_aov(se, pixel, pc.diffuseOutChannel, albedo);
    write_aov(se, pixel, pc.normalOutChannel, s.ffN);
  }
  float F0 = sqr((eta - 1) / (eta + 1));
  float3 f0 = mix3(f3(F0), albedo, metalness);
  float3 diffuse_term = albedo * (1 - metalness) * kInvPi;
  const float kDiffuseLobe = 0.2f;
  {
    bool visible;
    LightSample ls;
    float3 radiance = sample_lights(se, p, s.pos, s.ffN, visible, ls);
    float3 w = f3(0.0f);
    float bpdf = 0.0f;
    if (visible) {
      w = two_lobe_eval(ls.d, s, diffuse_term, f0, f3(1.0f), ax, ay);
      bpdf = two_lobe_pdf(ls.d, s, ax, ay, kDiffuseLobe, ls.flags);
    }
    store_direct(p, visible, w, bpdf, radiance, ls);
  }
  two_lobe_sample(se, p, s, diffuse_term, f0, f3(1.0f), ax, ay, kDiffuseLobe, true);
}

ADEV void shade_kang18(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m, const DInstance& in,
                       uint32_t pixel) {  // :141-249
  const AsunaState& pc = se.fp->pc;
  float3 kd = diffuse_of(se, m, s);
  float3 ks = m.metalnessTextureId >= 0 ? f3(tex(se, m.metalnessTextureId, s.uv)) : f3(m.rhoSpec);
  float2 alpha = make_float2(m.anisoAlpha[0], m.anisoAlpha[1]);
  if (m.roughnessTextureId >= 0) {
    float4 c = tex(se, m.roughnessTextureId, s.uv);
    alpha = make_float2(c.x, c.y);
  }
  float opacity = m.opacityTextureId >= 0 ? tex(se, m.opacityTextureId, s.uv).x : m.metalness;
  if (m.normalTextureId >= 0) {  // object-space normal map (:157-163)
    float3 n = 2.0f * f3(tex(se, m.normalTextureId, s.uv)) - 1.0f;
    s.N = make_normal(xf_normal(in.w2o, n));
    configure_frame(pc, s);
  }
  if (m.tangentTextureId >= 0) {  // (:165-169)
    float3 t = 2.0f * f3(tex(se, m.tangentTextureId, s.uv)) - 1.0f;
    s.X = make_normal(xf_normal(in.w2o, t));
  }
  s.Y = make_normal(cross(s.N, s.X));
  s.X = make_normal(cross(s.Y, s.N));
  s.ffN = dot(s.N, s.V) > 0 ? s.N : -s.N;
  if (pass_through(p, s, opacity)) return;
  float ax = fmaxf(alpha.x, kEps), ay = fmaxf(alpha.y, kEps);
  float eta = m.ior;
  if (p.depth == 1) {
    write_aov(se, pixel, pc.diffuseOutChannel, kd);
    write_aov(se, pixel, pc.normalOutChannel, s.ffN);
    write_aov(se, pixel, pc.specularOutChannel, ks);
    write_aov(se, pixel, pc.tangentOutChannel, s.X);
    write_aov(se, pixel, pc.roughnessOutChannel, f3(ax, ay, 0));
    write_aov(se, pixel, pc.positionOutChannel, s.pos);
    write_aov(se, pixel, pc.uvOutChannel, f3(s.uv.x, s.uv.y, 1));
  }
  float F0 = sqr((eta - 1) / (eta + 1));
  auto lum709 = [](float3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; };  // :57-59
  float pd = lum709(kd), ps = lum709(ks);
  float p_diffuse = pd / (pd + ps + kEps);
  float3 diffuse_term = kd * kInvPi;
  {
    bool visible;
    LightSample ls;
    float3 radiance = sample_lights(se, p, s.pos, s.ffN, visible, ls);
    float3 w = f3(0.0f);
    float bpdf = 0.0f;
    if (visible) {
      w = two_lobe_eval(ls.d, s, diffuse_term, f3(F0), ks, ax, ay);
      bpdf = two_lobe_pdf(ls.d, s, ax, ay, p_diffuse, ls.flags);
    }
    store_direct(p, visible, w, bpdf, radiance, ls);
  }
  two_lobe_sample(se, p, s, diffuse_term, f3(F0), ks, ax, ay, p_diffuse, false);
}

// ---- brdf_mirror.rchit:39-101 ------------------------------------------------------------------
ADEV void shade_mirror(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m) {
  float3 kd = diffuse_of(se, m, s);
  apply_normal_map(se, m, s);
  sample_lights_rng_only(se, p);  // eval(:12-19) is called with EArea and tests EDelta: identically 0
  float3 d = reflect(-s.V, s.ffN);  // sampleBsdf(:28-37): no random numbers, pdf 1
  if (length(kd) == 0.0f) {
    p.stop = true;
    return;
  }
  p.bsdf_pdf = 1.0f;
  p.bsdf_flags = kSpecularReflection;
  p.ray_o = offset_position_along_normal(s.pos, s.ffN);
  p.ray_d = d;
  p.throughput *= kd;  // divides by pdf (= 1), not pdf + EPS
}

// ---- brdf_rough_conductor.rchit -----------------------------------------------------------------
ADEV float3 conductor_reflectance3(float3 eta, float3 k, float c) {
  return f3(conductor_reflectance(eta.x, k.x, c), conductor_reflectance(eta.y, k.y, c), conductor_reflectance(eta.z, k.z, c));
}
ADEV float3 rough_conductor_term(float3 L, const Surface& s, float3 kd, float3 eta, float3 k, float ax, float ay, float NdotL,
                                 float NdotV) {  // :86-95 == :136-146
  const float3 N = s.ffN, V = s.V;
  float3 H = make_normal(L + V);
  float3 Fs = conductor_reflectance3(eta, k, NdotL);
  float Gs = ggx_g1(NdotV, dot(V, s.X), dot(V, s.Y), ax, ay);
  Gs *= ggx_g1(NdotL, dot(L, s.X), dot(L, s.Y), ax, ay);
  float Ds = ggx_d(dot(H, N), dot(H, s.X), dot(H, s.Y), ax, ay);
  return kd * Fs * Gs * Ds * NdotL;
}
ADEV void shade_rough_conductor(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m) {  // :151-220
  float3 kd = diffuse_of(se, m, s);
  float2 alpha = make_float2(m.anisoAlpha[0], m.anisoAlpha[1]);
  if (m.roughnessTextureId >= 0) {
    float4 c = tex(se, m.roughnessTextureId, s.uv);
    alpha = make_float2(c.x, c.y);
  }
  apply_normal_map(se, m, s);
  float3 eta = f3(m.radiance), k = f3(m.radianceFactor);
  float ax = fmaxf(alpha.x, kEps), ay = fmaxf(alpha.y, kEps);
  const float3 N = s.ffN, V = s.V;
  {
    bool visible;
    LightSample ls;
    float3 radiance = sample_lights(se, p, s.pos, s.ffN, visible, ls);
    float3 w = f3(0.0f);
    float bpdf = 0.0f;
    if (visible) {
      float NdotL = dot(ls.d, N), NdotV = dot(V, N);
      if (!(NdotL < 0 || NdotV < 0)) {
        w = rough_conductor_term(ls.d, s, kd, eta, k, ax, ay, NdotL, NdotV);
        if (ls.flags & kLightArea) {
          float3 wi = to_local(s.X, s.Y, N, ls.d), wo = to_local(s.X, s.Y, N, V);
          bpdf = ggx_pdf(make_normal(wi + wo), wo, ax, ay);
        }
      }
    }
    store_direct(p, visible, w, bpdf, radiance, ls);
  }
  float2 u = rnd2(p.seed);
  float NdotV = dot(V, N);
  if (NdotV <= 0) {
    p.stop = true;
    return;
  }
  float3 wo = to_local(s.X, s.Y, N, V);
  float3 wi = ggx_sample(u, wo, ax, ay);
  float3 L = to_world(s.X, s.Y, N, wi);
  float3 wh = make_normal(wi + wo);
  float NdotL = dot(N, L);
  float3 w = rough_conductor_term(L, s, kd, eta, k, ax, ay, NdotL, NdotV);
  next_ray(p, s, L, ggx_pdf(wh, wo, ax, ay), kGlossyReflection, w, s.ffN);
}

// ---- brdf_phong.rchit -----------------------------------------------------------------------------
ADEV float3 glsl_normalize(float3 v) { return v / length(v); }  // GLSL normalize: no zero-length guard
ADEV float3 phong_eval(float3 L, float3 V, float3 N, float3 diffuse, float3 specular, float shininess, float dw) {  // :12-31, EArea
  float NdotL = dot(N, L), NdotV = dot(N, V);
  if (NdotL < 0 || NdotV < 0) return f3(0.0f);
  float3 H = glsl_normalize(L + V);
  float3 diffuse_lobe = diffuse * kInvPi;
  float3 specular_lobe = specular * powf(fmaxf(dot(H, N), 0.0f), shininess) * (shininess + 2) * kInv2Pi;
  return (dw * diffuse_lobe + (1 - dw) * specular_lobe) * NdotL;
}
ADEV float phong_pdf(float3 L, float3 V, float3 N, float shininess) {  // :33-38
  float3 H = glsl_normalize(L + V);
  float NdotH = fmaxf(dot(H, N), 0.0f), VdotH = fmaxf(dot(H, V), 0.0f);
  return (shininess + 1) * kInv2Pi * powf(NdotH, shininess) * 0.25f / (VdotH + kEps);
}
ADEV void shade_phong(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& m, uint32_t pixel) {  // :113-193
  const AsunaState& pc = se.fp->pc;
  float3 diffuse = diffuse_of(se, m, s);
  apply_normal_map(se, m, s);
  float3 specular = f3(m.rhoSpec);
  float shininess = m.specular;
  if (p.depth == 1) {
    write_aov(se, pixel, pc.diffuseOutChannel, diffuse);
    write_aov(se, pixel, pc.normalOutChannel, s.ffN);
    write_aov(se, pixel, pc.specularOutChannel, specular);
    write_aov(se, pixel, pc.tangentOutChannel, s.X);
    write_aov(se, pixel, pc.roughnessOutChannel, f3(1, 1, 0));
    write_aov(se, pixel, pc.positionOutChannel, s.pos);
    write_aov(se, pixel, pc.uvOutChannel, f3(s.uv.x, s.uv.y, 1));
  }
  const float3 N = s.ffN, V = s.V;
  float db = luminance(diffuse), sb = luminance(specular);
  float dw = db / (db + sb), sw = 1 - dw;
  {
    bool visible;
    LightSample ls;
    float3 radiance = sample_lights(se, p, s.pos, s.ffN, visible, ls);
    float3 w = f3(0.0f);
    float bpdf = 0.0f;
    if (visible) {
      w = phong_eval(ls.d, V, N, diffuse, specular, shininess, dw);
      if (ls.flags & kLightArea) bpdf = dw * cosine_hemisphere_pdf(dot(N, ls.d)) + sw * phong_pdf(ls.d, V, N, shininess);
    }
    store_direct(p, visible, w, bpdf, radiance, ls);
  }
  float2 u = rnd2(p.seed);
  float3 d;
  float pdf;
  uint32_t flags;
  if (u.x < dw) {
    u.x /= dw;
    float3 wi = cosine_sample_hemisphere(u);
    pdf = cosine_hemisphere_pdf(wi.z);
    d = to_world(s.X, s.Y, N, wi);
    flags = kDiffuseReflection;
  } else {
    u.x = (u.x - dw) / sw;
    float ct = powf(u.x, 1 / (shininess + 1));
    float phi = kTwoPi * u.y;
    float st = safe_sqrt(1 - ct * ct);
    float3 H = to_world(s.X, s.Y, N, f3(st * sinf(phi), st * cosf(phi), ct));
    d = reflect(-V, H);
    pdf = phong_pdf(d, V, N, shininess);
    flags = kGlossyReflection;
  }
  float3 w = phong_eval(d, V, N, diffuse, specular, shininess, dw);
  if (pdf <= 0.0f || length(w) == 0.0f) {
    p.stop = true;
    return;
  }
  p.bsdf_pdf = pdf;
  p.bsdf_flags = flags;
  p.ray_o = offset_position_along_normal(s.pos, s.ffN);
  p.ray_d = d;
  p.throughput *= w / pdf;
}

// ---- brdf_disney.rchit ------------------------------------------------------------------------------
struct DisneyMat {  // :31-49 (the fields the shader reads)
  float3 base;
  float metallic, roughness, subsurface, specular_tint, sheen, sheen_tint, clearcoat, clearcoat_roughness, ax, ay;
};
ADEV float lum709(float3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }  // :51-53
ADEV float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
ADEV float gtr1(float NdotH, float a) {  // :55-60
  if (a >= 1.0f) return kInvPi;
  float a2 = a * a;
  float t = 1.0f + (a2 - 1.0f) * NdotH * NdotH;
  return (a2 - 1.0f) / (kPi * logf(a2) * t);
}
ADEV float3 sample_gtr1(float rgh, float r1) {  // :62-74
  float a = fmaxf(0.001f, rgh);
  float a2 = a * a;
  float phi = r1 * kTwoPi;
  float ct = sqrtf((1.0f - powf(a2, 1.0f - r1)) / (1.0f - a2));
  float st = clampf(sqrtf(1.0f - (ct * ct)), 0.0f, 1.0f);
  return f3(st * cosf(phi), st * sinf(phi), ct);
}
ADEV float3 sample_ggx_vndf(float3 V, float ax, float ay, float r1, float r2) {  // :96-114
  float3 Vh = glsl_normalize(f3(ax * V.x, ay * V.y, V.z));
  float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
  float3 T1 = lensq > 0 ? f3(-Vh.y, Vh.x, 0) * (1.0f / sqrtf(lensq)) : f3(1, 0, 0);
  float3 T2 = cross(Vh, T1);
  float r = sqrtf(r1);
  float phi = 2.0f * kPi * r2;
  float t1 = r * cosf(phi), t2 = r * sinf(phi);
  float sv = 0.5f * (1.0f + Vh.z);
  t2 = (1.0f - sv) * sqrtf(1.0f - t1 * t1) + sv * t2;
  float3 Nh = t1 * T1 + t2 * T2 + sqrtf(fmaxf(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
  return glsl_normalize(f3(ax * Nh.x, ay * Nh.y, fmaxf(0.0f, Nh.z)));
}
ADEV float gtr2_aniso(float NdotH, float HdotX, float HdotY, float ax, float ay) {  // :116-121
  float a = HdotX / ax, b = HdotY / ay;
  float c = a * a + b * b + NdotH * NdotH;
  return 1.0f / (kPi * ax * ay * c * c);
}
ADEV float smith_g(float NdotV, float alpha_g) {  // :134-138
  float a = alpha_g * alpha_g, b = NdotV * NdotV;
  return (2.0f * NdotV) / (NdotV + sqrtf(a + b - a * b));
}
ADEV float smith_g_aniso(float NdotV, float VdotX, float VdotY, float ax, float ay) {  // :140-145
  float a = VdotX * ax, b = VdotY * ay, c = NdotV;
  return (2.0f * NdotV) / (NdotV + sqrtf(a * a + b * b + c * c));
}
ADEV float schlick_fresnel(float u) {  // :147-151
  float m = clampf(1.0f - u, 0.0f, 1.0f);
  float m2 = m * m;
  return m2 * m2 * m;
}
ADEV float disney_fresnel(float metallic, float eta, float LdotH, float VdotH) {  // :167-171
  return mixf(dielectric_fresnel(fabsf(VdotH), eta), schlick_fresnel(LdotH), metallic);
}
ADEV float3 disney_diffuse(const DisneyMat& m, float3 Csheen, float3 V, float3 L, float3 H, float& pdf) {  // :173-197
  pdf = 0.0f;
  if (L.z <= 0.0f) return f3(0.0f);
  float FL = schlick_fresnel(L.z), FV = schlick_fresnel(V.z), FH = schlick_fresnel(dot(L, H));
  float Fd90 = 0.5f + 2.0f * dot(L, H) * dot(L, H) * m.roughness;
  float Fd = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
  float Fss90 = dot(L, H) * dot(L, H) * m.roughness;
  float Fss = mixf(1.0f, Fss90, FL) * mixf(1.0f, Fss90, FV);
  float ss = 1.25f * (Fss * (1.0f / (L.z + V.z) - 0.5f) + 0.5f);
  float3 Fsheen = FH * m.sheen * Csheen;
  pdf = L.z * kInvPi;
  return (1.0f - m.metallic) * (kInvPi * mixf(Fd, ss, m.subsurface) * m.base + Fsheen);
}
ADEV float3 disney_spec(const DisneyMat& m, float eta, float3 spec_col, float3 V, float3 L, float3 H, float& pdf) {  // :199-212
  pdf = 0.0f;
  if (L.z <= 0.0f) return f3(0.0f);
  float FM = disney_fresnel(m.metallic, eta, dot(L, H), dot(V, H));
  float3 F = mix3(spec_col, f3(1.0f), FM);
  float D = gtr2_aniso(H.z, H.x, H.y, m.ax, m.ay);
  float G1 = smith_g_aniso(fabsf(V.z), V.x, V.y, m.ax, m.ay);
  float G2 = G1 * smith_g_aniso(fabsf(L.z), L.x, L.y, m.ax, m.ay);
  pdf = G1 * D / (4.0f * V.z);
  return F * D * G2 / (4.0f * L.z * V.z);
}
ADEV float3 disney_clearcoat(const DisneyMat& m, float3 V, float3 L, float3 H, float& pdf) {  // :214-226
  pdf = 0.0f;
  if (L.z <= 0.0f) return f3(0.0f);
  float FH = dielectric_fresnel(dot(V, H), 1.0f / 1.5f);
  float F = mixf(0.04f, 1.0f, FH);
  float D = gtr1(H.z, m.clearcoat_roughness);
  float G = smith_g(L.z, 0.25f) * smith_g(V.z, 0.25f);
  float jacobian = 1.0f / (4.0f * dot(V, H));
  pdf = D * H.z * jacobian;
  return f3(0.25f) * m.clearcoat * F * D * G / (4.0f * L.z * V.z);
}
ADEV void disney_spec_color(const DisneyMat& m, float eta, float3& spec_col, float3& sheen_col) {  // :228-236
  float lum = lum709(m.base);
  float3 ctint = lum > 0.0f ? m.base / lum : f3(1.0f);
  float F0 = (1.0f - eta) / (1.0f + eta);
  spec_col = mix3(F0 * F0 * mix3(f3(1.0f), ctint, m.specular_tint), m.base, m.metallic);
  sheen_col = mix3(f3(1.0f), ctint, m.sheen_tint);
}
ADEV void disney_lobes(const DisneyMat& m, float3 spec_col, float approx_fresnel, float& wd, float& ws, float& wc) {  // :238-250
  wd = lum709(m.base) * (1.0f - m.metallic);
  ws = lum709(mix3(spec_col, f3(1.0f), approx_fresnel));
  wc = 0.25f * m.clearcoat * (1.0f - m.metallic);
  float total = wd + ws + wc;
  wd /= total, ws /= total, wc /= total;
}
ADEV float3 disney_eval(float3 Lw, const Surface& s, const DisneyMat& m, float eta, float& bsdf_pdf) {  // :252-300, EArea
  float3 weight = f3(0.0f);
  bsdf_pdf = 0.0f;
  float3 V = to_local(s.X, s.Y, s.ffN, s.V), L = to_local(s.X, s.Y, s.ffN, Lw);
  if (L.z <= 0 || V.z <= 0) return weight;
  float3 H = glsl_normalize(L + V);
  if (H.z < 0.0f) H = -H;
  float3 spec_col, sheen_col;
  disney_spec_color(m, eta, spec_col, sheen_col);
  float wd, ws, wc;
  disney_lobes(m, spec_col, disney_fresnel(m.metallic, eta, dot(L, H), dot(V, H)), wd, ws, wc);
  float pdf = 0.0f;
  if (wd > 0.0f) {
    weight += disney_diffuse(m, sheen_col, V, L, H, pdf);
    bsdf_pdf += pdf * wd;
  }
  if (ws > 0.0f) {
    weight += disney_spec(m, eta, spec_col, V, L, H, pdf);
    bsdf_pdf += pdf * ws;
  }
  if (wc > 0.0f) {
    weight += disney_clearcoat(m, V, L, H, pdf);
    bsdf_pdf += pdf * wc;
  }
  return weight * L.z;
}
ADEV void shade_disney(const ShadeEnv& se, PathRegs& p, Surface& s, const AsunaMaterial& mt) {  // :373-487
  float3 base = diffuse_of(se, mt, s);
  float metalness = mt.metalnessTextureId >= 0 ? tex(se, mt.metalnessTextureId, s.uv).x : mt.metalness;
  float roughness = mt.roughnessTextureId >= 0 ? tex(se, mt.roughnessTextureId, s.uv).x : mt.roughness;
  apply_normal_map(se, mt, s);
  float opacity = mt.opacityTextureId >= 0 ? tex(se, mt.opacityTextureId, s.uv).x : mt.rhoSpec[0];
  if (pass_through(p, s, opacity)) return;
  DisneyMat m;
  float aspect = sqrtf(1.0f - mt.anisotropic * 0.9f);
  m.ax = fmaxf(0.001f, roughness * roughness / aspect);
  m.ay = fmaxf(0.001f, roughness * roughness * aspect);
  m.base = base;
  m.metallic = metalness;
  m.roughness = fmaxf(roughness * roughness, 0.001f);
  m.subsurface = mt.subsurface, m.specular_tint = mt.specularTint, m.sheen = mt.sheen, m.sheen_tint = mt.sheenTint;
  m.clearcoat = mt.clearcoat;
  m.clearcoat_roughness = mixf(0.1f, 0.001f, mt.clearcoatGloss);
  float eta = dot(s.V, s.N) > 0.0f ? (1.0f / mt.ior) : mt.ior;
  {
    bool visible;
    LightSample ls;
    float3 radiance = sample_lights(se, p, s.pos, s.ffN, visible, ls);
    float3 w = f3(0.0f);
    float bpdf = 0.0f;
    if (visible) w = disney_eval(ls.d, s, m, eta, bpdf);
    store_direct(p, visible, w, bpdf, radiance, ls);
  }
  float2 u = rnd2(p.seed);  // sampleBsdf(:302-371)
  float r1 = u.x, r2 = u.y, pdf = 0.0f;
  float3 f = f3(0.0f), L;
  uint32_t flags;
  const float3 N = s.ffN;
  float3 V = to_local(s.X, s.Y, N, s.V);
  float3 spec_col, sheen_col;
  disney_spec_color(m, eta, spec_col, sheen_col);
  float wd, ws, wc;
  disney_lobes(m, spec_col, disney_fresnel(m.metallic, eta, V.z, V.z), wd, ws, wc);
  float cdf0 = wd, cdf1 = cdf0 + wc;
  if (r1 < cdf0) {
    r1 /= cdf0;
    L = cosine_sample_hemisphere(make_float2(r1, r2));
    f = disney_diffuse(m, sheen_col, V, L, glsl_normalize(L + V), pdf);
    pdf *= wd;
    flags = kDiffuseReflection;
  } else if (r1 < cdf1) {
    r1 = (r1 - cdf0) / (cdf1 - cdf0);
    float3 H = sample_gtr1(m.clearcoat_roughness, r1);
    if (H.z < 0.0f) H = -H;
    L = glsl_normalize(reflect(-V, H));
    f = disney_clearcoat(m, V, L, H, pdf);
    pdf *= wc;
    flags = kGlossyReflection;
  } else {
    r1 = (r1 - cdf1) / (1.0f - cdf1);
    float3 H = sample_ggx_vndf(V, m.ax, m.ay, r1, r2);
    if (H.z < 0.0f) H = -H;
    L = glsl_normalize(reflect(-V, H));
    f = disney_spec(m, eta, spec_col, V, L, H, pdf);
    pdf *= ws;
    flags = kGlossyReflection;
  }
  float3 d = to_world(s.X, s.Y, N, L);
  float3 w = f * fabsf(dot(N, L));  // :370 mixes the world-space normal with the local direction, as written
  if (pdf <= 0.0f || length(w) == 0.0f) {
    p.stop = true;
    return;
  }
  p.bsdf_pdf = pdf;
  p.bsdf_flags = flags;
  p.ray_o = offset_position_along_normal(s.pos, s.ffN);
  p.ray_d = d;
  p.throughput *= w / pdf;
}

// ---- miss: raytrace.default.rmiss:24-55 -------------------------------------------------------
ADEV void shade_miss(const ShadeEnv& se, PathRegs& p) {
  const AsunaState& pc = se.fp->pc;
  const AsunaSunSky& sk = se.fp->sunsky;
  p.stop = true;
  float3 d = p.ray_d, env;
  if (sk.in_use == 1) env = sun_and_sky(sk, se.fp->sky, d);
  else if (pc.hasEnvMap == 1) env = env_eval(se.env, d);
  else env = f3(pc.bgColor);
  float mis = 1.0f;
  if (p.depth != 1 && (p.bsdf_flags & kSmooth) != 0) {
    float env_pdf_v = (sk.in_use != 1 && pc.hasEnvMap == 1) ? env_pdf(se.env, d) : kInv4Pi;
    mis = power_heuristic(p.bsdf_pdf, env_pdf_v);
  }
  p.radiance += p.throughput * env * mis;
}

}  // namespace asuna
