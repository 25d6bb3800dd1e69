// TEST INFRASTRUCTURE ONLY.  Prologue of the generated translation unit oracle/_ref/ref_glsl_gen.cpp.
//
// The generator (oracle/refbuild/build_ref.py) reads the reference's own GLSL from /root/reference at
// build time, rewrites only what C++ cannot parse (parameter qualifiers, unsuffixed float literals,
// swizzles, layout declarations) and pastes it after this file, inside namespace refglsl.  Vector and
// matrix types and the GLSL built-ins come from the GLM 0.9.9 that is vendored in the reference tree
// (src/ext/nvpro_core/third_party/tinygltf/examples/common/glm); this prologue adds what GLM lacks:
//   * mixed int/float overloads (GLSL converts `2 * v`, `max(0, x)`, `pow(x, 2)` implicitly);
//   * rvalue swizzles as `v->*SW_xyz` (the generator rewrites `.xyz`);
//   * sampler2D / image2D over plain fp32 RGBA memory, texture() = bilinear, REPEAT, LOD 0, fp32
//     weights (the sampler the reference creates: core/texture.cpp:103-107, 244-248);
//   * the ray-tracing built-in variables as thread-local globals.
// No reference source is stored in this repository: the generated file lives in oracle/_ref/ (git-ignored).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define GLM_FORCE_RADIANS
#include <glm/glm.hpp>
#include <glm/gtc/matrix_access.hpp>

namespace refglsl {
using namespace glm;
using uint = unsigned int;

// ---------------------------------------------------------------- swizzles (rvalue only)
struct SWT_xyz {}; struct SWT_xy {}; struct SWT_rgb {}; struct SWT_rgba {}; struct SWT_rg {};
static const SWT_xyz SW_xyz{}; static const SWT_xy SW_xy{}; static const SWT_rgb SW_rgb{};
static const SWT_rgba SW_rgba{}; static const SWT_rg SW_rg{};
template <class T, precision P> inline vec<3, T, P> operator->*(vec<4, T, P> const& v, SWT_xyz) { return vec<3, T, P>(v.x, v.y, v.z); }
template <class T, precision P> inline vec<3, T, P> operator->*(vec<3, T, P> const& v, SWT_xyz) { return v; }
template <class T, precision P> inline vec<3, T, P> operator->*(vec<4, T, P> const& v, SWT_rgb) { return vec<3, T, P>(v.x, v.y, v.z); }
template <class T, precision P> inline vec<3, T, P> operator->*(vec<3, T, P> const& v, SWT_rgb) { return v; }
template <class T, precision P> inline vec<4, T, P> operator->*(vec<4, T, P> const& v, SWT_rgba) { return v; }
template <class T, precision P> inline vec<2, T, P> operator->*(vec<4, T, P> const& v, SWT_xy) { return vec<2, T, P>(v.x, v.y); }
template <class T, precision P> inline vec<2, T, P> operator->*(vec<3, T, P> const& v, SWT_xy) { return vec<2, T, P>(v.x, v.y); }
template <class T, precision P> inline vec<2, T, P> operator->*(vec<4, T, P> const& v, SWT_rg) { return vec<2, T, P>(v.x, v.y); }
template <class T, precision P> inline vec<2, T, P> operator->*(vec<3, T, P> const& v, SWT_rg) { return vec<2, T, P>(v.x, v.y); }

// ---------------------------------------------------------------- implicit int -> float conversions
#define REFGLSL_MIXED_OPS(V)                                               \
  inline V operator*(int a, V const& v) { return float(a) * v; }           \
  inline V operator*(V const& v, int a) { return v * float(a); }           \
  inline V operator/(int a, V const& v) { return float(a) / v; }           \
  inline V operator/(V const& v, int a) { return v / float(a); }           \
  inline V operator+(int a, V const& v) { return float(a) + v; }           \
  inline V operator+(V const& v, int a) { return v + float(a); }           \
  inline V operator-(int a, V const& v) { return float(a) - v; }           \
  inline V operator-(V const& v, int a) { return v - float(a); }           \
  inline V operator*(uint a, V const& v) { return float(a) * v; }          \
  inline V operator*(V const& v, uint a) { return v * float(a); }          \
  inline V operator/(V const& v, uint a) { return v / float(a); }
REFGLSL_MIXED_OPS(vec2)
REFGLSL_MIXED_OPS(vec3)
REFGLSL_MIXED_OPS(vec4)

using glm::max; using glm::min; using glm::clamp; using glm::pow; using glm::mix; using glm::abs;
using glm::sqrt; using glm::exp; using glm::log; using glm::sin; using glm::cos; using glm::tan;
using glm::acos; using glm::asin; using glm::atan; using glm::floor;
using glm::smoothstep; using glm::dot; using glm::length; using glm::normalize; using glm::cross;
using glm::reflect; using glm::refract; using glm::isnan; using glm::isinf; using glm::transpose;
using glm::inverse; using glm::floatBitsToInt; using glm::intBitsToFloat; using glm::exp2; using glm::log2;
using glm::fract; using glm::mod; using glm::radians; using glm::degrees; using glm::distance;

inline float max(int a, float b) { return glm::max(float(a), b); }
inline float max(float a, int b) { return glm::max(a, float(b)); }
inline float min(int a, float b) { return glm::min(float(a), b); }
inline float min(float a, int b) { return glm::min(a, float(b)); }
inline float clamp(float x, int a, int b) { return glm::clamp(x, float(a), float(b)); }
inline vec3 clamp(vec3 const& x, int a, int b) { return glm::clamp(x, float(a), float(b)); }
inline float pow(float a, int b) { return glm::pow(a, float(b)); }
inline float pow(int a, float b) { return glm::pow(float(a), b); }
inline float mix(float a, float b, int t) { return glm::mix(a, b, float(t)); }
inline float mix(int a, float b, float t) { return glm::mix(float(a), b, t); }
inline float mix(float a, int b, float t) { return glm::mix(a, float(b), t); }
inline float mix(int a, int b, float t) { return glm::mix(float(a), float(b), t); }
// glm 0.9.9.0's scalar sign() does not compile (lessThan on scalars); GLSL: 1, 0 or -1
inline float sign(float x) { return float((0.0f < x) - (x < 0.0f)); }
inline float abs(int a) { return float(a < 0 ? -a : a); }
inline float sqrt(int a) { return glm::sqrt(float(a)); }
inline float step(float e, float x) { return x < e ? 0.0f : 1.0f; }  // glm 0.9.9.0 scalar step() has the same defect
inline float smoothstep(int a, int b, float x) { return glm::smoothstep(float(a), float(b), x); }

// ---------------------------------------------------------------- resources
struct sampler2D {
  const float* rgba = nullptr;  // w*h RGBA32F, row-major
  int w = 0, h = 0;
};
struct image2D {
  float* rgba = nullptr;
  int w = 0, h = 0;
};
struct accelerationStructureEXT {};

// Bilinear fetch of an un-mipmapped RGBA32F image through a LINEAR / REPEAT sampler at LOD 0
// (Vulkan spec 16.9 "Texel filtering": unnormalised coordinate u*w - 0.5, i0 = floor, weight = fract,
// wrap = i mod size), with fp32 weights -- the convention SURVEY.md "hard part 4" fixes for both sides.
inline vec4 texture(sampler2D const& s, vec2 uv) {
  float x = uv.x * float(s.w) - 0.5f, y = uv.y * float(s.h) - 0.5f;
  if (!std::isfinite(x) || !std::isfinite(y)) return vec4(0.0f);
  float fx = std::floor(x), fy = std::floor(y);
  float a = x - fx, b = y - fy;
  auto wrap = [](long long i, int n) { long long m = i % n; return int(m < 0 ? m + n : m); };
  int x0 = wrap((long long)std::fmod(fx, float(s.w)), s.w), y0 = wrap((long long)std::fmod(fy, float(s.h)), s.h);
  int x1 = wrap(x0 + 1, s.w), y1 = wrap(y0 + 1, s.h);
  auto at = [&](int xi, int yi) { const float* p = s.rgba + 4 * (size_t(yi) * s.w + xi); return vec4(p[0], p[1], p[2], p[3]); };
  // Vulkan spec, texel filtering: tau = (1-a)(1-b) t00 + a(1-b) t10 + (1-a)b t01 + ab t11
  float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  return w00 * at(x0, y0) + w10 * at(x1, y0) + w01 * at(x0, y1) + w11 * at(x1, y1);
}
inline vec4 imageLoad(image2D const& im, ivec2 p) {
  const float* q = im.rgba + 4 * (size_t(p.y) * im.w + p.x);
  return vec4(q[0], q[1], q[2], q[3]);
}
inline void imageStore(image2D const& im, ivec2 p, vec4 v) {
  float* q = im.rgba + 4 * (size_t(p.y) * im.w + p.x);
  q[0] = v.x, q[1] = v.y, q[2] = v.z, q[3] = v.w;
}
#define nonuniformEXT(x) (x)
#define debugPrintfEXT(...) ((void)0)

// ---------------------------------------------------------------- ray-tracing built-ins
static thread_local uvec3 gl_LaunchIDEXT;
static thread_local uvec3 gl_LaunchSizeEXT;
static thread_local int gl_InstanceID;
static thread_local int gl_PrimitiveID;
static thread_local vec3 gl_WorldRayDirectionEXT;
static thread_local vec3 gl_WorldRayOriginEXT;
static thread_local mat4x3 gl_ObjectToWorldEXT;
static thread_local mat4x3 gl_WorldToObjectEXT;
static thread_local vec2 _bary;  // hitAttributeEXT (rchit_layouts.glsl:35)
static thread_local bool isShadowed;  // rayPayload location 1 (rgen:23, shadow.rmiss:10)
const uint gl_RayFlagsCullBackFacingTrianglesEXT = 16u;  // overridden by the instance's cull-disable flag
const uint gl_RayFlagsTerminateOnFirstHitEXT = 4u;
const uint gl_RayFlagsSkipClosestHitShaderEXT = 8u;

// <cmath> defines M_PI as a double; GLSL has no such macro and sun_and_sky.glsl:23-25 defines its own fp32 one.
#undef M_PI
