// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// GLSL-flavoured fp32 vector helpers used by the restatement of the reference shaders.
// Built-ins follow the GLSL 4.60 spec definitions (reflect, refract, mix, clamp,
// smoothstep, sign, normalize, atan(y,x)); nothing here comes from the product.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct vec2 {
  float x, y;
};
struct vec3 {
  float x, y, z;
  vec3() : x(0), y(0), z(0) {}
  explicit vec3(float s) : x(s), y(s), z(s) {}
  vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  explicit vec3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4 {
  float x, y, z, w;
};

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator+(vec3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(float s, vec3 a) { return {s - a.x, s - a.y, s - a.z}; }
inline vec3 operator-(vec3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline vec3& operator+=(vec3& a, vec3 b) { return a = a + b; }
inline vec3& operator*=(vec3& a, vec3 b) { return a = a * b; }
inline vec3& operator*=(vec3& a, float s) { return a = a * s; }
inline vec3& operator/=(vec3& a, float s) { return a = a / s; }
inline vec3& operator/=(vec3& a, vec3 b) { return a = a / b; }

inline vec2 operator*(vec2 a, float s) { return {a.x * s, a.y * s}; }
inline vec2 operator*(float s, vec2 a) { return {a.x * s, a.y * s}; }
inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }

inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / length(a); }
inline vec3 reflect(vec3 I, vec3 N) { return I - 2.0f * dot(N, I) * N; }
inline vec3 refract(vec3 I, vec3 N, float eta) {
  float d = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - d * d);
  if (k < 0.0f) return vec3(0.0f);
  return eta * I - (eta * d + std::sqrt(k)) * N;
}
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline vec3 clamp3(vec3 v, float lo, float hi) {
  return {clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)};
}
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix3(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline float smoothstepf(float e0, float e1, float x) {
  float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
inline float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline vec3 exp3(vec3 v) { return {std::exp(v.x), std::exp(v.y), std::exp(v.z)}; }
inline vec3 pow3(vec3 v, float e) { return {std::pow(v.x, e), std::pow(v.y, e), std::pow(v.z, e)}; }

inline int32_t floatBitsToInt(float f) {
  int32_t i;
  std::memcpy(&i, &f, 4);
  return i;
}
inline float intBitsToFloat(int32_t i) {
  float f;
  std::memcpy(&f, &i, 4);
  return f;
}

// Column-major 4x4 (nvmath::mat4f layout: m[col*4+row]).
struct mat4 {
  float m[16];
  float at(int r, int c) const { return m[c * 4 + r]; }
};
inline vec4 mul(const mat4& M, vec4 v) {
  return {M.at(0, 0) * v.x + M.at(0, 1) * v.y + M.at(0, 2) * v.z + M.at(0, 3) * v.w,
          M.at(1, 0) * v.x + M.at(1, 1) * v.y + M.at(1, 2) * v.z + M.at(1, 3) * v.w,
          M.at(2, 0) * v.x + M.at(2, 1) * v.y + M.at(2, 2) * v.z + M.at(2, 3) * v.w,
          M.at(3, 0) * v.x + M.at(3, 1) * v.y + M.at(3, 2) * v.z + M.at(3, 3) * v.w};
}
inline mat4 transpose(const mat4& M) {
  mat4 T;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) T.m[c * 4 + r] = M.m[r * 4 + c];
  return T;
}

}  // namespace orc
