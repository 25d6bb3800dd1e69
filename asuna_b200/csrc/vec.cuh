// float3 helpers for the device code (GLSL built-in semantics where a shader built-in is mirrored).
#pragma once
#include <cuda_runtime.h>

namespace asuna {

#define ADEV __device__ __forceinline__

ADEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
ADEV float3 f3(float s) { return make_float3(s, s, s); }
ADEV float3 f3(const float* p) { return make_float3(p[0], p[1], p[2]); }
ADEV float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
ADEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
ADEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
ADEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
ADEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
ADEV float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
ADEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
ADEV float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
// One IEEE reciprocal and three multiplies instead of three IEEE divisions (each ~20 instructions; they were 15 % of the
// PBR shader): within Vulkan's 2.5 ulp for OpFDiv, i.e. what a driver may do with the reference's `v / s` as well.
ADEV float3 operator/(float3 a, float s) {
  const float inv = 1.0f / s;
  return f3(a.x * inv, a.y * inv, a.z * inv);
}
ADEV float3 operator+(float3 a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
ADEV float3 operator-(float s, float3 a) { return f3(s - a.x, s - a.y, s - a.z); }
ADEV float3 operator-(float3 a, float s) { return f3(a.x - s, a.y - s, a.z - s); }
ADEV float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
ADEV float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
ADEV float3& operator*=(float3& a, float s) { a = a * s; return a; }
ADEV float3& operator/=(float3& a, float s) { a = a / s; return a; }
ADEV float3& operator/=(float3& a, float3 b) { a = a / b; return a; }
ADEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
ADEV float3 cross(float3 a, float3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
ADEV float length(float3 a) { return sqrtf(dot(a, a)); }
ADEV float3 normalize(float3 a) { return a / length(a); }
ADEV float3 reflect(float3 I, float3 N) { return I - 2.0f * dot(N, I) * N; }
ADEV float3 refract(float3 I, float3 N, float eta) {
  float d = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - d * d);
  if (k < 0.0f) return f3(0.0f);
  return eta * I - (eta * d + sqrtf(k)) * N;
}
ADEV float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
ADEV float3 clamp3(float3 v, float lo, float hi) { return f3(clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)); }
ADEV float3 mix3(float3 a, float3 b, float t) { return a * (1.0f - t) + b * t; }
ADEV float smoothstepf(float e0, float e1, float x) {
  float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
ADEV float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
ADEV float3 exp3(float3 v) { return f3(expf(v.x), expf(v.y), expf(v.z)); }
ADEV float3 pow3(float3 v, float e) { return f3(powf(v.x, e), powf(v.y, e), powf(v.z, e)); }
ADEV float comp(float3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// rows of a 3x4 affine transform
ADEV float3 xf_point(const float4* m, float3 p) {
  return f3(m[0].x * p.x + m[0].y * p.y + m[0].z * p.z + m[0].w, m[1].x * p.x + m[1].y * p.y + m[1].z * p.z + m[1].w,
            m[2].x * p.x + m[2].y * p.y + m[2].z * p.z + m[2].w);
}
ADEV float3 xf_vector(const float4* m, float3 v) {
  return f3(m[0].x * v.x + m[0].y * v.y + m[0].z * v.z, m[1].x * v.x + m[1].y * v.y + m[1].z * v.z,
            m[2].x * v.x + m[2].y * v.y + m[2].z * v.z);
}
// n * M : normals go through the transpose of world->object (GLSL `n * gl_WorldToObjectEXT`)
ADEV float3 xf_normal(const float4* m, float3 n) {
  return f3(m[0].x * n.x + m[1].x * n.y + m[2].x * n.z, m[0].y * n.x + m[1].y * n.y + m[2].y * n.z,
            m[0].z * n.x + m[1].z * n.y + m[2].z * n.z);
}

// column-major 4x4 (nvmath) applied to a point with w divide / to a vector
ADEV float3 mat4_point(const float* m, float3 p) {
  float x = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
  float y = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
  float z = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
  float w = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15];
  return f3(x / w, y / w, z / w);
}
ADEV float3 mat4_vector(const float* m, float3 v) {
  return f3(m[0] * v.x + m[4] * v.y + m[8] * v.z, m[1] * v.x + m[5] * v.y + m[9] * v.z,
            m[2] * v.x + m[6] * v.y + m[10] * v.z);
}
ADEV float3 mat4_vector_transposed(const float* m, float3 v) {
  return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
            m[8] * v.x + m[9] * v.y + m[10] * v.z);
}

}  // namespace asuna
