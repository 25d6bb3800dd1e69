// 4x4 float matrices for the scene host, following the conventions the reference takes from nvmath
// (src/ext/nvpro_core/nvmath): column-major storage at the ABI, JSON matrices row-major in the file
// (src/loader/utils.h:37-39).  Internally a Mat4 is addressed m(r, c); to_colmajor() produces the wire layout.
#pragma once
#include <cmath>
#include <cstring>

namespace asuna_host {

struct Vec3 {
  float x = 0, y = 0, z = 0;
  Vec3() = default;
  Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(Vec3 a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(Vec3 a) {
  float l = length(a);
  return l > 0 ? a * (1.0f / l) : a;
}

struct Mat4 {
  float m[4][4];  // m[row][col]
  float& operator()(int r, int c) { return m[r][c]; }
  float operator()(int r, int c) const { return m[r][c]; }
  static Mat4 identity() {
    Mat4 a;
    std::memset(a.m, 0, sizeof a.m);
    for (int i = 0; i < 4; i++) a.m[i][i] = 1.f;
    return a;
  }
  static Mat4 zero() {
    Mat4 a;
    std::memset(a.m, 0, sizeof a.m);
    return a;
  }
  void to_colmajor(float out[16]) const {
    for (int c = 0; c < 4; c++)
      for (int r = 0; r < 4; r++) out[c * 4 + r] = m[r][c];
  }
};
inline Mat4 operator*(const Mat4& a, const Mat4& b) {  // fp32 accumulation in index order, like numpy's matmul on f32
  Mat4 r = Mat4::zero();
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float s = 0.f;
      for (int k = 0; k < 4; k++) s += a.m[i][k] * b.m[k][j];
      r.m[i][j] = s;
    }
  return r;
}
inline Mat4 translation(Vec3 v) {
  Mat4 a = Mat4::identity();
  a(0, 3) = v.x, a(1, 3) = v.y, a(2, 3) = v.z;
  return a;
}
inline Mat4 scaling(Vec3 v) {
  Mat4 a = Mat4::identity();
  a(0, 0) = v.x, a(1, 1) = v.y, a(2, 2) = v.z;
  return a;
}
// nvmath rotation_mat4_x/y/z (ext/nvpro_core/nvmath/nvmath.inl)
inline Mat4 rotation_x(float a) {
  Mat4 r = Mat4::identity();
  float c = std::cos(a), s = std::sin(a);
  r(1, 1) = c, r(1, 2) = -s, r(2, 1) = s, r(2, 2) = c;
  return r;
}
inline Mat4 rotation_y(float a) {
  Mat4 r = Mat4::identity();
  float c = std::cos(a), s = std::sin(a);
  r(0, 0) = c, r(0, 2) = s, r(2, 0) = -s, r(2, 2) = c;
  return r;
}
inline Mat4 rotation_z(float a) {
  Mat4 r = Mat4::identity();
  float c = std::cos(a), s = std::sin(a);
  r(0, 0) = c, r(0, 1) = -s, r(1, 0) = s, r(1, 1) = c;
  return r;
}
// nvmath look_at (nvmath.inl:979-1019)
inline Mat4 look_at(Vec3 eye, Vec3 center, Vec3 up) {
  Vec3 z = normalize(eye - center);
  Vec3 x = normalize(cross(up, z));
  Vec3 y = cross(z, x);
  Mat4 r = Mat4::identity();
  r(0, 0) = x.x, r(0, 1) = x.y, r(0, 2) = x.z, r(0, 3) = -dot(x, eye);
  r(1, 0) = y.x, r(1, 1) = y.y, r(1, 2) = y.z, r(1, 3) = -dot(y, eye);
  r(2, 0) = z.x, r(2, 1) = z.y, r(2, 2) = z.z, r(2, 3) = -dot(z, eye);
  return r;
}
// nvmath invert_rot_trans (nvmath.inl:911-931): transpose the 3x3 block, rotate the negated translation
inline Mat4 invert_rot_trans(const Mat4& a) {
  Mat4 r = Mat4::identity();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(j, i);
  for (int i = 0; i < 3; i++) r(i, 3) = -(r(i, 0) * a(0, 3) + r(i, 1) * a(1, 3) + r(i, 2) * a(2, 3));
  return r;
}
inline Vec3 xf_point(const Mat4& m, Vec3 p, float w) {
  return {m(0, 0) * p.x + m(0, 1) * p.y + m(0, 2) * p.z + m(0, 3) * w, m(1, 0) * p.x + m(1, 1) * p.y + m(1, 2) * p.z + m(1, 3) * w,
          m(2, 0) * p.x + m(2, 1) * p.y + m(2, 2) * p.z + m(2, 3) * w};
}
// general 4x4 inverse in double (Gauss-Jordan with partial pivoting)
inline bool invert(const Mat4& a, Mat4& out) {
  double w[4][8];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) w[i][j] = a(i, j), w[i][4 + j] = i == j ? 1.0 : 0.0;
  for (int c = 0; c < 4; c++) {
    int p = c;
    for (int r = c + 1; r < 4; r++)
      if (std::fabs(w[r][c]) > std::fabs(w[p][c])) p = r;
    if (w[p][c] == 0.0) return false;
    if (p != c)
      for (int j = 0; j < 8; j++) std::swap(w[p][j], w[c][j]);
    double inv = 1.0 / w[c][c];
    for (int j = 0; j < 8; j++) w[c][j] *= inv;
    for (int r = 0; r < 4; r++)
      if (r != c) {
        double f = w[r][c];
        if (f != 0.0)
          for (int j = 0; j < 8; j++) w[r][j] -= f * w[c][j];
      }
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) out(i, j) = (float)w[i][4 + j];
  return true;
}

}  // namespace asuna_host
