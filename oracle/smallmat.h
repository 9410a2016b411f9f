// TEST INFRASTRUCTURE ONLY — part of the CPU oracle (see dfsph_oracle.cpp header).
// Minimal fixed-size linear algebra for the oracle. Deliberately independent of the
// product's device-side helpers (difffr_b200/csrc/dfr_math.cuh).
#pragma once
#include <cmath>

namespace orc {

template <int R, int C>
struct Mat {
  double a[R][C];
  static Mat zero() {
    Mat m;
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) m.a[i][j] = 0.0;
    return m;
  }
  static Mat identity() {
    Mat m = zero();
    for (int i = 0; i < (R < C ? R : C); i++) m.a[i][i] = 1.0;
    return m;
  }
  double &operator()(int i, int j) { return a[i][j]; }
  double operator()(int i, int j) const { return a[i][j]; }
  Mat &operator+=(const Mat &o) {
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) a[i][j] += o.a[i][j];
    return *this;
  }
};

typedef Mat<3, 1> Vec3;
typedef Mat<4, 1> Vec4;
typedef Mat<3, 3> Mat3;
typedef Mat<3, 4> Mat34;
typedef Mat<4, 3> Mat43;
typedef Mat<4, 4> Mat4;

template <int R, int C>
inline Mat<R, C> operator+(const Mat<R, C> &x, const Mat<R, C> &y) {
  Mat<R, C> m;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) m.a[i][j] = x.a[i][j] + y.a[i][j];
  return m;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C> &x, const Mat<R, C> &y) {
  Mat<R, C> m;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) m.a[i][j] = x.a[i][j] - y.a[i][j];
  return m;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C> &x) {
  Mat<R, C> m;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) m.a[i][j] = -x.a[i][j];
  return m;
}
template <int R, int C>
inline Mat<R, C> operator*(double s, const Mat<R, C> &x) {
  Mat<R, C> m;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) m.a[i][j] = s * x.a[i][j];
  return m;
}
template <int R, int C>
inline Mat<R, C> operator*(const Mat<R, C> &x, double s) {
  return s * x;
}
template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K> &x, const Mat<K, C> &y) {
  Mat<R, C> m;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) {
      double s = 0.0;
      for (int k = 0; k < K; k++) s += x.a[i][k] * y.a[k][j];
      m.a[i][j] = s;
    }
  return m;
}
template <int R, int C>
inline Mat<C, R> transpose(const Mat<R, C> &x) {
  Mat<C, R> m;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) m.a[j][i] = x.a[i][j];
  return m;
}

inline Vec3 vec3(double x, double y, double z) {
  Vec3 v;
  v.a[0][0] = x;
  v.a[1][0] = y;
  v.a[2][0] = z;
  return v;
}
inline double X(const Vec3 &v) { return v.a[0][0]; }
inline double Y(const Vec3 &v) { return v.a[1][0]; }
inline double Z(const Vec3 &v) { return v.a[2][0]; }
inline double dot(const Vec3 &p, const Vec3 &q) { return X(p) * X(q) + Y(p) * Y(q) + Z(p) * Z(q); }
inline double sqnorm(const Vec3 &p) { return dot(p, p); }
inline double norm(const Vec3 &p) { return std::sqrt(sqnorm(p)); }
inline Vec3 cross(const Vec3 &p, const Vec3 &q) {
  return vec3(Y(p) * Z(q) - Z(p) * Y(q), Z(p) * X(q) - X(p) * Z(q), X(p) * Y(q) - Y(p) * X(q));
}
// p q^T
inline Mat3 outer(const Vec3 &p, const Vec3 &q) { return p * transpose(q); }

inline Mat3 inverse(const Mat3 &m) {
  // adjugate / determinant, the closed form Eigen uses for 3x3
  const double c00 = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1);
  const double c01 = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2);
  const double c02 = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
  const double det = m(0, 0) * c00 + m(0, 1) * c01 + m(0, 2) * c02;
  const double id = 1.0 / det;
  Mat3 r;
  r(0, 0) = c00 * id;
  r(1, 0) = c01 * id;
  r(2, 0) = c02 * id;
  r(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) * id;
  r(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) * id;
  r(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) * id;
  r(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) * id;
  r(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) * id;
  r(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) * id;
  return r;
}

// Quaternion stored (w, x, y, z); Hamilton product like Eigen::Quaternion::operator*.
struct Quat {
  double w, x, y, z;
};
inline Quat qmul(const Quat &a, const Quat &b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
inline double qnorm(const Quat &q) { return std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); }
inline Quat qnormalized(const Quat &q) {
  const double n = qnorm(q);
  Quat r = {q.w / n, q.x / n, q.y / n, q.z / n};
  return r;
}
inline Mat3 qrot(const Quat &q) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3 r;
  r(0, 0) = 1.0 - (tyy + tzz);
  r(0, 1) = txy - twz;
  r(0, 2) = txz + twy;
  r(1, 0) = txy + twz;
  r(1, 1) = 1.0 - (txx + tzz);
  r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy;
  r(2, 1) = tyz + twx;
  r(2, 2) = 1.0 - (txx + tyy);
  return r;
}

}  // namespace orc
