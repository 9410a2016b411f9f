// Small fixed-size FP64 linear algebra for device (and host) code of the DiffDFSPH step.
// Everything is fully unrolled so matrices live in registers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define DFR_HD __host__ __device__ __forceinline__

namespace dfr {

struct d3 {
  double x, y, z;
};
DFR_HD d3 mk3(double x, double y, double z) {
  d3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
DFR_HD d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
DFR_HD d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
DFR_HD d3 operator-(d3 a) { return mk3(-a.x, -a.y, -a.z); }
DFR_HD d3 operator*(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
DFR_HD d3 operator*(d3 a, double s) { return mk3(s * a.x, s * a.y, s * a.z); }
DFR_HD void operator+=(d3 &a, d3 b) {
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
}
DFR_HD void operator-=(d3 &a, d3 b) {
  a.x -= b.x;
  a.y -= b.y;
  a.z -= b.z;
}
DFR_HD double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DFR_HD d3 cross(d3 a, d3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

template <int R, int C>
struct Mat {
  double a[R * C];
  DFR_HD double &operator()(int i, int j) { return a[i * C + j]; }
  DFR_HD double operator()(int i, int j) const { return a[i * C + j]; }
  DFR_HD static Mat zero() {
    Mat m;
#pragma unroll
    for (int i = 0; i < R * C; i++) m.a[i] = 0.0;
    return m;
  }
  DFR_HD static Mat identity() {
    Mat m = zero();
#pragma unroll
    for (int i = 0; i < (R < C ? R : C); i++) m.a[i * C + i] = 1.0;
    return m;
  }
};
typedef Mat<3, 3> m33;
typedef Mat<3, 4> m34;
typedef Mat<4, 3> m43;
typedef Mat<4, 4> m44;
typedef Mat<4, 1> v4;

template <int R, int C>
DFR_HD Mat<R, C> operator+(const Mat<R, C> &x, const Mat<R, C> &y) {
  Mat<R, C> m;
#pragma unroll
  for (int i = 0; i < R * C; i++) m.a[i] = x.a[i] + y.a[i];
  return m;
}
template <int R, int C>
DFR_HD Mat<R, C> operator-(const Mat<R, C> &x, const Mat<R, C> &y) {
  Mat<R, C> m;
#pragma unroll
  for (int i = 0; i < R * C; i++) m.a[i] = x.a[i] - y.a[i];
  return m;
}
template <int R, int C>
DFR_HD void operator+=(Mat<R, C> &x, const Mat<R, C> &y) {
#pragma unroll
  for (int i = 0; i < R * C; i++) x.a[i] += y.a[i];
}
template <int R, int C>
DFR_HD Mat<R, C> operator*(double s, const Mat<R, C> &x) {
  Mat<R, C> m;
#pragma unroll
  for (int i = 0; i < R * C; i++) m.a[i] = s * x.a[i];
  return m;
}
template <int R, int K, int C>
DFR_HD Mat<R, C> operator*(const Mat<R, K> &x, const Mat<K, C> &y) {
  Mat<R, C> m;
#pragma unroll
  for (int i = 0; i < R; i++)
#pragma unroll
    for (int j = 0; j < C; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) s += x.a[i * K + k] * y.a[k * C + j];
      m.a[i * C + j] = s;
    }
  return m;
}
template <int R, int C>
DFR_HD Mat<C, R> transpose(const Mat<R, C> &x) {
  Mat<C, R> m;
#pragma unroll
  for (int i = 0; i < R; i++)
#pragma unroll
    for (int j = 0; j < C; j++) m.a[j * R + i] = x.a[i * C + j];
  return m;
}
DFR_HD d3 operator*(const m33 &m, d3 v) {
  return mk3(m.a[0] * v.x + m.a[1] * v.y + m.a[2] * v.z, m.a[3] * v.x + m.a[4] * v.y + m.a[5] * v.z,
             m.a[6] * v.x + m.a[7] * v.y + m.a[8] * v.z);
}
DFR_HD m33 outer(d3 p, d3 q) {
  m33 m;
  m.a[0] = p.x * q.x; m.a[1] = p.x * q.y; m.a[2] = p.x * q.z;
  m.a[3] = p.y * q.x; m.a[4] = p.y * q.y; m.a[5] = p.y * q.z;
  m.a[6] = p.z * q.x; m.a[7] = p.z * q.y; m.a[8] = p.z * q.z;
  return m;
}
// [v]x  (GradientUtils.cpp:4-6)
DFR_HD m33 skew(d3 v) {
  m33 m;
  m.a[0] = 0.0;  m.a[1] = -v.z; m.a[2] = v.y;
  m.a[3] = v.z;  m.a[4] = 0.0;  m.a[5] = -v.x;
  m.a[6] = -v.y; m.a[7] = v.x;  m.a[8] = 0.0;
  return m;
}
DFR_HD m33 inverse(const m33 &m) {
  const double c00 = m.a[4] * m.a[8] - m.a[5] * m.a[7];
  const double c01 = m.a[5] * m.a[6] - m.a[3] * m.a[8];
  const double c02 = m.a[3] * m.a[7] - m.a[4] * m.a[6];
  const double id = 1.0 / (m.a[0] * c00 + m.a[1] * c01 + m.a[2] * c02);
  m33 r;
  r.a[0] = c00 * id;
  r.a[3] = c01 * id;
  r.a[6] = c02 * id;
  r.a[1] = (m.a[2] * m.a[7] - m.a[1] * m.a[8]) * id;
  r.a[4] = (m.a[0] * m.a[8] - m.a[2] * m.a[6]) * id;
  r.a[7] = (m.a[1] * m.a[6] - m.a[0] * m.a[7]) * id;
  r.a[2] = (m.a[1] * m.a[5] - m.a[2] * m.a[4]) * id;
  r.a[5] = (m.a[2] * m.a[3] - m.a[0] * m.a[5]) * id;
  r.a[8] = (m.a[0] * m.a[4] - m.a[1] * m.a[3]) * id;
  return r;
}

// quaternion (w, x, y, z), Hamilton product
struct quat {
  double w, x, y, z;
};
DFR_HD quat qmul(quat a, quat b) {
  quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
DFR_HD double qnorm(quat q) { return sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); }
DFR_HD m33 qrot(quat q) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  m33 r;
  r.a[0] = 1.0 - (tyy + tzz); r.a[1] = txy - twz;         r.a[2] = txz + twy;
  r.a[3] = txy + twz;         r.a[4] = 1.0 - (txx + tzz); r.a[5] = tyz - twx;
  r.a[6] = txz - twy;         r.a[7] = tyz + twx;         r.a[8] = 1.0 - (txx + tyy);
  return r;
}

// d(R(q) p)/dq and d(R(q)^T p)/dq, columns (w,x,y,z)   (GradientUtils.cpp:9-34); sgn = -1 / +1
DFR_HD m34 grad_Rqp_to_q(quat q, d3 p, double sgn) {
  const d3 qv = mk3(q.x, q.y, q.z);
  const d3 pc = cross(p, qv);
  const d3 t1 = 2.0 * (q.w * p + sgn * pc);
  const double qp = dot(qv, p);
  const m33 S = skew(p);
  const m33 qpT = outer(qv, p), pqT = outer(p, qv);
  m34 r;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      r.a[i * 4 + j + 1] = 2.0 * ((i == j ? qp : 0.0) + qpT.a[i * 3 + j] - pqT.a[i * 3 + j] + sgn * q.w * S.a[i * 3 + j]);
    }
  }
  r.a[0] = t1.x;
  r.a[4] = t1.y;
  r.a[8] = t1.z;
  return r;
}
// GradientUtils.cpp:36-43
DFR_HD m44 grad_pq_to_q(quat p) {
  m44 m;
  m.a[0] = p.w;  m.a[1] = -p.x; m.a[2] = -p.y;  m.a[3] = -p.z;
  m.a[4] = p.x;  m.a[5] = p.w;  m.a[6] = -p.z;  m.a[7] = p.y;
  m.a[8] = p.y;  m.a[9] = p.z;  m.a[10] = p.w;  m.a[11] = -p.x;
  m.a[12] = p.z; m.a[13] = -p.y; m.a[14] = p.x; m.a[15] = p.w;
  return m;
}
// GradientUtils.cpp:45-53
DFR_HD m43 grad_omega_q_to_omega(quat q) {
  m43 m;
  m.a[0] = -q.x; m.a[1] = -q.y; m.a[2] = -q.z;
  m.a[3] = q.w;  m.a[4] = q.z;  m.a[5] = -q.y;
  m.a[6] = -q.z; m.a[7] = q.w;  m.a[8] = q.x;
  m.a[9] = q.y;  m.a[10] = -q.x; m.a[11] = q.w;
  return m;
}

}  // namespace dfr
