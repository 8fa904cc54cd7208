// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference algorithm; never shipped,
// never on the product path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, link or execute anything under oracle/.
//
// PARITY UNPINNED: the reference (IRVLab/SOS-SLAM @ a4849ab) ships no tests, golden vectors or
// fixtures for this path and cannot be compiled here (Eigen3 — un-vendored, unpinned
// `find_package(Eigen3 REQUIRED)`, CMakeLists.txt:8 — plus Boost/OpenCV/Pangolin/PCL/ROS are absent).
// This oracle restates the published algorithm from the reference sources it cites and is pinned
// only by analytic self-checks (tests/test_oracle_*.py): finite-difference Jacobians, a dense fp64
// numpy Schur complement, Sophus' own exp/log/Adj round-trip properties.
//
// orc_math.h: minimal fixed-size linear algebra + SE3 (restating thirdparty/Sophus/sophus/so3.hpp,
// se3.hpp of Sophus 0.9a, vendored in the reference) + pivoted LDLT (restating the published
// algorithm of Eigen::LDLT, the solver called at EnergyFunctional.cpp:1147-1148 and
// CoarseTracker.cpp:423-443).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>

namespace orc {

// ---------------------------------------------------------------------------------------------
// 3x3 helpers (row-major), templated on scalar
template <class T> struct M3 { T m[9]; T &operator()(int r, int c) { return m[r * 3 + c]; } const T &operator()(int r, int c) const { return m[r * 3 + c]; } };
template <class T> struct V3 { T v[3]; T &operator[](int i) { return v[i]; } const T &operator[](int i) const { return v[i]; } };

template <class T> inline M3<T> mul(const M3<T> &a, const M3<T> &b) {
  M3<T> r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
  return r;
}
template <class T> inline V3<T> mul(const M3<T> &a, const V3<T> &x) {
  V3<T> r;
  for (int i = 0; i < 3; i++) r[i] = (a(i, 0) * x[0] + a(i, 1) * x[1]) + a(i, 2) * x[2];
  return r;
}
template <class T> inline M3<T> transpose(const M3<T> &a) {
  M3<T> r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(j, i);
  return r;
}
template <class T> inline M3<T> identity3() { M3<T> r; for (int i = 0; i < 9; i++) r.m[i] = 0; r(0, 0) = r(1, 1) = r(2, 2) = 1; return r; }
inline M3<double> hat(const V3<double> &w) {
  M3<double> r;
  r(0, 0) = 0; r(0, 1) = -w[2]; r(0, 2) = w[1];
  r(1, 0) = w[2]; r(1, 1) = 0; r(1, 2) = -w[0];
  r(2, 0) = -w[1]; r(2, 1) = w[0]; r(2, 2) = 0;
  return r;
}

// ---------------------------------------------------------------------------------------------
// SO3 / SE3, double, unit-quaternion storage as in Sophus 0.9a.
struct Quat { double w, x, y, z; };
inline Quat qmul(const Quat &a, const Quat &b) {
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Quat qnormalize(const Quat &q) {
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return Quat{q.w / n, q.x / n, q.y / n, q.z / n};
}
inline Quat qconj(const Quat &q) { return Quat{q.w, -q.x, -q.y, -q.z}; }
inline M3<double> qmatrix(const Quat &q) {  // Eigen::Quaternion::toRotationMatrix
  M3<double> r;
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
  r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
  return r;
}
inline Quat qfrommatrix(const M3<double> &m) {  // Eigen quaternion from rotation matrix (Shoemake)
  Quat q;
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t; t = 0.5 / t;
    q.x = (m(2, 1) - m(1, 2)) * t; q.y = (m(0, 2) - m(2, 0)) * t; q.z = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    double v[3];
    v[i] = 0.5 * t; t = 0.5 / t;
    q.w = (m(k, j) - m(j, k)) * t;
    v[j] = (m(j, i) + m(i, j)) * t; v[k] = (m(k, i) + m(i, k)) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}
inline V3<double> qrotate(const Quat &q, const V3<double> &v) {  // Eigen _transformVector
  V3<double> uv{{2 * (q.y * v[2] - q.z * v[1]), 2 * (q.z * v[0] - q.x * v[2]), 2 * (q.x * v[1] - q.y * v[0])}};
  return V3<double>{{v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]), v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]),
                     v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0])}};
}

static const double kSophusEps = 1e-10;  // SophusConstants<double>::epsilon, sophus.hpp:45-47

// so3.hpp:343-369 (the 1/384 coefficient is the reference's, kept verbatim)
inline Quat so3_exp_theta(const V3<double> &omega, double *theta) {
  const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  *theta = std::sqrt(theta_sq);
  const double half_theta = 0.5 * (*theta);
  double imag_factor, real_factor;
  if ((*theta) < kSophusEps) {
    const double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real_factor = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    const double s = std::sin(half_theta);
    imag_factor = s / (*theta);
    real_factor = std::cos(half_theta);
  }
  return Quat{real_factor, imag_factor * omega[0], imag_factor * omega[1], imag_factor * omega[2]};
}
// so3.hpp:491-531
inline V3<double> so3_log_theta(const Quat &q, double *theta) {
  const double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  const double n = std::sqrt(squared_n);
  const double w = q.w;
  double f;
  if (n < kSophusEps) {
    const double squared_w = w * w;
    f = 2.0 / w - 2.0 * squared_n / (w * squared_w);
  } else if (std::fabs(w) < kSophusEps) {
    f = (w > 0 ? M_PI : -M_PI) / n;
  } else {
    f = 2.0 * std::atan(n / w) / n;
  }
  *theta = f * n;
  return V3<double>{{f * q.x, f * q.y, f * q.z}};
}

struct SE3 {
  Quat q{1, 0, 0, 0};
  V3<double> t{{0, 0, 0}};
  M3<double> R() const { return qmatrix(q); }
  SE3 inverse() const {  // se3.hpp:168-173
    SE3 r; r.q = qconj(q);
    V3<double> mt{{-t[0], -t[1], -t[2]}};
    r.t = qrotate(r.q, mt);
    return r;
  }
  static SE3 from_Rt(const M3<double> &R, const V3<double> &t) { SE3 r; r.q = qnormalize(qfrommatrix(R)); r.t = t; return r; }
  static SE3 from_rowmajor34(const double *p) {
    M3<double> R; V3<double> t;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R(i, j) = p[i * 4 + j]; t[i] = p[i * 4 + 3]; }
    return from_Rt(R, t);
  }
  void to_rowmajor34(double *p) const {
    M3<double> Rm = R();
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) p[i * 4 + j] = Rm(i, j); p[i * 4 + 3] = t[i]; }
  }
};
inline SE3 operator*(const SE3 &a, const SE3 &b) {  // se3.hpp operator*: t += R*t2 ; q *= q2 ; normalize
  SE3 r;
  V3<double> rt = qrotate(a.q, b.t);
  r.t = V3<double>{{a.t[0] + rt[0], a.t[1] + rt[1], a.t[2] + rt[2]}};
  r.q = qnormalize(qmul(a.q, b.q));
  return r;
}
// se3.hpp:417-439 ; tangent = [upsilon(3) omega(3)]
inline SE3 se3_exp(const double a[6]) {
  V3<double> omega{{a[3], a[4], a[5]}}, ups{{a[0], a[1], a[2]}};
  double theta;
  SE3 r;
  r.q = so3_exp_theta(omega, &theta);
  M3<double> Omega = hat(omega), Omega_sq = mul(Omega, Omega), V;
  if (theta < kSophusEps) {
    V = qmatrix(r.q);
  } else {
    const double theta_sq = theta * theta;
    const double c1 = (1.0 - std::cos(theta)) / theta_sq, c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    V = identity3<double>();
    for (int i = 0; i < 9; i++) V.m[i] = V.m[i] + c1 * Omega.m[i] + c2 * Omega_sq.m[i];
  }
  r.t = mul(V, ups);
  return r;
}
// se3.hpp:560-586
inline void se3_log(const SE3 &T, double out[6]) {
  double theta;
  V3<double> omega = so3_log_theta(T.q, &theta);
  M3<double> Omega = hat(omega), Osq = mul(Omega, Omega), Vinv = identity3<double>();
  double c;
  if (std::fabs(theta) < kSophusEps) c = 1.0 / 12.0;
  else c = (1.0 - theta / (2.0 * std::tan(theta / 2.0))) / (theta * theta);
  for (int i = 0; i < 9; i++) Vinv.m[i] = Vinv.m[i] - 0.5 * Omega.m[i] + c * Osq.m[i];
  V3<double> u = mul(Vinv, T.t);
  out[0] = u[0]; out[1] = u[1]; out[2] = u[2]; out[3] = omega[0]; out[4] = omega[1]; out[5] = omega[2];
}
// se3.hpp:131-139 ; 6x6 row-major
inline void se3_adj(const SE3 &T, double A[36]) {
  M3<double> R = T.R(), tR = mul(hat(T.t), R);
  for (int i = 0; i < 36; i++) A[i] = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { A[i * 6 + j] = R(i, j); A[(i + 3) * 6 + (j + 3)] = R(i, j); A[i * 6 + (j + 3)] = tR(i, j); }
}

// ---------------------------------------------------------------------------------------------
// Pivoted LDL^T of a symmetric (possibly indefinite) matrix, restating Eigen::LDLT (robust Cholesky
// with diagonal pivoting): at step k bring the largest |diagonal| of the trailing block to position k
// by a symmetric transposition, then eliminate.  Solves A x = rhs.  A is n*n row-major (lower used).
inline void ldlt_solve(const double *A_in, const double *rhs, double *x, int n) {
  std::vector<double> m(A_in, A_in + (size_t)n * n), temp(n);
  std::vector<int> tr(n);
#define MAT(r, c) m[(size_t)(r) * n + (c)]
  for (int k = 0; k < n; k++) {
    int big = k; double bigv = std::fabs(MAT(k, k));
    for (int i = k + 1; i < n; i++) if (std::fabs(MAT(i, i)) > bigv) { bigv = std::fabs(MAT(i, i)); big = i; }
    tr[k] = big;
    if (big != k) {  // symmetric swap on the lower triangle
      int s = n - big - 1;
      for (int c = 0; c < k; c++) std::swap(MAT(k, c), MAT(big, c));
      for (int r = 0; r < s; r++) std::swap(MAT(big + 1 + r, k), MAT(big + 1 + r, big));
      std::swap(MAT(k, k), MAT(big, big));
      for (int i = k + 1; i < big; i++) std::swap(MAT(i, k), MAT(big, i));
    }
    int rs = n - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; j++) temp[j] = MAT(j, j) * MAT(k, j);
      double s = 0; for (int j = 0; j < k; j++) s += MAT(k, j) * temp[j];
      MAT(k, k) -= s;
      for (int i = 0; i < rs; i++) { double a = 0; for (int j = 0; j < k; j++) a += MAT(k + 1 + i, j) * temp[j]; MAT(k + 1 + i, k) -= a; }
    }
    double akk = MAT(k, k);
    if (rs > 0 && std::fabs(akk) > 0) for (int i = 0; i < rs; i++) MAT(k + 1 + i, k) /= akk;
  }
  std::vector<double> y(rhs, rhs + n);
  for (int k = 0; k < n; k++) std::swap(y[k], y[tr[k]]);                                   // P b
  for (int i = 0; i < n; i++) { double s = y[i]; for (int j = 0; j < i; j++) s -= MAT(i, j) * y[j]; y[i] = s; }  // L
  const double tol = 1.0 / std::numeric_limits<double>::max();
  for (int i = 0; i < n; i++) { if (std::fabs(MAT(i, i)) > tol) y[i] /= MAT(i, i); else y[i] = 0; }             // D
  for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int j = i + 1; j < n; j++) s -= MAT(j, i) * y[j]; y[i] = s; }  // L^T
  for (int k = n - 1; k >= 0; k--) std::swap(y[k], y[tr[k]]);                              // P^T
  for (int i = 0; i < n; i++) x[i] = y[i];
#undef MAT
}

}  // namespace orc
