// lm_math.cuh — fp64 device algebra of the direct-alignment control loops (k_lm.cu), for ONE thread:
//   * rigid motions as Sophus 0.9a keeps them: unit quaternion + translation (thirdparty/Sophus/sophus/se3.hpp, so3.hpp);
//     SE3::exp of [upsilon, omega] (se3.hpp:417-439 with SO3::expAndTheta so3.hpp:343-369), group product (se3.hpp: rotate the
//     right translation, multiply and re-normalise the quaternions), Eigen's quaternion -> matrix expansion;
//   * Eigen::LDLT (robust Cholesky with diagonal pivoting) of an n x n system, n <= 8: what `Hl.ldlt().solve(-b)` runs
//     (CoarseTracker.cpp:423-443).
// Compiled with -fmad=false like the rest of the path that feeds integer bookkeeping.
#pragma once
#include <math.h>

namespace lm {

struct Pose {
  double qx, qy, qz, qw;   // unit quaternion, Eigen coefficient order
  double t[3];
};

__device__ __forceinline__ void quat_matrix(const Pose &p, double R[9]) {   // Eigen::QuaternionBase::toRotationMatrix
  const double tx = 2 * p.qx, ty = 2 * p.qy, tz = 2 * p.qz;
  const double twx = tx * p.qw, twy = ty * p.qw, twz = tz * p.qw;
  const double txx = tx * p.qx, txy = ty * p.qx, txz = tz * p.qx;
  const double tyy = ty * p.qy, tyz = tz * p.qy, tzz = tz * p.qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

// v' = q v q^-1 the way Eigen evaluates it: uv = 2 (q_vec x v); v + w uv + q_vec x uv
__device__ __forceinline__ void quat_rotate(const Pose &p, const double v[3], double out[3]) {
  const double ux = 2 * (p.qy * v[2] - p.qz * v[1]), uy = 2 * (p.qz * v[0] - p.qx * v[2]), uz = 2 * (p.qx * v[1] - p.qy * v[0]);
  out[0] = v[0] + p.qw * ux + (p.qy * uz - p.qz * uy);
  out[1] = v[1] + p.qw * uy + (p.qz * ux - p.qx * uz);
  out[2] = v[2] + p.qw * uz + (p.qx * uy - p.qy * ux);
}

// SE3::exp(xi), xi = [upsilon(3), omega(3)]
__device__ __forceinline__ Pose se3_exp(const double xi[6]) {
  const double *ups = xi, *om = xi + 3;
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2], th = sqrt(th2), half = 0.5 * th;
  double imag, real;
  if (th < 1e-10) {   // SophusConstants<double>::epsilon
    const double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - 0.5 * th2 + (1.0 / 384.0) * th4;
  } else {
    imag = sin(half) / th;
    real = cos(half);
  }
  Pose r;
  r.qw = real; r.qx = imag * om[0]; r.qy = imag * om[1]; r.qz = imag * om[2];
  // V = I + c1 hat(om) + c2 hat(om)^2 (V = R(q) for tiny angles)
  const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double O2[9], V[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) O2[3 * i + j] = (O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j]) + O[3 * i + 2] * O[6 + j];
  if (th < 1e-10) quat_matrix(r, V);
  else {
    const double c1 = (1.0 - cos(th)) / th2, c2 = (th - sin(th)) / (th2 * th);
    for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * O[i] + c2 * O2[i];
  }
  for (int i = 0; i < 3; i++) r.t[i] = (V[3 * i] * ups[0] + V[3 * i + 1] * ups[1]) + V[3 * i + 2] * ups[2];
  return r;
}

// a * b
__device__ __forceinline__ Pose se3_mul(const Pose &a, const Pose &b) {
  Pose r;
  double rt[3];
  quat_rotate(a, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
  const double w = a.qw * b.qw - a.qx * b.qx - a.qy * b.qy - a.qz * b.qz;
  const double x = a.qw * b.qx + a.qx * b.qw + a.qy * b.qz - a.qz * b.qy;
  const double y = a.qw * b.qy + a.qy * b.qw + a.qz * b.qx - a.qx * b.qz;
  const double z = a.qw * b.qz + a.qz * b.qw + a.qx * b.qy - a.qy * b.qx;
  const double n = sqrt(w * w + x * x + y * y + z * z);
  r.qw = w / n; r.qx = x / n; r.qy = y / n; r.qz = z / n;
  return r;
}

// Solve A x = rhs with the in-place pivoted LDL^T of Eigen::LDLT<Lower>: at step k the largest |diagonal| of the trailing
// block is brought to position k by a symmetric transposition; x = P^T L^-T D^-1 L^-1 P rhs.  A: n x n row-major, lower
// triangle read, destroyed.
template <int NMAX>
__device__ __forceinline__ void ldlt_solve(double *A, int n, const double *rhs, double *x) {
  int perm[NMAX];
  double tmp[NMAX];
#define LM_A(r, c) A[(r) * n + (c)]
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = fabs(LM_A(k, k));
    for (int i = k + 1; i < n; i++) { const double v = fabs(LM_A(i, i)); if (v > best) { best = v; p = i; } }
    perm[k] = p;
    if (p != k) {
      for (int c = 0; c < k; c++) { const double s = LM_A(k, c); LM_A(k, c) = LM_A(p, c); LM_A(p, c) = s; }
      for (int r = p + 1; r < n; r++) { const double s = LM_A(r, k); LM_A(r, k) = LM_A(r, p); LM_A(r, p) = s; }
      { const double s = LM_A(k, k); LM_A(k, k) = LM_A(p, p); LM_A(p, p) = s; }
      for (int i = k + 1; i < p; i++) { const double s = LM_A(i, k); LM_A(i, k) = LM_A(p, i); LM_A(p, i) = s; }
    }
    if (k > 0) {
      double s = 0;
      for (int j = 0; j < k; j++) { tmp[j] = LM_A(j, j) * LM_A(k, j); s += LM_A(k, j) * tmp[j]; }
      LM_A(k, k) -= s;
      for (int i = k + 1; i < n; i++) {
        double a = 0;
        for (int j = 0; j < k; j++) a += LM_A(i, j) * tmp[j];
        LM_A(i, k) -= a;
      }
    }
    const double d = LM_A(k, k);
    if (fabs(d) > 0)
      for (int i = k + 1; i < n; i++) LM_A(i, k) /= d;
  }
  for (int i = 0; i < n; i++) x[i] = rhs[i];
  for (int k = 0; k < n; k++) { const double s = x[k]; x[k] = x[perm[k]]; x[perm[k]] = s; }
  for (int i = 0; i < n; i++) { double s = x[i]; for (int j = 0; j < i; j++) s -= LM_A(i, j) * x[j]; x[i] = s; }
  for (int i = 0; i < n; i++) x[i] = fabs(LM_A(i, i)) > 5.562684646268003e-309 ? x[i] / LM_A(i, i) : 0.0;   // 1 / highest()
  for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int j = i + 1; j < n; j++) s -= LM_A(j, i) * x[j]; x[i] = s; }
  for (int k = n - 1; k >= 0; k--) { const double s = x[k]; x[k] = x[perm[k]]; x[perm[k]] = s; }
#undef LM_A
}

// The same solve for a compile-time size, one thread, working arrays in SHARED memory (A: N x N row-major input, lower triangle
// read; W: N x N scratch; both with compile-time offsets after unrolling, so no local memory and no address arithmetic in the
// dependent chain).  Eigen's in-place LDLT is left-looking: when step k looks for the largest |diagonal| of the trailing block none
// of those entries has been updated yet, so the transposition sequence is a selection sort of the ORIGINAL |diagonal| (first
// maximum wins); the factorisation then runs on the permuted matrix without any search.
template <int N>
__device__ __forceinline__ void ldlt_solve_fixed(const double *__restrict__ A, double *__restrict__ W, const double *rhs, double *x) {
  int idx[N];
  double d[N];
#pragma unroll
  for (int i = 0; i < N; i++) { idx[i] = i; d[i] = fabs(A[i * N + i]); }
#pragma unroll
  for (int k = 0; k < N; k++) {
    int p = k;
    double best = d[k];
#pragma unroll
    for (int i = k + 1; i < N; i++) if (d[i] > best) { best = d[i]; p = i; }
    // swap entries k and p (p is dynamic: done with selects so that the arrays stay in registers)
    const double dk = d[k];
    const int ik = idx[k];
    int ip = ik;
#pragma unroll
    for (int i = k + 1; i < N; i++) if (i == p) { ip = idx[i]; idx[i] = ik; d[i] = dk; }
    idx[k] = ip; d[k] = best;
  }
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) { const int r = idx[i], c = idx[j]; W[i * N + j] = r >= c ? A[r * N + c] : A[c * N + r]; }
  double tmp[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = 0;
#pragma unroll
    for (int j = 0; j < k; j++) { tmp[j] = W[j * N + j] * W[k * N + j]; s += W[k * N + j] * tmp[j]; }
    const double dkk = W[k * N + k] - s;
    W[k * N + k] = dkk;
    const bool ok = fabs(dkk) > 0;
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      double a = 0;
#pragma unroll
      for (int j = 0; j < k; j++) a += W[i * N + j] * tmp[j];
      const double v = W[i * N + k] - a;
      W[i * N + k] = ok ? v / dkk : v;
    }
  }
  double y[N];
#pragma unroll
  for (int i = 0; i < N; i++) y[i] = rhs[idx[i]];
#pragma unroll
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int j = 0; j < i; j++) y[i] -= W[i * N + j] * y[j];
  }
#pragma unroll
  for (int i = 0; i < N; i++) y[i] = fabs(W[i * N + i]) > 5.562684646268003e-309 ? y[i] / W[i * N + i] : 0.0;
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
#pragma unroll
    for (int j = i + 1; j < N; j++) y[i] -= W[j * N + i] * y[j];
  }
#pragma unroll
  for (int i = 0; i < N; i++) x[idx[i]] = y[i];
}

}  // namespace lm
