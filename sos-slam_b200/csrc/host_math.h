// host_math.h — fp64 rigid-body algebra for the host side of the boundary (and, being __host__ __device__,
// for the single-CTA frame-step kernel).  Mirrors the semantics of Sophus 0.9a as the reference uses them
// (thirdparty/Sophus/sophus/se3.hpp: tangent = [translation, rotation]; exp :417-439, Adj :131-139,
// inverse :168-173) on plain rotation matrices.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SOSBA_HD __host__ __device__ inline
#else
#define SOSBA_HD inline
#endif

namespace sosba_math {

struct Rigid {     // x_out = R x + t, R row-major
  double R[9];
  double t[3];
};

SOSBA_HD void mat3_mul(const double *A, const double *B, double *C) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
SOSBA_HD void mat3_vec(const double *A, const double *x, double *y) {
  for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
SOSBA_HD Rigid rigid_identity() {
  Rigid r;
  for (int i = 0; i < 9; i++) r.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  r.t[0] = r.t[1] = r.t[2] = 0.0;
  return r;
}
SOSBA_HD Rigid rigid_mul(const Rigid &a, const Rigid &b) {
  Rigid r;
  mat3_mul(a.R, b.R, r.R);
  double rt[3];
  mat3_vec(a.R, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
  return r;
}
SOSBA_HD Rigid rigid_inverse(const Rigid &a) {
  Rigid r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.R[3 * i + j] = a.R[3 * j + i];
  double rt[3];
  mat3_vec(r.R, a.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = -rt[i];
  return r;
}
SOSBA_HD Rigid rigid_from34(const double *p) {
  Rigid r;
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) r.R[3 * i + j] = p[4 * i + j]; r.t[i] = p[4 * i + 3]; }
  return r;
}
SOSBA_HD void rigid_to34(const Rigid &r, double *p) {
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) p[4 * i + j] = r.R[3 * i + j]; p[4 * i + 3] = r.t[i]; }
}
SOSBA_HD void hat3(const double *w, double *O) {
  O[0] = 0; O[1] = -w[2]; O[2] = w[1];
  O[3] = w[2]; O[4] = 0; O[5] = -w[0];
  O[6] = -w[1]; O[7] = w[0]; O[8] = 0;
}
// SE3::exp of [upsilon, omega]
SOSBA_HD Rigid rigid_exp(const double *xi) {
  const double *u = xi, *w = xi + 3;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double O[9], O2[9];
  hat3(w, O);
  mat3_mul(O, O, O2);
  double A, B, C;  // R = I + A*O + B*O2 ; V = I + B*O + C*O2
  if (th < 1e-10) {
    A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0;
  } else {
    A = sin(th) / th; B = (1.0 - cos(th)) / th2; C = (th - sin(th)) / (th2 * th);
  }
  Rigid r;
  double V[9];
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    r.R[i] = I + A * O[i] + B * O2[i];
    V[i] = I + B * O[i] + C * O2[i];
  }
  mat3_vec(V, u, r.t);
  return r;
}
// 6x6 row-major adjoint [[R, hat(t) R], [0, R]]
SOSBA_HD void rigid_adj(const Rigid &T, double *A) {
  double O[9], tR[9];
  hat3(T.t, O);
  mat3_mul(O, T.R, tR);
  for (int i = 0; i < 36; i++) A[i] = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { A[6 * i + j] = T.R[3 * i + j]; A[6 * (i + 3) + j + 3] = T.R[3 * i + j]; A[6 * i + j + 3] = tR[3 * i + j]; }
}

}  // namespace sosba_math
