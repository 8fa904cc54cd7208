// k_solve.cu — EnergyFunctional::solveSystemF without IMU (EnergyFunctional.cpp:1029-1184) as ONE single-CTA
// fp64 kernel: assemble HFinal = HA + HL (+HM), b (+bM + HM*delta), damp diag*(1+1e-5), subtract H_sc/(1+1e-5),
// b_sc, Jacobi-precondition with 1/sqrt(diag+10), pivoted LDL^T (the left-looking algorithm with the
// largest-|diagonal| transposition rule of Eigen::LDLT, the solver called at :1147-1148), substitute, undo the
// scaling, and prepare xAd[h*nf+t] = x_h^T adHostF[h+nf*t] + x_t^T adTargetF[h+nf*t] for resubstituteF_MT (:496-524).
// D = 4 + 8*nf <= 132; the matrix lives in shared memory (D*(D+1) doubles).
#include <math.h>

#include "kernels.h"

namespace {

__global__ void __launch_bounds__(256) k_solve(SolveArgs a) {
  extern __shared__ double sm[];
  const int D = a.D, LD = D + 1, tid = threadIdx.x, nth = blockDim.x;
  double *M = sm;                 // [D][LD]
  double *bb = M + (size_t)D * LD;  // [D]
  double *S = bb + D;             // [D]
  double *temp = S + D;           // [D]
  double *y = temp + D;           // [D]
  __shared__ int tr[136];
  __shared__ int s_big;
  __shared__ double s_red[8];
  __shared__ int s_redi[8];
  const double lambda = 1e-5;                       // EnergyFunctional.cpp:1031
  const double sc = (double)(1.0f / (float)(1 + lambda));  // float-typed scalar (:1099)

  // ---- assemble --------------------------------------------------------------------------------
  for (int e = tid; e < D * D; e += nth) {
    const int r = e / D, c = e % D;
    double v = a.HA[e] + a.HL[e];
    if (a.HM) v += a.HM[e];
    if (r == c) v *= (1 + lambda);
    v -= a.Hsc[e] * sc;
    M[r * LD + c] = v;
    if (a.Hfinal) a.Hfinal[e] = v;
  }
  for (int r = tid; r < D; r += nth) {
    double v = a.bA[r] + a.bL[r];
    if (a.HM) {  // bM_top = bM + HM * getStitchedDeltaF() (:1070-1091)
      double s = 0;
      for (int c = 0; c < D; c++) {
        const double dl = c < 4 ? (double)a.cDeltaF[c] : a.wprior[4 + 16 * a.nf + (c - 4)];
        s += a.HM[(size_t)r * D + c] * dl;
      }
      v += a.bM[r] + s;
    }
    v -= a.bsc[r];
    bb[r] = v;
    if (a.bfinal) a.bfinal[r] = v;
  }
  __syncthreads();
  // ---- Jacobi preconditioning (:1143-1146) -------------------------------------------------------
  for (int i = tid; i < D; i += nth) S[i] = 1.0 / sqrt(M[i * LD + i] + 10.0);
  __syncthreads();
  for (int e = tid; e < D * D; e += nth) { const int r = e / D, c = e % D; M[r * LD + c] = S[r] * M[r * LD + c] * S[c]; }
  for (int i = tid; i < D; i += nth) y[i] = S[i] * bb[i];
  __syncthreads();

  // ---- pivoted LDL^T, lower triangle in place ------------------------------------------------------
  const int warp = tid >> 5, lane = tid & 31, nwarps = nth >> 5;
  for (int k = 0; k < D; k++) {
    // largest |diagonal| of the trailing block, first index on ties
    {
      double best = -1.0; int bi = D;
      for (int i = k + tid; i < D; i += nth) { const double v = fabs(M[i * LD + i]); if (v > best) { best = v; bi = i; } }
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) { s_red[warp] = best; s_redi[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        double bv = s_red[0]; int bidx = s_redi[0];
        for (int w = 1; w < nwarps; w++) if (s_red[w] > bv || (s_red[w] == bv && s_redi[w] < bidx)) { bv = s_red[w]; bidx = s_redi[w]; }
        s_big = bidx; tr[k] = bidx;
      }
      __syncthreads();
    }
    const int big = s_big;
    if (big != k) {  // symmetric transposition on the lower triangle
      for (int c = tid; c < k; c += nth) { const double t = M[k * LD + c]; M[k * LD + c] = M[big * LD + c]; M[big * LD + c] = t; }
      for (int r = big + 1 + tid; r < D; r += nth) { const double t = M[r * LD + k]; M[r * LD + k] = M[r * LD + big]; M[r * LD + big] = t; }
      for (int i = k + 1 + tid; i < big; i += nth) { const double t = M[i * LD + k]; M[i * LD + k] = M[big * LD + i]; M[big * LD + i] = t; }
      if (tid == 0) { const double t = M[k * LD + k]; M[k * LD + k] = M[big * LD + big]; M[big * LD + big] = t; }
      __syncthreads();
    }
    if (k > 0) {
      for (int j = tid; j < k; j += nth) temp[j] = M[j * LD + j] * M[k * LD + j];
      __syncthreads();
      // rows k..D-1: 4 lanes per row share the dot product over j < k (row k itself updates the pivot)
      const int sub = tid & 3;
      for (int base = k; base < D; base += (nth >> 2)) {  // uniform trip count: the shuffles need the whole warp
        const int row = base + (tid >> 2);
        double s = 0;
        if (row < D) for (int j = sub; j < k; j += 4) s += M[row * LD + j] * temp[j];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (row < D && sub == 0) M[row * LD + k] -= s;
      }
      __syncthreads();
    }
    const double akk = M[k * LD + k];
    if (fabs(akk) > 0) for (int i = k + 1 + tid; i < D; i += nth) M[i * LD + k] /= akk;
    __syncthreads();
  }
  // ---- solve: P b, L, D, L^T, P^T ------------------------------------------------------------------
  if (tid == 0) {
    for (int k = 0; k < D; k++) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
  }
  __syncthreads();
  for (int j = 0; j < D; j++) {  // forward, column oriented
    const double yj = y[j];
    for (int i = j + 1 + tid; i < D; i += nth) y[i] -= M[i * LD + j] * yj;
    __syncthreads();
  }
  for (int i = tid; i < D; i += nth) { const double d = M[i * LD + i]; y[i] = fabs(d) > 1.0 / 1.7976931348623157e308 ? y[i] / d : 0.0; }
  __syncthreads();
  for (int j = D - 1; j >= 0; j--) {  // backward with L^T
    const double yj = y[j];
    for (int i = tid; i < j; i += nth) y[i] -= M[j * LD + i] * yj;
    __syncthreads();
  }
  if (tid == 0) {
    for (int k = D - 1; k >= 0; k--) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
  }
  __syncthreads();
  bool bad = false;
  for (int i = tid; i < D; i += nth) { const double xi = S[i] * y[i]; a.x[i] = xi; y[i] = xi; if (!isfinite(xi)) bad = true; }
  if (bad && a.status) a.status[0] = 1;
  __syncthreads();
  // ---- xAd (EnergyFunctional.cpp:509-513) and xc ------------------------------------------------------
  if (a.xAd) {
    const int nf = a.nf;
    for (int e = tid; e < nf * nf * 8; e += nth) {
      const int c = e & 7, ht = e >> 3, h = ht / nf, t = ht % nf;   // xAd index = h*nf + t
      const float *AhF = a.adHostF + 64 * (size_t)(h + nf * t), *AtF = a.adTargetF + 64 * (size_t)(h + nf * t);
      float sh = 0.f, st = 0.f;
      for (int k = 0; k < 8; k++) { sh += (float)y[4 + 8 * h + k] * AhF[k * 8 + c]; st += (float)y[4 + 8 * t + k] * AtF[k * 8 + c]; }
      a.xAd[e] = sh + st;
    }
    if (tid < 4) a.xAd[(size_t)nf * nf * 8 + tid] = (float)y[tid];
  }
}

// resubstitute with a caller-provided x: only the xAd part of the kernel above
__global__ void __launch_bounds__(256) k_make_xad(const double *__restrict__ x, int nf, const float *__restrict__ adHostF,
                                                  const float *__restrict__ adTargetF, float *__restrict__ xAd) {
  const int tid = threadIdx.x;
  for (int e = tid; e < nf * nf * 8; e += blockDim.x) {
    const int c = e & 7, ht = e >> 3, h = ht / nf, t = ht % nf;
    const float *AhF = adHostF + 64 * (size_t)(h + nf * t), *AtF = adTargetF + 64 * (size_t)(h + nf * t);
    float sh = 0.f, st = 0.f;
    for (int k = 0; k < 8; k++) { sh += (float)x[4 + 8 * h + k] * AhF[k * 8 + c]; st += (float)x[4 + 8 * t + k] * AtF[k * 8 + c]; }
    xAd[e] = sh + st;
  }
  if (tid < 4) xAd[(size_t)nf * nf * 8 + tid] = (float)x[tid];
}

}  // namespace

size_t solve_smem_bytes(int D) { return ((size_t)D * (D + 1) + 4 * (size_t)D) * sizeof(double); }

void launch_solve(sosba *h, const SolveArgs &a) {
  const size_t smem = solve_smem_bytes(a.D);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  k_solve<<<1, 256, smem, h->stream>>>(a);
  h->launches++;
}

void launch_make_xad(sosba *h, const double *d_x, int nf, const float *adHostF, const float *adTargetF, float *xAd) {
  k_make_xad<<<1, 256, 0, h->stream>>>(d_x, nf, adHostF, adTargetF, xAd);
  h->launches++;
}
