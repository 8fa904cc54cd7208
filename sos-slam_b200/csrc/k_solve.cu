// k_solve.cu — EnergyFunctional::solveSystemF without IMU (EnergyFunctional.cpp:1029-1184) as ONE single-CTA
// fp64 kernel.  Input is the raw stitch of the top blocks (A and L passes summed, not yet symmetrised), the
// stitched-space Schur Gram matrix, the priors and (optionally) the marginalisation prior HM, bM:
//   HFinal = sym(Htop) + priors (+HM), b = btop + prior*delta_prior (+bM + HM*delta); diag *= (1+1e-5);
//   HFinal -= H_sc/(1+1e-5); b -= b_sc; Jacobi scaling 1/sqrt(diag+10); LDL^T with the transposition order of
//   Eigen::LDLT (the solver called at :1147-1148); back-substitution; x = S * y; then
//   xAd[h*nf+t] = x_h^T adHostF[h+nf*t] + x_t^T adTargetF[h+nf*t] for resubstituteF_MT (:496-524).
// D = 4 + 8*nf <= 132; the (D+1)^2 augmented matrix lives in shared memory.
#include <math.h>

#include "kernels.h"

namespace {

// Pivot order.  Eigen::LDLT is left-looking: when it searches the largest |diagonal| of the trailing block at
// step k, none of those entries has been touched yet, so the transposition sequence depends only on the
// diagonal of the (preconditioned) input.  One warp replays it on a copy of the diagonal; perm[i] = original
// index that ends up at position i.  The factorisation itself can then run without pivoting, in any order.
__device__ void pivot_order(const double *__restrict__ diag_in, double *__restrict__ d, int *__restrict__ perm, int D, int lane) {
  for (int i = lane; i < D; i += 32) { d[i] = fabs(diag_in[i]); perm[i] = i; }
  __syncwarp();
  for (int k = 0; k < D; k++) {
    double best = -1.0; int bi = D;
    for (int i = k + lane; i < D; i += 32) { const double v = d[i]; if (v > best) { best = v; bi = i; } }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0 && bi != k) { const double t = d[k]; d[k] = d[bi]; d[bi] = t; const int p = perm[k]; perm[k] = perm[bi]; perm[bi] = p; }
    __syncwarp();
  }
}

// final (undamped) top value at (r,c), r >= c, from the raw stitch (AccumulatedTopHessian.h:107-126 epilogue)
__device__ __forceinline__ double top_entry(const double *__restrict__ Hraw, int D, int r, int c) {
  if (c >= 4 && ((r - 4) >> 3) != ((c - 4) >> 3)) return Hraw[(size_t)r * D + c] + Hraw[(size_t)c * D + r];
  return Hraw[(size_t)r * D + c];
}

__global__ void __launch_bounds__(256) k_solve(SolveArgs a) {
  extern __shared__ double sm[];
  const int D = a.D, LD = D + 1, tid = threadIdx.x, nth = blockDim.x, nf = a.nf;
  double *M = sm;                          // [D+1][LD]: permuted, preconditioned lower triangle + rhs row
  double *S = M + (size_t)(D + 1) * LD;    // [D] Jacobi scaling
  double *bb = S + D;                      // [D] unscaled rhs
  double *cbuf = bb + D;                   // [D+1]
  double *lbuf = cbuf + D + 1;             // [D+1]
  double *dtmp = lbuf + D + 1;             // [D]
  double *delta = dtmp + D;                // [D]
  __shared__ int perm[136];
  const double lambda = 1e-5;                               // EnergyFunctional.cpp:1031
  const double sc = (double)(1.0f / (float)(1 + lambda));   // float-typed scalar (:1099)
  const double *cPrior = a.wprior, *fprior = a.wprior + 4, *fdp = a.wprior + 4 + 8 * nf, *fdelta = a.wprior + 4 + 16 * nf;
  const int DP = D + 1;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nth >> 5;

  // ---- diagonal, rhs --------------------------------------------------------------------------------
  for (int i = tid; i < D; i += nth) delta[i] = i < 4 ? (double)a.cDeltaF[i] : fdelta[i - 4];
  __syncthreads();
  for (int r = warp; r < D; r += nwarps) {   // bM_top = bM + HM * delta (:1070-1091)
    double s = 0;
    if (a.HM) for (int c = lane; c < D; c += 32) s += a.HM[(size_t)r * D + c] * delta[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const double pr = r < 4 ? cPrior[r] : fprior[r - 4];
      const double dpr = r < 4 ? (double)a.cDeltaF[r] : fdp[r - 4];
      double v = a.btop[r] + pr * dpr;                     // AccumulatedTopHessian.cpp:292-300 (L pass carries the priors)
      if (a.HM) v += a.bM[r] + s;
      v -= a.accSC[(size_t)r * DP + D];
      bb[r] = v;
      if (a.bfinal) a.bfinal[r] = v;
      double dg = a.Htop[(size_t)r * D + r] + pr;
      if (a.HM) dg += a.HM[(size_t)r * D + r];
      dg *= (1 + lambda);
      dg -= a.accSC[(size_t)r * DP + r] * sc;
      dtmp[r] = dg;
      S[r] = 1.0 / sqrt(dg + 10.0);                        // :1143-1146
    }
  }
  __syncthreads();
  if (warp == 0) {
    for (int i = lane; i < D; i += 32) cbuf[i] = S[i] * dtmp[i] * S[i];
    __syncwarp();
    pivot_order(cbuf, lbuf, perm, D, lane);
  }
  __syncthreads();
  // ---- assemble the permuted, preconditioned lower triangle --------------------------------------------
  for (int e = tid; e < D * D; e += nth) {
    const int i = e / D, j = e % D;
    if (j > i) continue;
    int r = perm[i], c = perm[j];
    if (r < c) { const int t = r; r = c; c = t; }
    double v;
    if (r == c) v = dtmp[r];
    else {
      v = top_entry(a.Htop, D, r, c);
      if (a.HM) v += a.HM[(size_t)r * D + c];
      v -= a.accSC[(size_t)c * DP + r] * sc;
    }
    if (a.Hfinal) { a.Hfinal[(size_t)r * D + c] = v; a.Hfinal[(size_t)c * D + r] = v; }
    M[i * LD + j] = S[r] * v * S[c];
  }
  for (int j = tid; j < D; j += nth) M[D * LD + j] = S[perm[j]] * bb[perm[j]];
  __syncthreads();

  // ---- right-looking LDL^T with the rhs as an extra row: row D ends as D^-1 L^-1 P b ----------------------
  const int ty = tid >> 4, tx = tid & 15;
  for (int k = 0; k < D; k++) {
    const double dk = M[k * LD + k];
    for (int i = k + 1 + tid; i <= D; i += nth) {
      const double v = M[i * LD + k];
      cbuf[i] = v;
      const double l = dk != 0.0 ? v / dk : 0.0;
      lbuf[i] = l;
      M[i * LD + k] = l;
    }
    __syncthreads();
    for (int i = k + 1 + ty; i <= D; i += 16) {
      const double li = lbuf[i];
      const int jmax = i < D ? i : D - 1;
      for (int j = k + 1 + tx; j <= jmax; j += 16) M[i * LD + j] -= li * cbuf[j];
    }
    __syncthreads();
  }
  // ---- L^T w = z, undo permutation and scaling -----------------------------------------------------------
  double *w = cbuf;
  for (int j = tid; j < D; j += nth) w[j] = M[D * LD + j];
  __syncthreads();
  for (int j = D - 1; j > 0; j--) {
    const double wj = w[j];
    for (int i = tid; i < j; i += nth) w[i] -= M[j * LD + i] * wj;
    __syncthreads();
  }
  double *y = lbuf;   // x in original order
  bool bad = false;
  for (int j = tid; j < D; j += nth) { const int r = perm[j]; const double xi = S[r] * w[j]; y[r] = xi; a.x[r] = xi; if (!isfinite(xi)) bad = true; }
  if (bad && a.status) a.status[0] = 1;
  __syncthreads();
  // ---- xAd (EnergyFunctional.cpp:509-513) and xc ------------------------------------------------------
  if (a.xAd) {
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

size_t solve_smem_bytes(int D) { return ((size_t)(D + 1) * (D + 1) + 6 * (size_t)D + 8) * sizeof(double); }

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
