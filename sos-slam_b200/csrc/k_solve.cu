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

// final (undamped) top value at (r,c), r >= c, from the raw stitch (AccumulatedTopHessian.h:107-126 epilogue)
__device__ __forceinline__ double top_entry(const double *__restrict__ Hraw, int D, int r, int c) {
  if (c >= 4 && ((r - 4) >> 3) != ((c - 4) >> 3)) return Hraw[(size_t)r * D + c] + Hraw[(size_t)c * D + r];
  return Hraw[(size_t)r * D + c];
}

// 256 threads as a 16x16 grid; thread (ty,tx) keeps the entries {(ty+16a, tx+16b)} of the augmented, permuted,
// preconditioned lower triangle in registers for the whole factorisation (T = ceil((D+1)/16)).
// Pivot order: Eigen::LDLT is left-looking, so when it searches the largest |diagonal| of the trailing block none
// of those entries has been updated yet — the transposition sequence depends only on the input diagonal and equals
// a descending sort of |diag| (ties broken by index).  The factorisation itself then needs no pivot search and can
// run right-looking with one barrier per column; the right-hand side rides along as row D, ending as D^-1 L^-1 P b.
template <int T>
__global__ void __launch_bounds__(256) k_solve(SolveArgs a) {
  extern __shared__ double sm[];
  const int D = a.D, LD = D + 1, tid = threadIdx.x, nf = a.nf;
  double *M = sm;                          // [D][LD] unit lower factor, for the back substitution
  double *S = M + (size_t)D * LD;          // [D] Jacobi scaling
  double *bb = S + D;                      // [D] unscaled rhs
  double *dtmp = bb + D;                   // [D] damped diagonal
  double *delta = dtmp + D;                // [D]
  double *key = delta + D;                 // [D]
  double *col = key + D;                   // [2][D+2] column broadcast, double buffered
  double *z = col + 2 * (D + 2);           // [D]
  __shared__ int perm[144];
  __shared__ double s_dk[2];
  const double lambda = 1e-5;                               // EnergyFunctional.cpp:1031
  const double sc = (double)(1.0f / (float)(1 + lambda));   // float-typed scalar (:1099)
  const double *cPrior = a.wprior, *fprior = a.wprior + 4, *fdp = a.wprior + 4 + 8 * nf, *fdelta = a.wprior + 4 + 16 * nf;
  const int DP = D + 1;
  const int warp = tid >> 5, lane = tid & 31;
  const int ty = tid >> 4, tx = tid & 15;
#define SOLVE_TS(n) do { if (a.dbg && tid == 0) a.dbg[n] = clock64(); } while (0)
  SOLVE_TS(0);

  // ---- diagonal, rhs --------------------------------------------------------------------------------
  for (int i = tid; i < D; i += 256) delta[i] = i < 4 ? (double)a.cDeltaF[i] : fdelta[i - 4];
  __syncthreads();
  double *hmd = key;   // HM * delta, bM_top = bM + HM * delta (:1070-1091)
  if (a.HM) {
    for (int r = warp; r < D; r += 8) {
      double s = 0;
      for (int c = lane; c < D; c += 32) s += a.HM[(size_t)r * D + c] * delta[c];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) hmd[r] = s;
    }
    __syncthreads();
  }
  if (tid < D) {
    const int r = tid;
    const double pr = r < 4 ? cPrior[r] : fprior[r - 4];
    const double dpr = r < 4 ? (double)a.cDeltaF[r] : fdp[r - 4];
    const double bt = a.btop[r], scb = a.accSC[(size_t)r * DP + D], ht = a.Htop[(size_t)r * D + r], scd = a.accSC[(size_t)r * DP + r];
    double v = bt + pr * dpr;                            // AccumulatedTopHessian.cpp:292-300 (the L pass carries the priors)
    double dg = ht + pr;
    if (a.HM) { v += a.bM[r] + hmd[r]; dg += a.HM[(size_t)r * D + r]; }
    v -= scb;
    bb[r] = v;
    if (a.bfinal) a.bfinal[r] = v;
    dg *= (1 + lambda);
    dg -= scd * sc;
    dtmp[r] = dg;
    const double sr = 1.0 / sqrt(dg + 10.0);             // :1143-1146
    S[r] = sr;
  }
  __syncthreads();
  if (tid < D) key[tid] = fabs(S[tid] * dtmp[tid] * S[tid]);
  __syncthreads();
  SOLVE_TS(1);
  // ---- pivot order = descending |diag| ---------------------------------------------------------------------
  if (tid < D) {
    const double mine = key[tid];
    int rank = 0;
    for (int j = 0; j < D; j++) { const double o = key[j]; rank += (o > mine || (o == mine && j < tid)) ? 1 : 0; }
    perm[rank] = tid;
  }
  __syncthreads();
  SOLVE_TS(2);
  // ---- assemble into registers ---------------------------------------------------------------------------
  double reg[T][T];
#pragma unroll
  for (int ia = 0; ia < T; ia++) {
    // one register-tile row at a time: indices, then all global loads of the row in flight, then the arithmetic
    int er[T], ec[T];
    double h1[T], h2[T], hs[T], hm[T];
    const int i = ty + 16 * ia;
#pragma unroll
    for (int jb = 0; jb <= ia; jb++) {
      const int j = tx + 16 * jb;
      const bool ok = i < D && j <= i;
      int r = perm[ok ? i : 0], c = perm[ok ? j : 0];
      if (r < c) { const int t = r; r = c; c = t; }
      er[jb] = r; ec[jb] = c;
    }
#pragma unroll
    for (int jb = 0; jb <= ia; jb++) {
      const int r = er[jb], c = ec[jb];
      h1[jb] = a.Htop[(size_t)r * D + c];
      h2[jb] = a.Htop[(size_t)c * D + r];
      hs[jb] = a.accSC[(size_t)c * DP + r];
      hm[jb] = a.HM ? a.HM[(size_t)r * D + c] : 0.0;
    }
#pragma unroll
    for (int jb = 0; jb < T; jb++) {
      double v = 0.0;
      if (jb <= ia) {
        const int j = tx + 16 * jb;
        if (i == D && j < D) {
          const int c = perm[j];
          v = S[c] * bb[c];
        } else if (i < D && j <= i) {
          const int r = er[jb], c = ec[jb];
          double u;
          if (r == c) u = dtmp[r];
          else {
            const bool offdiag = c >= 4 && ((r - 4) >> 3) != ((c - 4) >> 3);   // AccumulatedTopHessian.h:107-126 epilogue
            u = (offdiag ? h1[jb] + h2[jb] : h1[jb]) + hm[jb];
            u -= hs[jb] * sc;
          }
          if (a.Hfinal) { a.Hfinal[(size_t)r * D + c] = u; a.Hfinal[(size_t)c * D + r] = u; }
          v = S[r] * u * S[c];
        }
      }
      reg[ia][jb] = v;
    }
  }
  SOLVE_TS(3);
  // ---- right-looking LDL^T, one barrier per column ------------------------------------------------------------
#pragma unroll
  for (int kb = 0; kb < T; kb++) {     // column block: static, so every register index below is a constant
    for (int km = 0; km < 16; km++) {
      const int k = 16 * kb + km;
      if (k >= D) break;
      double *cb = col + (k & 1) * (D + 2);
      if (tx == km) {
#pragma unroll
        for (int ia = kb; ia < T; ia++) {
          const int i = ty + 16 * ia;
          const double v = reg[ia][kb];
          if (i > k && i <= D) cb[i] = v;
          if (i == k) s_dk[k & 1] = __drcp_rn(v);
        }
      }
      __syncthreads();
      const double rcp0 = s_dk[k & 1];
      const double rcp = isfinite(rcp0) ? rcp0 : 0.0;   // zero pivot: leave the column (Eigen: pivot_is_valid)
      double li[T], cj[T];
#pragma unroll
      for (int ia = kb; ia < T; ia++) { const int i = ty + 16 * ia; li[ia] = (i > k && i <= D) ? cb[i] * rcp : 0.0; }
#pragma unroll
      for (int jb = kb; jb < T; jb++) { const int j = tx + 16 * jb; cj[jb] = (j > k && j < D) ? cb[j] : 0.0; }
#pragma unroll
      for (int ia = kb; ia < T; ia++)
#pragma unroll
        for (int jb = kb; jb <= ia; jb++) reg[ia][jb] -= li[ia] * cj[jb];   // entries outside the trailing block see li or cj == 0
      if (tx == km) {
#pragma unroll
        for (int ia = kb; ia < T; ia++)
          if (ty + 16 * ia > k) reg[ia][kb] = li[ia];   // keep the unit-lower factor in place
      }
    }
  }
  SOLVE_TS(4);
  // ---- spill the factor, back substitution L^T w = z in one warp ----------------------------------------------
#pragma unroll
  for (int ia = 0; ia < T; ia++)
#pragma unroll
    for (int jb = 0; jb < T; jb++) {
      const int i = ty + 16 * ia, j = tx + 16 * jb;
      if (i < D && j < i) M[i * LD + j] = reg[ia][jb];
      if (i == D && j < D) z[j] = reg[ia][jb];
    }
  __syncthreads();
  if (warp == 0) {
    constexpr int W = (16 * T + 31) / 32;
    double w[W];
#pragma unroll
    for (int m = 0; m < W; m++) { const int i = lane + 32 * m; w[m] = i < D ? z[i] : 0.0; }
    for (int j = D - 1; j > 0; j--) {
      const int jm = j >> 5, jl = j & 31;
      double own = 0.0;
#pragma unroll
      for (int m = 0; m < W; m++) own = m == jm ? w[m] : own;
      const double wj = __shfl_sync(0xffffffffu, own, jl);
      const double *Lj = M + j * LD;
#pragma unroll
      for (int m = 0; m < W; m++) { const int i = lane + 32 * m; if (i < j) w[m] -= Lj[i] * wj; }
    }
#pragma unroll
    for (int m = 0; m < W; m++) { const int i = lane + 32 * m; if (i < D) z[i] = w[m]; }
  }
  __syncthreads();
  SOLVE_TS(5);
  double *y = key;   // x in original order
  bool bad = false;
  for (int j = tid; j < D; j += 256) { const int r = perm[j]; const double xi = S[r] * z[j]; y[r] = xi; a.x[r] = xi; if (!isfinite(xi)) bad = true; }
  if (bad && a.status) a.status[0] = 1;
  __syncthreads();
  SOLVE_TS(6);
  // ---- xAd (EnergyFunctional.cpp:509-513) and xc ------------------------------------------------------
  if (a.xAd) {
    for (int e = tid; e < nf * nf * 8; e += 256) {
      const int c = e & 7, ht = e >> 3, h = ht / nf, t = ht % nf;   // xAd index = h*nf + t
      const float *AhF = a.adHostF + 64 * (size_t)(h + nf * t), *AtF = a.adTargetF + 64 * (size_t)(h + nf * t);
      float ah[8], at[8];
#pragma unroll
      for (int k = 0; k < 8; k++) { ah[k] = __ldg(AhF + k * 8 + c); at[k] = __ldg(AtF + k * 8 + c); }
      float sh = 0.f, st = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) { sh += (float)y[4 + 8 * h + k] * ah[k]; st += (float)y[4 + 8 * t + k] * at[k]; }
      a.xAd[e] = sh + st;
    }
    if (tid < 4) a.xAd[(size_t)nf * nf * 8 + tid] = (float)y[tid];
  }
  SOLVE_TS(7);
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

size_t solve_smem_bytes(int D) { return ((size_t)D * (D + 1) + 6 * (size_t)D + 2 * (size_t)(D + 2) + 8) * sizeof(double); }

void launch_solve(sosba *h, const SolveArgs &a) {
  const size_t smem = solve_smem_bytes(a.D);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(k_solve<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_solve<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_solve<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  const int T = (a.D + 1 + 15) / 16;
  if (T <= 5) k_solve<5><<<1, 256, smem, h->stream>>>(a);
  else if (T <= 7) k_solve<7><<<1, 256, smem, h->stream>>>(a);
  else k_solve<9><<<1, 256, smem, h->stream>>>(a);
  h->launches++;
}

void launch_make_xad(sosba *h, const double *d_x, int nf, const float *adHostF, const float *adTargetF, float *xAd) {
  k_make_xad<<<1, 256, 0, h->stream>>>(d_x, nf, adHostF, adTargetF, xAd);
  h->launches++;
}
