// k_accum.cu — block-Hessian accumulation, Schur complement and back-substitution.
//
//   top_accumulate   AccumulatedTopHessianSSE::addPoint<mode> block part + AccumulatorApprox::update /
//                    updateTopRight / updateBotRight     AccumulatedTopHessian.cpp:35-147, MatrixAccumulators.h:928-1112
//   point_sums       the per-point Hdd/bd/Hcd sums of addPoint<mode>            AccumulatedTopHessian.cpp:124-146
//   sc_accumulate    AccumulatedSCHessianSSE::addPoint                          AccumulatedSCHessian.cpp:32-79
//   stitch_top       AccumulatedTopHessianSSE::stitchDoubleInternal + MT epilogue  AccumulatedTopHessian.cpp:231-301, .h:80-127
//   finalize_sc      AccumulatedSCHessianSSE::stitchDoubleInternal + MT epilogue   AccumulatedSCHessian.cpp:80-158, .h:88-124
//   resubstitute     EnergyFunctional::resubstituteFPt                          EnergyFunctional.cpp:526-551
//
// Re-design notes: the reference keeps 6 thread-private copies of nf^2 13x13 float blocks with
// 3-tier float buffers and sums them in the stitch.  Here residuals are visited in (host,target)-block order by
// warps that own the 91 upper-triangle entries of one block in registers (3 per lane), so a block change is the
// only time anything leaves the SM: one fp64 red.global per entry into the nf^2 x 92 fp64 block table.  The
// Schur term is accumulated directly in stitched space: every point contributes w * g g^T with
// g = [Hcd ; sum_r adHost*JpJdF (host rows) ; adTarget*JpJdF (target rows) ; bdSum], a (D+1)^2 rank-1 update
// done as a register-tiled SYRK over 32-point tiles in shared memory — nf^3 8x8 blocks never exist.
#include <math.h>

#include "energy_th.cuh"
#include "kernels.h"
#include "resub.cuh"
#include "top_entries.cuh"

namespace {

__device__ __forceinline__ float entry_value(const float *s, const EntryDesc &d) {
  if (d.kind == 0) {
    const float xr = s[CR_X + d.p], yr = s[CR_Y + d.p], xc = s[CR_X + d.q], yc = s[CR_Y + d.q];
    const float a = s[CR_A], b = s[CR_A + 1], c = s[CR_A + 2];
    return a * xc * xr + c * yc * yr + b * (xc * yr + yc * xr);
  }
  if (d.kind == 1) return s[CR_X + d.p] * s[CR_TR + 2 * d.q] + s[CR_Y + d.p] * s[CR_TR + 2 * d.q + 1];
  if (d.kind == 2) return s[CR_BR + d.p];
  return 0.f;
}

// The 91 entries of one block over the 32 lanes without divergence: every lane runs the same four expressions.
//   q0: 10x10 entry `lane` (0..31)   q1: 10x10 entry 32 + lane (lane < 23)   tr: TopRight entry lane (< 30)   br: BotRight entry lane (< 6)
// 10x10 entry (r,c) = x_r (a x_c + b y_c) + y_r (b x_c + c y_c)   (AccumulatorApprox::update, MatrixAccumulators.h:928-1050)
struct LaneEntries {
  int r0, c0, r1, c1, trp, trq, brk;   // clamped to valid offsets for idle lanes
  bool v1, vtr, vbr;
};
__device__ __forceinline__ LaneEntries lane_entries(int lane) {
  LaneEntries L;
  EntryDesc d = entry_desc(lane);
  L.r0 = d.p; L.c0 = d.q;
  L.v1 = lane < 23;
  d = entry_desc(L.v1 ? lane + 32 : 32);
  L.r1 = d.p; L.c1 = d.q;
  L.vtr = lane < 30;
  L.trp = L.vtr ? lane / 3 : 0; L.trq = L.vtr ? lane % 3 : 0;
  L.vbr = lane < 6;
  L.brk = L.vbr ? lane : 0;
  return L;
}
struct LaneAcc { float q0, q1, tr, br; };
__device__ __forceinline__ void lane_accumulate(const float *__restrict__ s, const LaneEntries &L, LaneAcc &A) {
  const float4 abc = *(const float4 *)(s + CR_A);   // a, b, c, (TR[0])
  const float xr0 = s[CR_X + L.r0], yr0 = s[CR_Y + L.r0], xc0 = s[CR_X + L.c0], yc0 = s[CR_Y + L.c0];
  const float xr1 = s[CR_X + L.r1], yr1 = s[CR_Y + L.r1], xc1 = s[CR_X + L.c1], yc1 = s[CR_Y + L.c1];
  const float xp = s[CR_X + L.trp], yp = s[CR_Y + L.trp], t0 = s[CR_TR + 2 * L.trq], t1 = s[CR_TR + 2 * L.trq + 1];
  const float bv = s[CR_BR + L.brk];
  A.q0 += xr0 * (abc.x * xc0 + abc.y * yc0) + yr0 * (abc.y * xc0 + abc.z * yc0);
  A.q1 += xr1 * (abc.x * xc1 + abc.y * yc1) + yr1 * (abc.y * xc1 + abc.z * yc1);
  A.tr += xp * t0 + yp * t1;
  A.br += bv;
}
__device__ __forceinline__ void lane_flush(double *__restrict__ dst, int lane, const LaneEntries &L, const LaneAcc &A, int n) {
  atomicAdd(dst + lane, (double)A.q0);
  if (L.v1) atomicAdd(dst + 32 + lane, (double)A.q1);
  if (L.vtr) atomicAdd(dst + 55 + lane, (double)A.tr);
  if (L.vbr) atomicAdd(dst + 85 + lane, (double)A.br);
  if (lane == 31) atomicAdd(dst + 91, (double)n);
}

// One warp walks a contiguous chunk of the block-ordered residual list.
constexpr int TOP_WARPS = 8;
__global__ void __launch_bounds__(TOP_WARPS * 32) k_top_accumulate(AccArgs a, int chunk) {
  __shared__ __align__(16) float s_rec[TOP_WARPS][SOSBA_CREC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * TOP_WARPS + warp;
  const int k0 = gw * chunk, k1 = min(k0 + chunk, a.n_list);
  if (k0 >= k1) return;
  const LaneEntries LE = lane_entries(lane);
  LaneAcc A = {0.f, 0.f, 0.f, 0.f};
  int cur = -1, nacc = 0, ntotal = 0;
  float *s = s_rec[warp];

  auto flush = [&]() {
    if (cur >= 0 && nacc > 0) lane_flush(a.accTop + (size_t)cur * SOSBA_TOPB, lane, LE, A, nacc);
    A.q0 = A.q1 = A.tr = A.br = 0.f; nacc = 0;
  };

  // software prefetch: the record of residual k+1 is in flight while k is accumulated
  auto wanted = [&](int r) -> bool {
    if (a.r_dropped[r] || !a.r_is_active[r]) return false;
    if (a.mode == 0) return !a.r_is_lin[r];
    if (a.mode == 1) return a.r_is_lin[r];
    return true;
  };
  int r = a.list ? a.list[k0] : k0;
  bool use = wanted(r);
  float v0 = 0.f, v1 = 0.f;
  if (use) { const float *rec = a.rec + (size_t)r * SOSBA_CREC; v0 = rec[lane]; if (lane < SOSBA_CREC - 32) v1 = rec[32 + lane]; }
  for (int k = k0; k < k1; k++) {
    const int rc = r; const bool usec = use; const float c0 = v0, c1 = v1;
    if (k + 1 < k1) {
      r = a.list ? a.list[k + 1] : k + 1;
      use = wanted(r);
      if (use) { const float *rec = a.rec + (size_t)r * SOSBA_CREC; v0 = rec[lane]; if (lane < SOSBA_CREC - 32) v1 = rec[32 + lane]; }
    }
    if (!usec) continue;
    const int blk = a.r_host[rc] + a.r_target[rc] * a.nf;
    if (blk != cur) { flush(); cur = blk; }
    __syncwarp();
    s[lane] = c0;
    if (lane < SOSBA_CREC - 32) s[32 + lane] = c1;
    __syncwarp();
    lane_accumulate(s, LE, A);
    nacc++; ntotal++;
  }
  flush();
  if (lane == 0 && ntotal) atomicAdd(a.n_acc, ntotal);
}

// ------------------------------------------------------------------------------------------------
// Per-point sums + Schur complement in stitched space.  CTA = 256 threads, tiles of 32 points.
//   phase 1: 8 lanes per point, lane q owns residual res_begin[p] + q (+8, +16 for nf > 9):
//            (a) its share of Hdd / bd / Hcd (AccumulatedTopHessian.cpp:124-127), summed across the lanes in
//                residual order into the A (non-linearised) and L (linearised) point sums (:132-146);
//            (b) adHost * JpJdF (host rows, summed over the lanes) and adTarget * JpJdF (target rows);
//            then HdiF, bdSumF (AccumulatedSCHessian.cpp:46-56) and g -> shared Gs[32][DPAD]
//   phase 2: register-tiled SYRK acc += w * g g^T; thread owns up to 3 4x4 tiles of the (D+1)^2 upper triangle
constexpr int SC_TP = 32;       // points per tile
constexpr int SC_MAXT = 3;      // 4x4 tiles per thread (nf <= 16)
__global__ void __launch_bounds__(256) k_point_sc(SCArgs a, int DP, int DPAD, int ntiles4, int tiles_total) {
  extern __shared__ float smem[];
  float *Gs = smem;                    // [SC_TP][DPAD]
  float *Ws = smem + SC_TP * DPAD;     // [SC_TP]
  const int tid = threadIdx.x;
  const int nt4 = DPAD / 4;
  int my_ti[SC_MAXT], my_tj[SC_MAXT];
  float acc[SC_MAXT][16];
#pragma unroll
  for (int m = 0; m < SC_MAXT; m++) {
    int t = tid + m * 256;
    my_ti[m] = -1; my_tj[m] = 0;
    if (t < ntiles4) {
      int ti = 0, base = 0;
      while (t >= base + (nt4 - ti)) { base += nt4 - ti; ti++; }
      my_ti[m] = ti; my_tj[m] = ti + (t - base);
    }
#pragma unroll
    for (int q = 0; q < 16; q++) acc[m][q] = 0.f;
  }
  const int np = a.plist ? a.n_plist : a.P;
  const int D = a.D, nf = a.nf;
  const int lp = tid >> 3, sub = tid & 7;
  const unsigned gmask = 0xFFu << ((tid & 31) & ~7);

  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
    for (int i = tid; i < SC_TP * DPAD; i += 256) Gs[i] = 0.f;
    if (tid < SC_TP) Ws[tid] = 0.f;
    __syncthreads();
    {
      const int k = tile * SC_TP + lp;
      if (k < np) {   // uniform across the 8 lanes of a point
        const int p = a.plist ? a.plist[k] : k;
        const int rb = a.res_begin[p], re = a.res_begin[p + 1];
        const int host = a.p_host[p];
        float *g = Gs + lp * DPAD;
        float HddA = 0.f, bdA = 0.f, HcdA[4] = {0.f, 0.f, 0.f, 0.f}, HddL = 0.f, bdL = 0.f, HcdL[4] = {0.f, 0.f, 0.f, 0.f};
        float gh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int ngood = 0;
        for (int base = rb; base < re; base += 8) {   // one round for nf <= 9
          const int r = base + sub;
          const bool use = r < re && a.r_is_active[r] && !a.r_dropped[r];
          const bool lin = use && (a.mode == 2 || a.r_is_lin[r]);
          float c_bd = 0.f, c_Hdd = 0.f, c_Hcd[4] = {0.f, 0.f, 0.f, 0.f};
          float sh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (use) {
            const float4 *rec4 = (const float4 *)(a.rec + (size_t)r * SOSBA_CREC);
            const float4 x03 = rec4[0];                    // CR_X 0..3
            const float4 y_a = rec4[2], y_b = rec4[3];     // floats 8..15: x[8],x[9],y[0],y[1] | y[2],y[3],y[4],y[5]
            const float4 f20 = rec4[5], f24 = rec4[6], f28 = rec4[7], f32 = rec4[8], f36 = rec4[9];
            const float4 v03 = rec4[10], v47 = rec4[11];   // JpJdF
            const float a00 = f20.x, a01 = f20.y, a11 = f20.z;               // CR_A 20..22
            const float JI_r0 = f24.w, JI_r1 = f28.x;                         // CR_TR+4 = 27, 28
            const float Jpdd0 = f32.w, Jpdd1 = f36.x;                         // CR_JPDD = 35, 36
            const float v0 = a00 * Jpdd0 + a01 * Jpdd1, v1 = a01 * Jpdd0 + a11 * Jpdd1;  // JIdx2 * Jpdd
            c_bd = JI_r0 * Jpdd0 + JI_r1 * Jpdd1;
            c_Hdd = v0 * Jpdd0 + v1 * Jpdd1;
            const float xs[4] = {x03.x, x03.y, x03.z, x03.w}, ys[4] = {y_a.z, y_a.w, y_b.x, y_b.y};
#pragma unroll
            for (int i = 0; i < 4; i++) c_Hcd[i] = xs[i] * v0 + ys[i] * v1;
            const int t = a.r_target[r];
            const float4 *Ah = (const float4 *)(a.adHostF + 64 * (size_t)(host + t * nf));
            const float4 *At = (const float4 *)(a.adTargetF + 64 * (size_t)(host + t * nf));
            const float v[8] = {v03.x, v03.y, v03.z, v03.w, v47.x, v47.y, v47.z, v47.w};
#pragma unroll
            for (int row = 0; row < 8; row++) {
              const float4 h0 = __ldg(Ah + 2 * row), h1 = __ldg(Ah + 2 * row + 1), t0 = __ldg(At + 2 * row), t1 = __ldg(At + 2 * row + 1);
              sh[row] = h0.x * v[0] + h0.y * v[1] + h0.z * v[2] + h0.w * v[3] + h1.x * v[4] + h1.y * v[5] + h1.z * v[6] + h1.w * v[7];
              const float st = t0.x * v[0] + t0.y * v[1] + t0.z * v[2] + t0.w * v[3] + t1.x * v[4] + t1.y * v[5] + t1.z * v[6] + t1.w * v[7];
              atomicAdd(&g[4 + 8 * t + row], st);   // one residual per (point, target): uncontended
            }
          }
          ngood += __popc(__ballot_sync(gmask, use) & gmask);
          const int cnt = min(8, re - base);
          for (int q = 0; q < cnt; q++) {   // residual-order sums (deterministic, reference order)
            const bool ql = __shfl_sync(gmask, (int)lin, q, 8) != 0;
            const float q_bd = __shfl_sync(gmask, c_bd, q, 8), q_Hdd = __shfl_sync(gmask, c_Hdd, q, 8);
            if (ql) { bdL += q_bd; HddL += q_Hdd; } else { bdA += q_bd; HddA += q_Hdd; }
#pragma unroll
            for (int i = 0; i < 4; i++) { const float qc = __shfl_sync(gmask, c_Hcd[i], q, 8); if (ql) HcdL[i] += qc; else HcdA[i] += qc; }
          }
#pragma unroll
          for (int row = 0; row < 8; row++) {
            float s = sh[row];
            s += __shfl_xor_sync(gmask, s, 1, 8); s += __shfl_xor_sync(gmask, s, 2, 8); s += __shfl_xor_sync(gmask, s, 4, 8);
            gh[row] += s;
          }
        }
        if (sub == 0) {
          a.HddA[p] = HddA; a.bdA[p] = bdA; a.HddL[p] = HddL; a.bdL[p] = bdL;
          for (int i = 0; i < 4; i++) { a.HcdA[4 * p + i] = HcdA[i]; a.HcdL[4 * p + i] = HcdL[i]; }
        }
        if (ngood == 0) {
          if (sub == 0) { a.HdiF[p] = 0.f; a.bdSumF[p] = 0.f; a.idepth_hessian[p] = 0.f; a.maxRelBaseline[p] = 0.f; }
        } else {
          float H = HddA + HddL + a.priorF[p];
          if (H < 1e-10) H = 1e-10;
          const float HdiF = (float)(1.0 / (double)H);
          float bdSum = bdA + bdL;
          if (a.shiftPriorToZero) bdSum += a.priorF[p] * a.deltaF[p];
          if (sub == 0) {
            a.HdiF[p] = HdiF; a.bdSumF[p] = bdSum; a.idepth_hessian[p] = H;
            Ws[lp] = HdiF;
            g[D] = bdSum;
          }
#pragma unroll
          for (int i = 0; i < 4; i++) if (sub == i) g[i] = HcdA[i] + HcdL[i];
#pragma unroll
          for (int i = 0; i < 8; i++) if (sub == i) atomicAdd(&g[4 + 8 * host + i], gh[i]);
        }
      }
    }
    __syncthreads();
    // phase 2
#pragma unroll
    for (int m = 0; m < SC_MAXT; m++) {
      if (my_ti[m] < 0) continue;
      const int i0 = my_ti[m] * 4, j0 = my_tj[m] * 4;
      for (int k = 0; k < SC_TP; k++) {
        const float w = Ws[k];
        if (w == 0.f) continue;
        const float4 gi = *(const float4 *)(Gs + k * DPAD + i0);
        const float4 gj = *(const float4 *)(Gs + k * DPAD + j0);
        const float wi[4] = {gi.x * w, gi.y * w, gi.z * w, gi.w * w};
        const float gjv[4] = {gj.x, gj.y, gj.z, gj.w};
#pragma unroll
        for (int ii = 0; ii < 4; ii++)
#pragma unroll
          for (int jj = 0; jj < 4; jj++) acc[m][ii * 4 + jj] += wi[ii] * gjv[jj];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < SC_MAXT; m++) {
    if (my_ti[m] < 0) continue;
    const int i0 = my_ti[m] * 4, j0 = my_tj[m] * 4;
#pragma unroll
    for (int ii = 0; ii < 4; ii++)
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const int i = i0 + ii, j = j0 + jj;
        if (i <= j && j < DP && acc[m][ii * 4 + jj] != 0.f) atomicAdd(a.accSC + (size_t)i * DP + j, (double)acc[m][ii * 4 + jj]);
      }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused accumulation of one Gauss-Newton iteration (accumulateAF_MT + accumulateLF_MT + accumulateSCF_MT,
// EnergyFunctional.cpp:197-254) over tiles of SC_TP points: the commit records of the tile's residuals (contiguous,
// point-major) arrive in shared memory by ONE TMA bulk copy and are read from there by all three consumers:
//   top blocks   warp w walks target t = w, w+8, ... over the tile's points (same host for a whole tile in practice:
//                points are listed host by host), 91 block entries in registers, flushed with fp64 red.global
//   point sums   8 lanes per point (Hdd / bd / Hcd, A and L), HdiF, bdSumF, and g
//   Schur        register-tiled SYRK acc += w g g^T in stitched space
// The last CTA of the grid does not accumulate: it runs the pending setNewFrameEnergyTH selection of the preceding
// linearisation, off the critical path.
__device__ __forceinline__ unsigned smem_u32a(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

constexpr int ADJ_ST = 68;   // floats per staged 8x8 adjoint (64 + 4: float4 rows of different targets fall into different banks)

constexpr int ACC_THREADS = 512;   // warps 0..7: point sums (8 lanes per point), warps 8..15: top blocks (one target each), concurrently
__global__ void __launch_bounds__(ACC_THREADS) k_accumulate_fused(FusedAccArgs a, int DP, int DPAD, int ntiles4, int tiles_total, int max_res, int smem_words) {
  extern __shared__ __align__(16) float smem[];
  PDL_ENTER_T(a.trace);
  if (a.gate && *a.gate) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (a.do_th && blockIdx.x == gridDim.x - 1) { energy_th_body(a.th, (unsigned *)smem, smem_words); TRACE_EXIT(a.trace); return; }
  const int nf = a.nf, D = a.D;
  float *Rs = smem;                                   // [max_res][SOSBA_CREC]
  float *Gs = Rs + (size_t)max_res * SOSBA_CREC;      // [SC_TP][DPAD]
  float *Ws = Gs + SC_TP * DPAD;                      // [SC_TP]
  float *Adh = Ws + SC_TP;                            // [nf][ADJ_ST] adHostF[h0 + t*nf]
  float *Adt = Adh + nf * ADJ_ST;                     // [nf][ADJ_ST] adTargetF[h0 + t*nf]
  int *s_rb = (int *)(Adt + nf * ADJ_ST);             // [SC_TP + 1]
  unsigned short *s_list = (unsigned short *)(s_rb + SC_TP + 1);   // [nf][2][SC_TP] tile-local residuals per target: active | linearised
  unsigned char *s_flag = (unsigned char *)(s_list + nf * 2 * SC_TP);   // [max_res] bit0 use, bit1 linearised
  unsigned char *s_tgt = s_flag + max_res;            // [max_res]
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int s_nacc[2];
  const int nt4 = DPAD / 4;
  int my_ti[SC_MAXT], my_tj[SC_MAXT];
  float acc[SC_MAXT][16];
#pragma unroll
  for (int m = 0; m < SC_MAXT; m++) {
    int t = tid + m * ACC_THREADS;
    my_ti[m] = -1; my_tj[m] = 0;
    if (t < ntiles4) {
      int ti = 0, base = 0;
      while (t >= base + (nt4 - ti)) { base += nt4 - ti; ti++; }
      my_ti[m] = ti; my_tj[m] = ti + (t - base);
    }
#pragma unroll
    for (int q = 0; q < 16; q++) acc[m][q] = 0.f;
  }
  const LaneEntries LE = lane_entries(lane);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32a(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_nacc[0] = s_nacc[1] = 0;
  }
  const int lp = tid >> 3, sub = tid & 7;
  const unsigned gmask = 0xFFu << ((tid & 31) & ~7);
  unsigned phase = 0;
  const int nworkers = a.do_th ? gridDim.x - 1 : gridDim.x;
#define ACC_TS(n) do { if (a.dbg && blockIdx.x == 0 && tid == 32) a.dbg[n] = clock64(); } while (0)
#define ACC_TS2(n) do { if (a.dbg && blockIdx.x == 0 && tid == 288) a.dbg[n] = clock64(); } while (0)
  ACC_TS(0);
  for (int tile = blockIdx.x; tile < tiles_total; tile += nworkers) {
    const int4 td = a.tiles[tile];   // first point, points (<= SC_TP, one host), first residual, residuals
    const int p0 = td.x, np = td.y, rb = td.z, nres = td.w;
    __syncthreads();   // previous tile fully consumed (and the mbarrier initialised)
    if (tid == 0 && nres > 0) {
      const unsigned bytes = (unsigned)nres * SOSBA_CREC * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32a(&mbar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32a(Rs)),
                   "l"(a.rec + (size_t)rb * SOSBA_CREC), "r"(bytes), "r"(smem_u32a(&mbar))
                   : "memory");
    }
    const int host = a.p_host[p0];
    for (int i = tid; i < nres; i += ACC_THREADS) {
      const int r = rb + i;
      const bool use = a.r_is_active[r] && !a.r_dropped[r];
      s_flag[i] = (unsigned char)((use ? 1 : 0) | (a.r_is_lin[r] ? 2 : 0));
      s_tgt[i] = (unsigned char)a.r_target[r];
    }
    for (int e = tid; e < nf * 16; e += ACC_THREADS) {   // the 2 x nf adjoints of this host, 16 float4 each
      const int t = e >> 4, k = e & 15;
      *(float4 *)(Adh + t * ADJ_ST + 4 * k) = __ldg((const float4 *)(a.adHostF + 64 * (size_t)(host + t * nf)) + k);
      *(float4 *)(Adt + t * ADJ_ST + 4 * k) = __ldg((const float4 *)(a.adTargetF + 64 * (size_t)(host + t * nf)) + k);
    }
    if (tid <= np) s_rb[tid] = a.res_begin[p0 + tid] - rb;
    for (int i = tid; i < SC_TP * DPAD; i += ACC_THREADS) Gs[i] = 0.f;
    if (tid < SC_TP) Ws[tid] = 0.f;
    __syncthreads();
    if (nres > 0) {
      unsigned done = 0;
      while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32a(&mbar)), "r"(phase) : "memory");
      phase ^= 1;
    }
    ACC_TS(1);
    // ---- top blocks: warp w owns target t = w, w+8, ...  Lane lp looks up the residual of ITS point towards t (a point has
    // at most one: a scan of its <= nf-1 entries), a vote compacts the hits into lane order = residual order (the tile is
    // point-major), and the warp sums them: no shared-memory lists, the same deterministic order as the reference's loop.
    if (warp >= 8) {
      int nA = 0, nL = 0;
      const int lb = lane < np ? s_rb[lane] : 0, le = lane < np ? s_rb[lane + 1] : 0;
      int deg = le - lb;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) deg = max(deg, __shfl_xor_sync(0xffffffffu, deg, o));
      for (int t = warp - 8; t < nf; t += 8) {
        int idx = 0, f = 0;
        for (int k = 0; k < deg; k++) {
          const int i = lb + k;
          if (i < le && s_tgt[i] == t) { idx = i; f = s_flag[i]; }
        }
#pragma unroll
        for (int l = 0; l < 2; l++) {
          const bool m = (f & 1) && ((f >> 1) == l);
          const unsigned bal = __ballot_sync(0xffffffffu, m);
          const int c = __popc(bal);
          if (c == 0) continue;
          // lane j takes the residual of the j-th voting lane
          const int src = lane < c ? (int)__fns(bal, 0, lane + 1) : 0;
          const int cidx = __shfl_sync(0xffffffffu, idx, src);
          LaneAcc A = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
          for (int n = 0; n < c; n++) lane_accumulate(Rs + (size_t)__shfl_sync(0xffffffffu, cidx, n) * SOSBA_CREC, LE, A);
          if (l == 0 && a.dbg && blockIdx.x == 0 && tid == 288) a.dbg[8] = clock64() + (long long)(A.q0 == 123.f);
          lane_flush(a.accTop + ((size_t)l * nf * nf + host + t * nf) * SOSBA_TOPB, lane, LE, A, c);
          if (l == 0) nA += c; else nL += c;
        }
        ACC_TS2(9);
      }
      if (lane == 0) { if (nA) atomicAdd(&s_nacc[0], nA); if (nL) atomicAdd(&s_nacc[1], nL); }
      ACC_TS2(2);
    }
    // ---- point sums + g (warps 0..7, concurrently with the top blocks) ------------------------------------------
    if (warp < 8 && lp < np) {   // uniform across the 8 lanes of a point
      const int p = p0 + lp;
      const int lb = s_rb[lp], le = s_rb[lp + 1];
      float *g = Gs + lp * DPAD;
      float HddA = 0.f, bdA = 0.f, HcdA[4] = {0.f, 0.f, 0.f, 0.f}, HddL = 0.f, bdL = 0.f, HcdL[4] = {0.f, 0.f, 0.f, 0.f};
      float gh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int ngood = 0;
      for (int base = lb; base < le; base += 8) {   // one round for nf <= 9
        const int i = base + sub;
        const int f = i < le ? s_flag[i] : 0;
        const bool use = f & 1;
        const bool lin = use && (f & 2);
        float c_bd = 0.f, c_Hdd = 0.f, c_Hcd[4] = {0.f, 0.f, 0.f, 0.f};
        float sh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (use) {
          const float4 *rec4 = (const float4 *)(Rs + (size_t)i * SOSBA_CREC);
          const float4 x03 = rec4[0];                    // CR_X 0..3
          const float4 y_a = rec4[2], y_b = rec4[3];     // floats 8..15: x[8],x[9],y[0],y[1] | y[2],y[3],y[4],y[5]
          const float4 f20 = rec4[5], f24 = rec4[6], f28 = rec4[7], f32 = rec4[8], f36 = rec4[9];
          const float4 v03 = rec4[10], v47 = rec4[11];   // JpJdF
          const float a00 = f20.x, a01 = f20.y, a11 = f20.z;               // CR_A 20..22
          const float JI_r0 = f24.w, JI_r1 = f28.x;                         // CR_TR+4 = 27, 28
          const float Jpdd0 = f32.w, Jpdd1 = f36.x;                         // CR_JPDD = 35, 36
          const float v0 = a00 * Jpdd0 + a01 * Jpdd1, v1 = a01 * Jpdd0 + a11 * Jpdd1;  // JIdx2 * Jpdd
          c_bd = JI_r0 * Jpdd0 + JI_r1 * Jpdd1;
          c_Hdd = v0 * Jpdd0 + v1 * Jpdd1;
          const float xs[4] = {x03.x, x03.y, x03.z, x03.w}, ys[4] = {y_a.z, y_a.w, y_b.x, y_b.y};
#pragma unroll
          for (int q = 0; q < 4; q++) c_Hcd[q] = xs[q] * v0 + ys[q] * v1;
          const int t = s_tgt[i];
          const float4 *Ah = (const float4 *)(Adh + t * ADJ_ST), *At = (const float4 *)(Adt + t * ADJ_ST);
          const float v[8] = {v03.x, v03.y, v03.z, v03.w, v47.x, v47.y, v47.z, v47.w};
          float st[8];
#pragma unroll
          for (int row = 0; row < 8; row++) {
            const float4 h0 = Ah[2 * row], h1 = Ah[2 * row + 1], t0 = At[2 * row], t1 = At[2 * row + 1];
            sh[row] = h0.x * v[0] + h0.y * v[1] + h0.z * v[2] + h0.w * v[3] + h1.x * v[4] + h1.y * v[5] + h1.z * v[6] + h1.w * v[7];
            st[row] = t0.x * v[0] + t0.y * v[1] + t0.z * v[2] + t0.w * v[3] + t1.x * v[4] + t1.y * v[5] + t1.z * v[6] + t1.w * v[7];
          }
          // one residual per (point, target): this lane is the only writer of the target rows
          *(float4 *)(g + 4 + 8 * t) = make_float4(st[0], st[1], st[2], st[3]);
          *(float4 *)(g + 8 + 8 * t) = make_float4(st[4], st[5], st[6], st[7]);
        }
        ngood += __popc(__ballot_sync(gmask, use) & gmask);
        const int cnt = min(8, le - base);
        for (int q = 0; q < cnt; q++) {   // residual-order sums (deterministic, reference order)
          const bool ql = __shfl_sync(gmask, (int)lin, q, 8) != 0;
          const float q_bd = __shfl_sync(gmask, c_bd, q, 8), q_Hdd = __shfl_sync(gmask, c_Hdd, q, 8);
          if (ql) { bdL += q_bd; HddL += q_Hdd; } else { bdA += q_bd; HddA += q_Hdd; }
#pragma unroll
          for (int k = 0; k < 4; k++) { const float qc = __shfl_sync(gmask, c_Hcd[k], q, 8); if (ql) HcdL[k] += qc; else HcdA[k] += qc; }
        }
#pragma unroll
        for (int row = 0; row < 8; row++) {
          float sv = sh[row];
          sv += __shfl_xor_sync(gmask, sv, 1, 8); sv += __shfl_xor_sync(gmask, sv, 2, 8); sv += __shfl_xor_sync(gmask, sv, 4, 8);
          gh[row] += sv;
        }
      }
      if (sub == 0) {
        a.HddA[p] = HddA; a.bdA[p] = bdA; a.HddL[p] = HddL; a.bdL[p] = bdL;
        *(float4 *)(a.HcdA + 4 * (size_t)p) = make_float4(HcdA[0], HcdA[1], HcdA[2], HcdA[3]);
        *(float4 *)(a.HcdL + 4 * (size_t)p) = make_float4(HcdL[0], HcdL[1], HcdL[2], HcdL[3]);
      }
      if (ngood == 0) {
        if (sub == 0) { a.HdiF[p] = 0.f; a.bdSumF[p] = 0.f; a.idepth_hessian[p] = 0.f; a.maxRelBaseline[p] = 0.f; }
      } else {
        float H = HddA + HddL + a.priorF[p];
        if (H < 1e-10) H = 1e-10;
        const float HdiF = (float)(1.0 / (double)H);
        float bdSum = bdA + bdL;
        if (a.shiftPriorToZero) bdSum += a.priorF[p] * a.deltaF[p];
        if (sub == 0) {
          a.HdiF[p] = HdiF; a.bdSumF[p] = bdSum; a.idepth_hessian[p] = H;
          Ws[lp] = HdiF;
          g[D] = bdSum;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) if (sub == q) g[q] = HcdA[q] + HcdL[q];
#pragma unroll
        for (int q = 0; q < 8; q++) if (sub == q) g[4 + 8 * host + q] = gh[q];   // no residual targets the host itself
      }
    }
    ACC_TS(6);
    __syncthreads();
    if (a.dbg && blockIdx.x == 0 && tid == 32) a.dbg[3] = clock64() + (long long)(Ws[0] == 123.f);
    // ---- Schur: acc += w g g^T (a point without good residuals has w = 0 and an all-zero g) ----------------------
#pragma unroll
    for (int m = 0; m < SC_MAXT; m++) {
      if (my_ti[m] < 0) continue;
      const int i0 = my_ti[m] * 4, j0 = my_tj[m] * 4;
#pragma unroll 4
      for (int k = 0; k < SC_TP; k++) {
        const float w = Ws[k];
        const float4 gi = *(const float4 *)(Gs + k * DPAD + i0);
        const float4 gj = *(const float4 *)(Gs + k * DPAD + j0);
        const float wi[4] = {gi.x * w, gi.y * w, gi.z * w, gi.w * w};
        const float gjv[4] = {gj.x, gj.y, gj.z, gj.w};
#pragma unroll
        for (int ii = 0; ii < 4; ii++)
#pragma unroll
          for (int jj = 0; jj < 4; jj++) acc[m][ii * 4 + jj] += wi[ii] * gjv[jj];
      }
    }
  }
  ACC_TS(4);
#pragma unroll
  for (int m = 0; m < SC_MAXT; m++) {
    if (my_ti[m] < 0) continue;
    const int i0 = my_ti[m] * 4, j0 = my_tj[m] * 4;
#pragma unroll
    for (int ii = 0; ii < 4; ii++)
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const int i = i0 + ii, j = j0 + jj;
        if (i <= j && j < DP && acc[m][ii * 4 + jj] != 0.f) atomicAdd(a.accSC + (size_t)i * DP + j, (double)acc[m][ii * 4 + jj]);
      }
  }
  ACC_TS(5);
  __syncthreads();
  if (tid < 2 && s_nacc[tid]) atomicAdd(a.n_acc + tid, s_nacc[tid]);
  TRACE_EXIT(a.trace);
}

// ------------------------------------------------------------------------------------------------
// One CTA (64 threads) per (host,target) block: accH (13x13 fp64) -> H, b with the adjoint sandwich.
__global__ void __launch_bounds__(64) k_stitch_top(const double *__restrict__ accTop, const double *__restrict__ adHost,
                                                   const double *__restrict__ adTarget, int nf, double *__restrict__ H, double *__restrict__ b) {
  __shared__ double accH[13][13];
  __shared__ double Ah[64], At[64], AhP[64], AtP[64];
  PDL_ENTER();
  const int blk = blockIdx.x % (nf * nf);   // = h + nf*t ; blockIdx.x / nf^2 selects the table (A pass, L pass)
  const int h = blk % nf, t = blk / nf;
  const double *src = accTop + (size_t)blockIdx.x * SOSBA_TOPB;
  if (src[91] == 0.0) return;          // num == 0 (AccumulatedTopHessian.cpp:168-170)
  const int tid = threadIdx.x;
  const int D = 4 + 8 * nf;
  for (int e = tid; e < 91; e += 64) {
    const EntryDesc d = entry_desc(e);
    int r, c;
    if (d.kind == 0) { r = d.p; c = d.q; }
    else if (d.kind == 1) { r = d.p; c = 10 + d.q; }
    else { const int rr[6] = {10, 10, 10, 11, 11, 12}, cc[6] = {10, 11, 12, 11, 12, 12}; r = rr[d.p]; c = cc[d.p]; }
    accH[r][c] = src[e]; accH[c][r] = src[e];
  }
  Ah[tid] = adHost[64 * (size_t)blk + tid];
  At[tid] = adTarget[64 * (size_t)blk + tid];
  __syncthreads();
  const int i = tid >> 3, j = tid & 7;
  {  // AhP = Ah * P, AtP = At * P with P = accH[4:12, 4:12]
    double sh = 0, st = 0;
    for (int k = 0; k < 8; k++) { sh += Ah[i * 8 + k] * accH[4 + k][4 + j]; st += At[i * 8 + k] * accH[4 + k][4 + j]; }
    AhP[tid] = sh; AtP[tid] = st;
  }
  __syncthreads();
  const int hIdx = 4 + 8 * h, tIdx = 4 + 8 * t;
  {
    double hh = 0, tt = 0, ht = 0;
    for (int k = 0; k < 8; k++) { hh += AhP[i * 8 + k] * Ah[j * 8 + k]; tt += AtP[i * 8 + k] * At[j * 8 + k]; ht += AhP[i * 8 + k] * At[j * 8 + k]; }
    atomicAdd(&H[(size_t)(hIdx + i) * D + hIdx + j], hh);
    atomicAdd(&H[(size_t)(tIdx + i) * D + tIdx + j], tt);
    atomicAdd(&H[(size_t)(hIdx + i) * D + tIdx + j], ht);
  }
  if (j < 4) {  // frame-calib blocks
    double sh = 0, st = 0;
    for (int k = 0; k < 8; k++) { sh += Ah[i * 8 + k] * accH[4 + k][j]; st += At[i * 8 + k] * accH[4 + k][j]; }
    atomicAdd(&H[(size_t)(hIdx + i) * D + j], sh);
    atomicAdd(&H[(size_t)(tIdx + i) * D + j], st);
  }
  if (j == 4) {  // b
    double sh = 0, st = 0;
    for (int k = 0; k < 8; k++) { sh += Ah[i * 8 + k] * accH[4 + k][12]; st += At[i * 8 + k] * accH[4 + k][12]; }
    atomicAdd(&b[hIdx + i], sh);
    atomicAdd(&b[tIdx + i], st);
  }
  if (tid < 16) atomicAdd(&H[(size_t)(tid >> 2) * D + (tid & 3)], accH[tid >> 2][tid & 3]);
  if (tid >= 16 && tid < 20) atomicAdd(&b[tid - 16], accH[tid - 16][12]);
}

// priors + symmetrisation epilogue of stitchDoubleMT (AccumulatedTopHessian.h:107-126, .cpp:292-300). One CTA.
__global__ void __launch_bounds__(256) k_finalize_top(int nf, double *__restrict__ H, double *__restrict__ b, int usePrior,
                                                      const double *__restrict__ wprior, const float *__restrict__ cDeltaF) {
  const int D = 4 + 8 * nf, tid = threadIdx.x;
  if (usePrior) {
    const double *cPrior = wprior, *fprior = wprior + 4, *fdp = wprior + 4 + 8 * nf;
    for (int i = tid; i < D; i += 256) {
      if (i < 4) { H[(size_t)i * D + i] += cPrior[i]; b[i] += cPrior[i] * (double)cDeltaF[i]; }
      else { H[(size_t)i * D + i] += fprior[i - 4]; b[i] += fprior[i - 4] * fdp[i - 4]; }
    }
  }
  __syncthreads();
  // calib-frame blocks: H[0:4, hIdx+j] = H[hIdx+j, 0:4]
  for (int e = tid; e < 4 * 8 * nf; e += 256) {
    const int i = e / (8 * nf), c = 4 + e % (8 * nf);
    H[(size_t)i * D + c] = H[(size_t)c * D + i];
  }
  // off-diagonal frame blocks: H[h,t] += H[t,h]^T ; H[t,h] = H[h,t]^T   (h < t)
  for (int e = tid; e < 64 * nf * nf; e += 256) {
    const int blk = e >> 6, i = (e >> 3) & 7, j = e & 7;
    const int h = blk % nf, t = blk / nf;
    if (h >= t) continue;
    const size_t ht = (size_t)(4 + 8 * h + i) * D + 4 + 8 * t + j, th = (size_t)(4 + 8 * t + j) * D + 4 + 8 * h + i;
    const double v = H[ht] + H[th];
    H[ht] = v; H[th] = v;
  }
}

// the priors the linearised pass carries (AccumulatedTopHessian.cpp:292-300) onto an already symmetric H, b. One CTA.
__global__ void __launch_bounds__(128) k_add_priors(int nf, double *__restrict__ H, double *__restrict__ b, const double *__restrict__ wprior,
                                                    const float *__restrict__ cDeltaF) {
  const int D = 4 + 8 * nf;
  const double *cPrior = wprior, *fprior = wprior + 4, *fdp = wprior + 4 + 8 * nf;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const double pr = i < 4 ? cPrior[i] : fprior[i - 4], dp = i < 4 ? (double)cDeltaF[i] : fdp[i - 4];
    H[(size_t)i * D + i] += pr;
    b[i] += pr * dp;
  }
}

// accSC (upper tiles of the (D+1)^2 Gram matrix) -> full symmetric Hsc and bsc. One CTA.
__global__ void __launch_bounds__(256) k_finalize_sc(const double *__restrict__ accSC, int D, double *__restrict__ H, double *__restrict__ b) {
  const int DP = D + 1;
  for (int e = threadIdx.x; e < D * D; e += 256) {
    const int i = e / D, j = e % D;
    // tiles with ti <= tj are stored; read the upper triangle only so that the result is exactly symmetric
    H[e] = i <= j ? accSC[(size_t)i * DP + j] : accSC[(size_t)j * DP + i];
  }
  for (int i = threadIdx.x; i < D; i += 256) b[i] = accSC[(size_t)i * DP + D];
}

// ------------------------------------------------------------------------------------------------
// a11  resubstituteFPt (+ the point part of backupState / doStepFromBackup when do_step); 8 lanes per point,
// lane q owns residual res_begin[p] + q; the subtraction chain runs in residual order like the reference.
__global__ void __launch_bounds__(256) k_resubstitute(ResubArgs a) {
  PDL_ENTER();
  if (a.gate && *a.gate) return;
  resubstitute_body(a, blockIdx.x, gridDim.x);
}

// EnergyFunctional::marginalizePointsF: p->priorF *= setting_idepthFixPriorMargFac (EnergyFunctional.cpp:901)
__global__ void k_scale_prior(float *__restrict__ priorF, const int *__restrict__ ids, int n, float fac) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) priorF[ids[i]] *= fac;
}

}  // namespace

namespace {
__global__ void k_fill_f32(float *__restrict__ dst, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
}  // namespace
void launch_fill_f32(sosba *h, float *dst, int n, float v) {
  if (n == 0) return;
  k_fill_f32<<<(n + 255) / 256, 256, 0, h->stream>>>(dst, n, v);
  h->launches++;
}

void launch_scale_prior(sosba *h, float *priorF, const int *ids, int n, float fac) {
  if (n == 0) return;
  k_scale_prior<<<(n + 255) / 256, 256, 0, h->stream>>>(priorF, ids, n, fac);
  h->launches++;
}

// ------------------------------------------------------------------------------------------------
void launch_top_accumulate(sosba *h, const AccArgs &a) {
  if (a.n_list == 0) return;
  // enough warps to fill the machine, at least 4 residuals per warp so a flush is amortised
  int warps = h->sm_count * TOP_WARPS * 2;
  int chunk = (a.n_list + warps - 1) / warps;
  if (chunk < 4) chunk = 4;
  warps = (a.n_list + chunk - 1) / chunk;
  const int blocks = (warps + TOP_WARPS - 1) / TOP_WARPS;
  k_top_accumulate<<<blocks, TOP_WARPS * 32, 0, h->stream>>>(a, chunk);
  h->launches++;
}

void launch_point_sc(sosba *h, const SCArgs &a) {
  const int np = a.plist ? a.n_plist : a.P;
  if (np == 0) return;
  const int DP = a.D + 1, DPAD = (DP + 3) / 4 * 4, nt4 = DPAD / 4;
  const int ntiles4 = nt4 * (nt4 + 1) / 2;
  const int tiles_total = (np + SC_TP - 1) / SC_TP;
  int blocks = tiles_total < 2 * h->sm_count ? tiles_total : 2 * h->sm_count;
  const size_t smem = (size_t)(SC_TP * DPAD + SC_TP) * sizeof(float);
  k_point_sc<<<blocks, 256, smem, h->stream>>>(a, DP, DPAD, ntiles4, tiles_total);
  h->launches++;
}

void launch_stitch_top(sosba *h, const double *accTop, const double *adHost, const double *adTarget, int nf, double *H, double *b, int usePrior,
                       const double *wprior, const float *cDeltaF) {
  k_stitch_top<<<nf * nf, 64, 0, h->stream>>>(accTop, adHost, adTarget, nf, H, b);
  k_finalize_top<<<1, 256, 0, h->stream>>>(nf, H, b, usePrior, wprior, cDeltaF);
  h->launches += 2;
}

void launch_add_priors(sosba *h, int nf, double *H, double *b, const double *wprior, const float *cDeltaF) {
  k_add_priors<<<1, 128, 0, h->stream>>>(nf, H, b, wprior, cDeltaF);
  h->launches++;
}

void launch_finalize_sc(sosba *h, const double *accSC, int nf, double *H, double *b) {
  k_finalize_sc<<<1, 256, 0, h->stream>>>(accSC, 4 + 8 * nf, H, b);
  h->launches++;
}

void launch_resubstitute(sosba *h, const ResubArgs &a) {
  if (a.P == 0) return;
  launch_pdl(k_resubstitute, (a.P * 8 + 255) / 256, 256, 0, h->stream, a);
  h->launches++;
}

// mode-0 accumulation (A, L and Schur tables) in one launch.  Returns false when the tile cannot be staged in shared
// memory (the caller falls back to the separate kernels).
bool launch_accumulate_fused(sosba *h, const FusedAccArgs &a, int max_res_per_tile, int tiles_total) {
  if (a.P == 0 || tiles_total == 0) return true;
  const int DP = a.D + 1, DPAD = (DP + 3) / 4 * 4, nt4 = DPAD / 4;
  const int ntiles4 = nt4 * (nt4 + 1) / 2;
  if (ntiles4 > SC_MAXT * ACC_THREADS || a.nf > 255) return false;
  const int max_res = (max_res_per_tile > 0 ? max_res_per_tile + 3 : 4) & ~3;
  if (max_res > 60000) return false;
  size_t smem = (size_t)max_res * SOSBA_CREC * 4 + (size_t)(SC_TP * DPAD + SC_TP) * 4 + (size_t)2 * a.nf * ADJ_ST * 4 + (size_t)(SC_TP + 1) * 4 +
                (size_t)a.nf * 2 * SC_TP * 2 + 2 * (size_t)max_res;
  smem = (smem + 15) & ~(size_t)15;
  if (smem > 200 * 1024) return false;
  static bool configured[64] = {};   // the attribute is per device: one flag per device ordinal (several handles in one process)
  if (smem > 48 * 1024 && !(h->device >= 0 && h->device < 64 && configured[h->device])) {
    if (cudaFuncSetAttribute(k_accumulate_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return false;
    if (h->device >= 0 && h->device < 64) configured[h->device] = true;
  }
  int workers = tiles_total < 2 * h->sm_count ? tiles_total : 2 * h->sm_count;
  launch_pdl(k_accumulate_fused, workers + (a.do_th ? 1 : 0), ACC_THREADS, smem, h->stream, a, DP, DPAD, ntiles4, tiles_total, max_res, (int)(smem / 4));
  h->launches++;
  return true;
}
