// k_exact.cu — kernels whose per-element float expressions decide integer bookkeeping (residual state,
// counts) and therefore follow the reference's operation order with no FMA contraction.  This file is
// compiled with -fmad=false (the parity definition pins -ffp-contract=off on the CPU side).
//
//   make_images          FrameHessian::makeImages                     src/FullSystem/HessianBlocks.cpp:121-176
//   linearize            PointFrameResidual::linearize                src/FullSystem/Residuals.cpp:77-271
//                        projectPoint x2                              src/FullSystem/ResidualProjections.h:43-73
//                        getInterpolatedElement33                     src/util/globalFuncs.h:68-82
//   apply_res            PointFrameResidual::applyRes + takeDataF     Residuals.cpp:304-321, EnergyFunctionalStructs.cpp:36-45
//                        + the fixLinearization bookkeeping of linearizeAll_Reductor  FullSystemOptimize.cpp:52-74
//   fix_linearization    EFResidual::fixLinearizationF                EnergyFunctionalStructs.cpp:75-103
//   prep_records         resApprox / JI_r / Jab_r / rr of addPoint<1>, addPoint<2>   AccumulatedTopHessian.cpp:68-110
//   energy_th            FullSystem::setNewFrameEnergyTH              FullSystemOptimize.cpp:84-124
//   track_res            CoarseTracker::calcResPose / ScaleOptimizer::calcResScale   CoarseTracker.cpp:612-764, ScaleOptimizer.cpp:273-437
#include <math.h>
#include <stdlib.h>

#include "energy_th.cuh"
#include "host_math.h"
#include "kernels.h"

namespace {

__constant__ int c_pattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};  // settings.cpp:307-309

constexpr float kSCALE_IDEPTH = 1.0f, kSCALE_F = 50.0f, kSCALE_C = 50.0f;  // HessianBlocks.h:53-60

// ------------------------------------------------------------------------------------------------
// a1  image pyramid.  Level 0 reads the irradiance plane; level l recomputes the 2x2 means of level l-1
// for the pixel and its 4 flat-index neighbours (the reference differentiates over a flat index, so the
// first/last column difference reaches into the neighbouring row — reproduced here).  Rows 0 and h-1
// keep dx = dy = absSquaredGrad = 0 (uninitialised in the reference).
__device__ __forceinline__ float down4(const float *__restrict__ prev, int wm, int wl, int j) {
  int x = j % wl, y = j / wl;
  const float *p = prev + 2 * x + 2 * y * wm;
  return 0.25f * (((p[0] + p[1]) + p[wm]) + p[1 + wm]);
}

__global__ void __launch_bounds__(256) k_pyr(const float *__restrict__ src, int wm, float4 *__restrict__ img, float *__restrict__ plane, int w, int h,
                                             int lvl, const float *__restrict__ B, int gamma) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w * h) return;
  float I, dx = 0.f, dy = 0.f, ab = 0.f;
  const bool inner = idx >= w && idx < w * (h - 1);
  if (lvl == 0) {
    I = src[idx];
    if (inner) {
      dx = 0.5f * (src[idx + 1] - src[idx - 1]);
      dy = 0.5f * (src[idx + w] - src[idx - w]);
    }
  } else {
    I = down4(src, wm, w, idx);
    if (inner) {
      dx = 0.5f * (down4(src, wm, w, idx + 1) - down4(src, wm, w, idx - 1));
      dy = 0.5f * (down4(src, wm, w, idx + w) - down4(src, wm, w, idx - w));
    }
  }
  if (inner) {
    if (!isfinite(dx)) dx = 0.f;
    if (!isfinite(dy)) dy = 0.f;
    ab = dx * dx + dy * dy;
    if (gamma == 1 && B != nullptr) {  // CalibHessian::getBGradOnly, HessianBlocks.h:519-526
      int c = (int)(I + 0.5f);
      if (c < 5) c = 5;
      if (c > 250) c = 250;
      float gw = B[c + 1] - B[c];
      ab *= gw * gw;
    }
  }
  img[idx] = make_float4(I, dx, dy, ab);
  plane[idx] = I;
}

// a1, ALL LEVELS IN ONE LAUNCH (the default where the shape allows it, launch_make_images): a CTA owns a 64 x 32 tile of level 0 and everything below it.  The
// irradiance tile plus a halo of 2^(levels-1) pixels is staged in shared memory by TMA bulk copies (one cp.async.bulk per row,
// completion on an mbarrier), the 2x2 means of every level are formed there (same expression, same float values as the
// level-by-level kernel reads back from its planar copies), and each level's {I, dx, dy, absSquaredGrad} texels are written
// from shared-memory neighbours.  The only values a tile cannot see are the flat-index neighbours of the first / last column
// (the reference differentiates over a flat index, so they sit at the other end of the neighbouring row): those few are
// rebuilt from the level-0 image with the same expression tree.
constexpr int PYR_TW = 64, PYR_TH = 32, PYR_MAXL = 5;
struct PyrArgs {
  const float *src;          // level-0 irradiance, w * h
  float4 *img[PYR_MAXL];     // per level
  int w, h, levels, gamma;
  const float *B;
};
__device__ float pyr_level_value(const float *__restrict__ src, int w0, int l, int x, int y) {   // I_l(x, y) from level 0
  if (l == 0) return __ldg(src + (size_t)y * w0 + x);
  const float a = pyr_level_value(src, w0, l - 1, 2 * x, 2 * y), b = pyr_level_value(src, w0, l - 1, 2 * x + 1, 2 * y);
  const float c = pyr_level_value(src, w0, l - 1, 2 * x, 2 * y + 1), d = pyr_level_value(src, w0, l - 1, 2 * x + 1, 2 * y + 1);
  return 0.25f * (((a + b) + c) + d);
}
__global__ void __launch_bounds__(256) k_pyr_fused(PyrArgs a) {
  extern __shared__ __align__(16) float s_pl[];   // planes of the levels, each (SW >> l) x (SH >> l), level 0 first
  __shared__ __align__(8) unsigned long long mbar;
  const int L = a.levels, H = 1 << (L - 1);
  const int SW = PYR_TW + 2 * H, SH = PYR_TH + 2 * H;
  const int X0 = blockIdx.x * PYR_TW - H, Y0 = blockIdx.y * PYR_TH - H;   // level-0 coordinates of the staged region's corner
  const int tid = threadIdx.x;
  // ---- stage level 0: rows inside the image, columns clipped to the image (all multiples of 4 floats)
  const int xs = max(X0, 0), xe = min(X0 + SW, a.w), ys = max(Y0, 0), ye = min(Y0 + SH, a.h);
  const unsigned row_bytes = (unsigned)(xe - xs) * 4u;
  const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(row_bytes * (unsigned)(ye - ys)) : "memory");
  }
  __syncthreads();
  for (int r = ys + tid; r < ye; r += blockDim.x) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(s_pl + (size_t)(r - Y0) * SW + (xs - X0));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(a.src + (size_t)r * a.w + xs), "r"(row_bytes), "r"(mb)
                 : "memory");
  }
  {
    unsigned done = 0;
    while (!done) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(mb), "r"(0u) : "memory");
  }
  // ---- the 2x2 means, level by level, over the part of the staged region that lies inside the image
  int off_prev = 0;
  for (int l = 1; l < L; l++) {
    const int wp = SW >> (l - 1), wl = SW >> l, hl = SH >> l;
    const int off = off_prev + wp * (SH >> (l - 1));
    const int gx0 = X0 >> l, gy0 = Y0 >> l, gw = a.w >> l, gh = a.h >> l;   // (X0, Y0 are multiples of 2^l)
    for (int e = tid; e < wl * hl; e += blockDim.x) {
      const int x = e % wl, y = e / wl;
      const int gx = gx0 + x, gy = gy0 + y;
      if (gx < 0 || gy < 0 || gx >= gw || gy >= gh) continue;
      const float *p = s_pl + off_prev + 2 * x + 2 * y * wp;
      s_pl[off + e] = 0.25f * (((p[0] + p[1]) + p[wp]) + p[1 + wp]);
    }
    off_prev = off;
    __syncthreads();
  }
  // ---- texels of every level
  int off = 0;
  for (int l = 0; l < L; l++) {
    const int wl = SW >> l, tw = PYR_TW >> l, th = PYR_TH >> l, hh = H >> l;
    const int gw = a.w >> l, gh = a.h >> l;
    const int gx0 = (X0 + H) >> l, gy0 = (Y0 + H) >> l;   // first level-l pixel of this tile
    float4 *img = a.img[l];
    for (int e = tid; e < tw * th; e += blockDim.x) {
      const int x = e % tw, y = e / tw;
      const int gx = gx0 + x, gy = gy0 + y;
      if (gx >= gw || gy >= gh) continue;
      const float *c = s_pl + off + (x + hh) + (y + hh) * wl;
      const float I = c[0];
      float dx = 0.f, dy = 0.f, ab = 0.f;
      if (gy >= 1 && gy <= gh - 2) {
        const float left = gx > 0 ? c[-1] : pyr_level_value(a.src, a.w, l, gw - 1, gy - 1);        // flat index - 1
        const float right = gx < gw - 1 ? c[1] : pyr_level_value(a.src, a.w, l, 0, gy + 1);        // flat index + 1
        dx = 0.5f * (right - left);
        dy = 0.5f * (c[wl] - c[-wl]);
        if (!isfinite(dx)) dx = 0.f;
        if (!isfinite(dy)) dy = 0.f;
        ab = dx * dx + dy * dy;
        if (a.gamma == 1 && a.B != nullptr) {  // CalibHessian::getBGradOnly, HessianBlocks.h:519-526
          int ci = (int)(I + 0.5f);
          if (ci < 5) ci = 5;
          if (ci > 250) ci = 250;
          const float gwt = a.B[ci + 1] - a.B[ci];
          ab *= gwt * gwt;
        }
      }
      img[(size_t)gy * gw + gx] = make_float4(I, dx, dy, ab);
    }
    off += wl * (SH >> l);
  }
}

// ------------------------------------------------------------------------------------------------
// globalFuncs.h:68-82 on the float4 image: weights and summation order verbatim.
__device__ __forceinline__ float3 interp33(const float4 *__restrict__ img, float x, float y, int width) {
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy;
  float dxdy = dx * dy;
  const float4 *bp = img + ix + iy * width;
  float4 t11 = __ldg(bp + 1 + width), t01 = __ldg(bp + width), t10 = __ldg(bp + 1), t00 = __ldg(bp);
  float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  float3 o;
  o.x = w11 * t11.x + w01 * t01.x + w10 * t10.x + w00 * t00.x;
  o.y = w11 * t11.y + w01 * t01.y + w10 * t10.y + w00 * t00.y;
  o.z = w11 * t11.z + w01 * t01.z + w10 * t10.z + w00 * t00.z;
  return o;
}

// sequential (index-order) sum of one value per pattern lane, identical on all 8 lanes of the group
__device__ __forceinline__ float seqsum8(unsigned mask, float v) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) s += __shfl_sync(mask, v, j, 8);
  return s;
}

__device__ __forceinline__ float4 pick4(int k, float4 c0, float4 c1, float4 c2, float4 c3, float4 c4, float4 c5, float4 c6, float4 c7) {
  float4 r = c0;
  r = k == 1 ? c1 : r; r = k == 2 ? c2 : r; r = k == 3 ? c3 : r; r = k == 4 ? c4 : r;
  r = k == 5 ? c5 : r; r = k == 6 ? c6 : r; r = k == 7 ? c7 : r;
  return r;
}

__device__ __forceinline__ float bfly8(unsigned mask, float v);

// a3  one residual = 8 consecutive lanes (one per pattern sample); 32 residuals per 256-thread block.
// APPLY: PointFrameResidual::applyRes(true) + EFResidual::takeDataF of the same residual fused in (the commit record is
//        written straight from registers; FullSystemOptimize.cpp:363-372 runs linearizeAll then applyRes back to back).
// WRITE_J: materialise the candidate RawResidualJacobian record and projectedTo (the Gauss-Newton loop needs neither).
// TH: the last CTA to finish runs setNewFrameEnergyTH (FullSystemOptimize.cpp:84-124) — no separate launch.
template <bool APPLY, bool WRITE_J, bool TH>
__global__ void __launch_bounds__(256) k_linearize(LinArgs a) {
  __shared__ double s_we[8];    // per-warp energy sums (no shared-memory fp64 atomics: those are CAS loops)
  __shared__ int s_wc[8][3];    // per-warp state histogram
  __shared__ int s_last;
  PDL_ENTER();
  // first round trip: the loop gate and every per-residual id / flag at once (independent loads)
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = gid >> 3, idx = gid & 7;
  const bool inr = r < a.R;
  const int gate = a.gate ? *a.gate : 0;
  const int m_lin = inr ? a.r_is_lin[r] : 1, m_drop = inr ? a.r_dropped[r] : 1, m_state = inr ? a.r_state[r] : 0;
  const int m_pt = inr ? a.r_point[r] : 0, m_host = inr ? a.r_host[r] : 0, m_target = inr ? a.r_target[r] : 0;
  const float m_energy = inr ? a.r_energy[r] : 0.f;
  if (gate) return;   // the Gauss-Newton loop already converged (device-side break)
  if (a.zero_n > 0) {              // clear the block tables of the accumulation that follows (one memset less on the stream)
    double2 *zb = (double2 *)a.zero_buf;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.zero_n; i += gridDim.x * blockDim.x) zb[i] = make_double2(0.0, 0.0);
  }
  __syncthreads();

  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  const bool live = inr && !(m_lin | m_drop);
  int outcome = -1;       // state_NewState once decided
  float ret_energy = 0.f; // linearize() return value
  if (live) {
    const int pt = m_pt, host = m_host, target = m_target;
    const float old_energy = m_energy;
    float energyWO = -1.f;
    bool done = false;
    if (m_state == SOSBA_RES_OOB) { outcome = SOSBA_RES_OOB; ret_energy = old_energy; done = true; }

    if (!done) {
      const float *pc = a.precalc + (size_t)(host * a.nf + target) * SOSBA_PRECALC_FLOATS;
      const float4 q0 = __ldg((const float4 *)pc), q1 = __ldg((const float4 *)pc + 1), q2 = __ldg((const float4 *)pc + 2),
                   q3 = __ldg((const float4 *)pc + 3), q4 = __ldg((const float4 *)pc + 4), q5 = __ldg((const float4 *)pc + 5),
                   q6 = __ldg((const float4 *)pc + 6);
      const float R0[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
      const float t0[3] = {q2.y, q2.z, q2.w};
      const float KRKi[9] = {q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w, q5.x};
      const float Kt[3] = {q5.y, q5.z, q5.w};
      const float affLL0 = q6.x, affLL1 = q6.y, b0 = q6.z;
      const float fxl = a.calib[0], fyl = a.calib[1], cxl = a.calib[2], cyl = a.calib[3], fxli = a.calib[4], fyli = a.calib[5];
      const float pu = a.p_u[pt], pv = a.p_v[pt], idepth = a.p_idepth[pt] * kSCALE_IDEPTH, idepth_zero = a.p_idepth_zero[pt] * kSCALE_IDEPTH;
      const float color = a.p_color[(size_t)pt * 8 + idx], weight = a.p_weights[(size_t)pt * 8 + idx];

      float d_xi_x[6], d_xi_y[6], d_C_x[4], d_C_y[4], d_d_x, d_d_y;
      float cKu = 0.f, cKv = 0.f, c_new_idepth = 0.f;
      {  // centre projection at the evaluation point (Residuals.cpp:102-169)
        float KliP[3] = {(pu + 0 - cxl) * fxli, (pv + 0 - cyl) * fyli, 1.f};
        float ptp[3];
#pragma unroll
        for (int i = 0; i < 3; i++) ptp[i] = ((R0[3 * i] * KliP[0] + R0[3 * i + 1] * KliP[1]) + R0[3 * i + 2] * KliP[2]) + t0[i] * idepth_zero;
        float drescale = 1.0f / ptp[2];
        float new_idepth = idepth_zero * drescale;
        float u = ptp[0] * drescale, v = ptp[1] * drescale;
        float Ku = u * fxl + cxl, Kv = v * fyl + cyl;
        if (!(drescale > 0) || !(Ku > 1.1f && Kv > 1.1f && Ku < a.wM3G && Kv < a.hM3G)) {
          outcome = SOSBA_RES_OOB; ret_energy = old_energy; done = true;
        }
        cKu = Ku; cKv = Kv; c_new_idepth = new_idepth;
        d_d_x = drescale * (t0[0] - t0[2] * u) * kSCALE_IDEPTH * fxl;
        d_d_y = drescale * (t0[1] - t0[2] * v) * kSCALE_IDEPTH * fyl;

        d_C_x[2] = drescale * (R0[6] * u - R0[0]);
        d_C_x[3] = fxl * drescale * (R0[7] * u - R0[1]) * fyli;
        d_C_x[0] = KliP[0] * d_C_x[2];
        d_C_x[1] = KliP[1] * d_C_x[3];

        d_C_y[2] = fyl * drescale * (R0[6] * v - R0[3]) * fxli;
        d_C_y[3] = drescale * (R0[7] * v - R0[4]);
        d_C_y[0] = KliP[0] * d_C_y[2];
        d_C_y[1] = KliP[1] * d_C_y[3];

        d_C_x[0] = (d_C_x[0] + u) * kSCALE_F;
        d_C_x[1] *= kSCALE_F;
        d_C_x[2] = (d_C_x[2] + 1) * kSCALE_C;
        d_C_x[3] *= kSCALE_C;

        d_C_y[0] *= kSCALE_F;
        d_C_y[1] = (d_C_y[1] + v) * kSCALE_F;
        d_C_y[2] *= kSCALE_C;
        d_C_y[3] = (d_C_y[3] + 1) * kSCALE_C;

        d_xi_x[0] = new_idepth * fxl;
        d_xi_x[1] = 0;
        d_xi_x[2] = -new_idepth * u * fxl;
        d_xi_x[3] = -u * v * fxl;
        d_xi_x[4] = (1 + u * u) * fxl;
        d_xi_x[5] = -v * fxl;

        d_xi_y[0] = 0;
        d_xi_y[1] = new_idepth * fyl;
        d_xi_y[2] = -new_idepth * v * fyl;
        d_xi_y[3] = -(1 + v * v) * fyl;
        d_xi_y[4] = u * v * fyl;
        d_xi_y[5] = u * fyl;
      }

      if (!done) {
        // pattern sample idx at the current state (Residuals.cpp:177-256)
        const float up = pu + c_pattern[idx][0], vp = pv + c_pattern[idx][1];
        float ptp[3];
#pragma unroll
        for (int i = 0; i < 3; i++) ptp[i] = ((KRKi[3 * i] * up + KRKi[3 * i + 1] * vp) + KRKi[3 * i + 2] * 1.0f) + Kt[i] * idepth;
        const float Ku = ptp[0] / ptp[2], Kv = ptp[1] / ptp[2];
        bool bad = !(Ku > 1.1f && Kv > 1.1f && Ku < a.wM3G && Kv < a.hM3G);
        float3 hit = make_float3(0.f, 0.f, 0.f);
        if (!bad) {
          hit = interp33(a.img[target], Ku, Kv, a.w);
          if (!isfinite(hit.x)) bad = true;
        }
        if (__any_sync(gmask, bad)) {
          outcome = SOSBA_RES_OOB; ret_energy = old_energy; done = true;
        } else {
          const float residual = hit.x - (float)(affLL0 * color + affLL1);
          const float drdA = (color - b0);
          float w = sqrtf(a.outlierTHSum / (a.outlierTHSum + (hit.y * hit.y + hit.z * hit.z)));
          w = 0.5f * (w + weight);
          float hw = fabsf(residual) < a.huberTH ? 1 : a.huberTH / fabsf(residual);
          const float e_term = w * w * hw * residual * residual * (2 - hw);
          if (hw < 1) hw = sqrtf(hw);
          hw = hw * w;
          const float hx = hit.y * hw, hy = hit.z * hw;
          const float resF = residual * hw;
          float jab0 = drdA * hw, jab1 = hw;

          const float energyLeft0 = seqsum8(gmask, e_term);
          const float JIdxJIdx_00 = seqsum8(gmask, hx * hx);
          const float JIdxJIdx_11 = seqsum8(gmask, hy * hy);
          const float JIdxJIdx_10 = seqsum8(gmask, hx * hy);
          const float JabJIdx_00 = seqsum8(gmask, drdA * hw * hx);
          const float JabJIdx_01 = seqsum8(gmask, drdA * hw * hy);
          const float JabJIdx_10 = seqsum8(gmask, hw * hx);
          const float JabJIdx_11 = seqsum8(gmask, hw * hy);
          const float JabJab_00 = seqsum8(gmask, drdA * drdA * hw * hw);
          const float JabJab_01 = seqsum8(gmask, drdA * hw * hw);
          const float JabJab_11 = seqsum8(gmask, hw * hw);
          const float wJI2_sum = seqsum8(gmask, hw * hw * (hx * hx + hy * hy));
          if (a.affModeA < 0) jab0 = 0;
          if (a.affModeB < 0) jab1 = 0;

          // candidate record: PointFrameResidual::J
          if (WRITE_J) {
          float *J = (a.r_sel[r] ? a.J1 : a.J0) + (size_t)r * SOSBA_JREC;
          J[JR_RES + idx] = resF;
          J[JR_JIDX0 + idx] = hx;
          J[JR_JIDX1 + idx] = hy;
          J[JR_JAB0 + idx] = jab0;
          J[JR_JAB1 + idx] = jab1;
          {
            const float4 g0 = make_float4(d_xi_x[0], d_xi_x[1], d_xi_x[2], d_xi_x[3]);
            const float4 g1 = make_float4(d_xi_x[4], d_xi_x[5], d_xi_y[0], d_xi_y[1]);
            const float4 g2 = make_float4(d_xi_y[2], d_xi_y[3], d_xi_y[4], d_xi_y[5]);
            const float4 g3 = make_float4(d_C_x[0], d_C_x[1], d_C_x[2], d_C_x[3]);
            const float4 g4 = make_float4(d_C_y[0], d_C_y[1], d_C_y[2], d_C_y[3]);
            const float4 g5 = make_float4(d_d_x, d_d_y, JIdxJIdx_00, JIdxJIdx_10);
            const float4 g6 = make_float4(JIdxJIdx_10, JIdxJIdx_11, JabJIdx_00, JabJIdx_01);
            const float4 g7 = make_float4(JabJIdx_10, JabJIdx_11, JabJab_00, JabJab_01);
            ((float4 *)(J + JR_GEO))[idx] = pick4(idx, g0, g1, g2, g3, g4, g5, g6, g7);
            if (idx == 0) J[JR_JAB2 + 2] = JabJab_01;
            if (idx == 1) J[JR_JAB2 + 3] = JabJab_11;
          }
          a.proj[(size_t)r * 16 + idx * 2] = Ku;
          a.proj[(size_t)r * 16 + idx * 2 + 1] = Kv;
          }

          energyWO = energyLeft0;
          float energyLeft = energyLeft0;
          const float th = fmaxf(a.frameEnergyTH[host], a.frameEnergyTH[target]);
          if (energyLeft > th || wJI2_sum < 2) { energyLeft = th; outcome = SOSBA_RES_OUTLIER; }
          else outcome = SOSBA_RES_IN;
          ret_energy = energyLeft;
          if (APPLY && outcome == SOSBA_RES_IN) {   // EFResidual::takeDataF on the registers of this linearisation
            const float JI_r0 = bfly8(gmask, resF * hx), JI_r1 = bfly8(gmask, resF * hy);
            const float Jab_r0 = bfly8(gmask, resF * jab0), Jab_r1 = bfly8(gmask, resF * jab1), rr = bfly8(gmask, resF * resF);
            float *rec = a.rec + (size_t)r * SOSBA_CREC;
            if (idx == 0) *(float4 *)(rec + CR_X) = make_float4(d_C_x[0], d_C_x[1], d_C_x[2], d_C_x[3]);
            if (idx == 1) { *(float4 *)(rec + CR_X + 4) = make_float4(d_xi_x[0], d_xi_x[1], d_xi_x[2], d_xi_x[3]); *(float2 *)(rec + CR_X + 8) = make_float2(d_xi_x[4], d_xi_x[5]); }
            if (idx == 2) { *(float2 *)(rec + CR_Y) = make_float2(d_C_y[0], d_C_y[1]); *(float2 *)(rec + CR_Y + 2) = make_float2(d_C_y[2], d_C_y[3]); }
            if (idx == 3) { *(float2 *)(rec + CR_Y + 4) = make_float2(d_xi_y[0], d_xi_y[1]); *(float4 *)(rec + CR_Y + 6) = make_float4(d_xi_y[2], d_xi_y[3], d_xi_y[4], d_xi_y[5]); }
            if (idx == 4) {
              *(float4 *)(rec + CR_A) = make_float4(JIdxJIdx_00, JIdxJIdx_10, JIdxJIdx_11, JabJIdx_00);
              *(float4 *)(rec + CR_TR + 1) = make_float4(JabJIdx_01, JabJIdx_10, JabJIdx_11, JI_r0);
              rec[CR_TR + 5] = JI_r1;
            }
            if (idx == 5) {
              rec[CR_BR + 0] = JabJab_00; rec[CR_BR + 1] = JabJab_01; rec[CR_BR + 2] = Jab_r0;
              *(float4 *)(rec + CR_BR + 3) = make_float4(JabJab_11, Jab_r1, rr, d_d_x);
              rec[CR_JPDD + 1] = d_d_y;
            }
            const float v0 = JIdxJIdx_00 * d_d_x + JIdxJIdx_10 * d_d_y, v1 = JIdxJIdx_10 * d_d_x + JIdxJIdx_11 * d_d_y;
            float jp;
            if (idx < 6) jp = (idx == 0 ? d_xi_x[0] : idx == 1 ? d_xi_x[1] : idx == 2 ? d_xi_x[2] : idx == 3 ? d_xi_x[3] : idx == 4 ? d_xi_x[4] : d_xi_x[5]) * v0 +
                              (idx == 0 ? d_xi_y[0] : idx == 1 ? d_xi_y[1] : idx == 2 ? d_xi_y[2] : idx == 3 ? d_xi_y[3] : idx == 4 ? d_xi_y[4] : d_xi_y[5]) * v1;
            else if (idx == 6) jp = JabJIdx_00 * d_d_x + JabJIdx_01 * d_d_y;
            else jp = JabJIdx_10 * d_d_x + JabJIdx_11 * d_d_y;
            rec[CR_JPJDF + idx] = jp;
          }
          if (idx == 0) {
            a.r_new_energy[r] = energyLeft;
            a.center[(size_t)r * 3 + 0] = cKu; a.center[(size_t)r * 3 + 1] = cKv; a.center[(size_t)r * 3 + 2] = c_new_idepth;
            if (target == a.nf - 1) {  // input of setNewFrameEnergyTH
              int pos = atomicAdd(a.newE_count, 1);
              a.newE[pos] = energyWO;
            }
          }
        }
      }
    }
    if (APPLY && WRITE_J) __syncwarp(gmask);   // every lane has used r_sel before lane 0 flips it
    if (idx == 0) {
      a.r_new_state[r] = (uint8_t)outcome;
      a.r_new_energy_wo[r] = energyWO;
      if (APPLY && m_state != SOSBA_RES_OOB) {   // applyRes(true): "can never go back from OOB" (Residuals.cpp:306-309)
        a.r_is_active[r] = outcome == SOSBA_RES_IN ? 1 : 0;
        a.r_state[r] = (uint8_t)outcome;
        // state_energy = state_NewEnergy, which a linearisation that left early (new state OOB) did not refresh
        a.r_energy[r] = outcome == SOSBA_RES_OOB ? a.r_new_energy[r] : ret_energy;
        if (WRITE_J && outcome == SOSBA_RES_IN) a.r_sel[r] ^= 1;   // std::swap(J, data->J)
      }
    }
  }
  {  // energy and state histogram: 4 residual leaders per warp -> warp sum -> CTA sum (fixed order) -> one atomic each
    const bool lead = live && idx == 0;
    double e = lead ? (double)ret_energy : 0.0;
    int c0 = lead && outcome == 0, c1 = lead && outcome == 1, c2 = lead && outcome == 2;
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
      e += __shfl_xor_sync(0xffffffffu, e, o);
      c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if ((threadIdx.x & 31) == 0) { const int w = threadIdx.x >> 5; s_we[w] = e; s_wc[w][0] = c0; s_wc[w][1] = c1; s_wc[w][2] = c2; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0.0;
    int c[3] = {0, 0, 0};
    for (int w = 0; w < 8; w++) { e += s_we[w]; c[0] += s_wc[w][0]; c[1] += s_wc[w][1]; c[2] += s_wc[w][2]; }
    if (e != 0.0) atomicAdd(&a.stats[0], e);
    if (c[0]) atomicAdd(&a.counts[0], c[0]);
    if (c[1]) atomicAdd(&a.counts[1], c[1]);
    if (c[2]) atomicAdd(&a.counts[2], c[2]);
  }
  if (TH) {   // last CTA done: the newest-frame energies of every CTA are in place
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(a.ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      energy_th_body(a.th);
      if (threadIdx.x == 0) *a.ticket = 0;
    }
  }
}

// 128-bit read-only load that the compiler may not reorder against its siblings: a thread's taps stay back to back
__device__ __forceinline__ float4 ldg_nc_v4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// a3  LANES (1, 2 or 4) THREADS PER RESIDUAL, 8 / LANES pattern samples each (the default since round 2; the
// 8-lanes-per-residual kernel above stays selectable with SOSBA_LIN_LANES=8).  A lane walks its samples like the scalar
// loop of Residuals.cpp:177-256; the 2x2 sums are in-order float additions: lane 0 sums its samples from zero, every
// following lane continues the running sums of its left neighbour (one shuffle per sum and hop), so the result is
// bit-identical to the reference's sequential loop for every LANES.  LANES = 1: no shuffles at all, the centre
// projection and the geometric derivatives once per residual, 32 independent 128-bit taps in flight per thread —
// 7x fewer warp instructions than the 8-lane kernel: the variant for large windows (bandwidth regime).  LANES = 4: a
// quarter of the dependent instruction chain per thread: the variant for the 10^4-residual window of the benchmark,
// where the launch is one partial wave and its duration is the length of one thread's chain.
// The last lane of a group decides the state and writes the records (12 x 16 B per commit record).
// FIX: the bookkeeping of linearizeAll(true) (FullSystemOptimize.cpp:52-74, :148-179) fused in: maxRelBaseline / numGood of
//      new residuals, removal of residuals that did not stay active.
constexpr int LIN_T = 64;

template <int LANES, bool APPLY, bool WRITE_J, bool FIX>
__global__ void __launch_bounds__(LIN_T) k_linearize_t(LinArgs a) {
  constexpr int SPL = 8 / LANES;           // samples per lane
  constexpr int NS = APPLY ? 17 : 12;      // running sums
  __shared__ double s_we[LIN_T / 32];
  __shared__ int s_wc[LIN_T / 32][4];
  PDL_ENTER_T(a.trace);
  const int gid = blockIdx.x * LIN_T + threadIdx.x;
  const int r = gid / LANES, sub = gid % LANES;
  const int lane = threadIdx.x & 31;
  const unsigned gmask = LANES == 1 ? (1u << lane) : (((1u << LANES) - 1u) << (lane & ~(LANES - 1)));
  const bool writer = sub == LANES - 1;
  const bool inr = r < a.R;
  const int gate = a.gate ? *a.gate : 0;
  // first round trip: every per-residual id / flag at once (independent loads)
  const int m_lin = inr ? a.r_is_lin[r] : 1, m_drop = inr ? a.r_dropped[r] : 1, m_state = inr ? a.r_state[r] : 0;
  const int pt = inr ? a.r_point[r] : 0, host = inr ? a.r_host[r] : 0, target = inr ? a.r_target[r] : 0;
  const float old_energy = inr ? a.r_energy[r] : 0.f;
  if (gate) return;   // the Gauss-Newton loop already converged (device-side break)
  if (a.zero_n > 0) {   // clear the block tables of the accumulation that follows
    double2 *zb = (double2 *)a.zero_buf;
    for (int i = gid; i < a.zero_n; i += gridDim.x * LIN_T) zb[i] = make_double2(0.0, 0.0);
  }
#define LIN_TS(n, dep) do { if (a.trace && threadIdx.x == 0 && blockIdx.x == gridDim.x / 2) a.trace[4 + (n)] = clock64() + (long long)((dep) == 1.2345e-30f); } while (0)
  LIN_TS(0, old_energy);
  const bool live = inr && !(m_lin | m_drop);
  int outcome = -1;        // state_NewState once decided
  float ret_energy = 0.f;  // linearize() return value
  float energyWO = -1.f;
  int removed = 0;
  // Straight-line code with group-uniform stage predicates instead of nested early returns: the votes and the shuffles of
  // the running sums sit at convergent points of the warp (plain SHFL / VOTE, no per-shuffle reconvergence scaffolding).
  // Lanes whose residual left early carry harmless garbage through the arithmetic; every memory access is guarded.
  const bool s1 = live && m_state != SOSBA_RES_OOB;
  if (live) { outcome = SOSBA_RES_OOB; ret_energy = old_energy; }   // every early return of the reference leaves exactly this
  const float *pc = a.precalc + (size_t)(host * a.nf + target) * SOSBA_PRECALC_FLOATS;
  const float4 q0 = __ldg((const float4 *)pc), q1 = __ldg((const float4 *)pc + 1), q2 = __ldg((const float4 *)pc + 2),
               q3 = __ldg((const float4 *)pc + 3), q4 = __ldg((const float4 *)pc + 4), q5 = __ldg((const float4 *)pc + 5),
               q6 = __ldg((const float4 *)pc + 6);
  const float R0[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
  const float t0[3] = {q2.y, q2.z, q2.w};
  const float KRKi[9] = {q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w, q5.x};
  const float Kt[3] = {q5.y, q5.z, q5.w};
  const float affLL0 = q6.x, affLL1 = q6.y, b0 = q6.z;
  const float fxl = a.calib[0], fyl = a.calib[1], cxl = a.calib[2], cyl = a.calib[3], fxli = a.calib[4], fyli = a.calib[5];
  const float pu = a.p_u[pt], pv = a.p_v[pt], idepth = a.p_idepth[pt] * kSCALE_IDEPTH, idepth_zero = a.p_idepth_zero[pt] * kSCALE_IDEPTH;
  float color[SPL], weight[SPL];
  if (SPL == 8) {
    const float4 col0 = __ldg((const float4 *)(a.p_color + (size_t)pt * 8)), col1 = __ldg((const float4 *)(a.p_color + (size_t)pt * 8) + 1);
    const float4 wgt0 = __ldg((const float4 *)(a.p_weights + (size_t)pt * 8)), wgt1 = __ldg((const float4 *)(a.p_weights + (size_t)pt * 8) + 1);
    const float c8[8] = {col0.x, col0.y, col0.z, col0.w, col1.x, col1.y, col1.z, col1.w};
    const float w8[8] = {wgt0.x, wgt0.y, wgt0.z, wgt0.w, wgt1.x, wgt1.y, wgt1.z, wgt1.w};
#pragma unroll
    for (int k = 0; k < SPL; k++) { color[k] = c8[k % 8]; weight[k] = w8[k % 8]; }
  } else {
#pragma unroll
    for (int k = 0; k < SPL; k++) { color[k] = __ldg(a.p_color + (size_t)pt * 8 + sub * SPL + k); weight[k] = __ldg(a.p_weights + (size_t)pt * 8 + sub * SPL + k); }
  }
  const float th = fmaxf(a.frameEnergyTH[host], a.frameEnergyTH[target]);
  int sel = 0;
  if (WRITE_J) sel = inr ? a.r_sel[r] : 0;

  // centre projection at the evaluation point (Residuals.cpp:102-169); every lane of the group evaluates it.  Only the
  // bounds test is needed before the taps go out; the derivatives follow below, while the taps are in flight.
  const float KliP[3] = {(pu + 0 - cxl) * fxli, (pv + 0 - cyl) * fyli, 1.f};
  float cptp[3];
#pragma unroll
  for (int i = 0; i < 3; i++) cptp[i] = ((R0[3 * i] * KliP[0] + R0[3 * i + 1] * KliP[1]) + R0[3 * i + 2] * KliP[2]) + t0[i] * idepth_zero;
  const float drescale = 1.0f / cptp[2];
  const float c_new_idepth = idepth_zero * drescale;
  const float cu = cptp[0] * drescale, cv = cptp[1] * drescale;
  const float cKu = cu * fxl + cxl, cKv = cv * fyl + cyl;
  const bool centre_ok = (drescale > 0) && (cKu > 1.1f && cKv > 1.1f && cKu < a.wM3G && cKv < a.hM3G);
  const bool s2 = s1 && centre_ok;   // uniform across the lanes of a group
  LIN_TS(1, cKu + th);

  // this lane's pattern samples at the current state (Residuals.cpp:177-256): all projections first, then all taps
  float Ku[SPL], Kv[SPL];
  bool bad = false;
#pragma unroll
  for (int k = 0; k < SPL; k++) {
    const int idx = sub * SPL + k;
    const float up = pu + c_pattern[idx][0], vp = pv + c_pattern[idx][1];
    float ptp[3];
#pragma unroll
    for (int i = 0; i < 3; i++) ptp[i] = ((KRKi[3 * i] * up + KRKi[3 * i + 1] * vp) + KRKi[3 * i + 2] * 1.0f) + Kt[i] * idepth;
    Ku[k] = ptp[0] / ptp[2]; Kv[k] = ptp[1] / ptp[2];
    bad |= !(Ku[k] > 1.1f && Kv[k] > 1.1f && Ku[k] < a.wM3G && Kv[k] < a.hM3G);
  }
  if (LANES > 1) bad = (__ballot_sync(0xffffffffu, bad) & gmask) != 0u;
  const bool s3 = s2 && !bad;
  LIN_TS(2, Ku[0] + Kv[SPL - 1]);
  float d_xi_x[6], d_xi_y[6], d_C_x[4], d_C_y[4], d_d_x, d_d_y;
  float3 hit[SPL];
  {   // getInterpolatedElement33 (globalFuncs.h:68-82) of all samples: the taps go out back to back
    float4 t11[SPL], t01[SPL], t10[SPL], t00[SPL];
#pragma unroll
    for (int k = 0; k < SPL; k++) t11[k] = t01[k] = t10[k] = t00[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s3) {
      const float4 *img = a.img[target];
#pragma unroll
      for (int k = 0; k < SPL; k++) {
        const float4 *bp = img + (int)Ku[k] + (int)Kv[k] * a.w;
        t11[k] = ldg_nc_v4(bp + 1 + a.w); t01[k] = ldg_nc_v4(bp + a.w); t10[k] = ldg_nc_v4(bp + 1); t00[k] = ldg_nc_v4(bp);
      }
    }
    {  // the geometric derivatives (Residuals.cpp:121-169): independent of the taps, evaluated while they are in flight
      const float u = cu, v = cv, new_idepth = c_new_idepth;
      d_d_x = drescale * (t0[0] - t0[2] * u) * kSCALE_IDEPTH * fxl;
      d_d_y = drescale * (t0[1] - t0[2] * v) * kSCALE_IDEPTH * fyl;

      d_C_x[2] = drescale * (R0[6] * u - R0[0]);
      d_C_x[3] = fxl * drescale * (R0[7] * u - R0[1]) * fyli;
      d_C_x[0] = KliP[0] * d_C_x[2];
      d_C_x[1] = KliP[1] * d_C_x[3];

      d_C_y[2] = fyl * drescale * (R0[6] * v - R0[3]) * fxli;
      d_C_y[3] = drescale * (R0[7] * v - R0[4]);
      d_C_y[0] = KliP[0] * d_C_y[2];
      d_C_y[1] = KliP[1] * d_C_y[3];

      d_C_x[0] = (d_C_x[0] + u) * kSCALE_F;
      d_C_x[1] *= kSCALE_F;
      d_C_x[2] = (d_C_x[2] + 1) * kSCALE_C;
      d_C_x[3] *= kSCALE_C;

      d_C_y[0] *= kSCALE_F;
      d_C_y[1] = (d_C_y[1] + v) * kSCALE_F;
      d_C_y[2] *= kSCALE_C;
      d_C_y[3] = (d_C_y[3] + 1) * kSCALE_C;

      d_xi_x[0] = new_idepth * fxl;
      d_xi_x[1] = 0;
      d_xi_x[2] = -new_idepth * u * fxl;
      d_xi_x[3] = -u * v * fxl;
      d_xi_x[4] = (1 + u * u) * fxl;
      d_xi_x[5] = -v * fxl;

      d_xi_y[0] = 0;
      d_xi_y[1] = new_idepth * fyl;
      d_xi_y[2] = -new_idepth * v * fyl;
      d_xi_y[3] = -(1 + v * v) * fyl;
      d_xi_y[4] = u * v * fyl;
      d_xi_y[5] = u * fyl;
    }
    // every weight below depends on ALL taps (through a word that is zero at run time, or-ed from their unused 4th
    // components), so no consumer can be scheduled between the loads: all of them are in flight before the first wait
    unsigned zr = 0u;
#pragma unroll
    for (int k = 0; k < SPL; k++)
      zr |= __float_as_uint(t11[k].w) | __float_as_uint(t01[k].w) | __float_as_uint(t10[k].w) | __float_as_uint(t00[k].w);
    zr &= a.opaque_zero;
    bad = false;
#pragma unroll
    for (int k = 0; k < SPL; k++) {
      const float x = __uint_as_float(__float_as_uint(Ku[k]) | zr), y = Kv[k];
      const int ix = (int)x, iy = (int)y;
      const float dx = x - ix, dy = y - iy;
      const float dxdy = dx * dy;
      const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
      hit[k].x = w11 * t11[k].x + w01 * t01[k].x + w10 * t10[k].x + w00 * t00[k].x;
      hit[k].y = w11 * t11[k].y + w01 * t01[k].y + w10 * t10[k].y + w00 * t00[k].y;
      hit[k].z = w11 * t11[k].z + w01 * t01[k].z + w10 * t10[k].z + w00 * t00[k].z;
      bad |= !isfinite(hit[k].x);
    }
    if (LANES > 1) bad = (__ballot_sync(0xffffffffu, bad) & gmask) != 0u;
  }
  const bool reached = s3 && !bad;   // the energies are evaluated (no early OOB return)
  LIN_TS(3, hit[0].x + hit[SPL - 1].z);
  // the terms of the running sums, sample by sample: 0 energyLeft, 1..3 JIdx2 (00, 11, 10), 4..7 JabJIdx (00, 01, 10, 11),
  // 8..10 Jab2 (00, 01, 11), 11 wJI2_sum, 12..16 JI_r[0], JI_r[1], Jab_r[0], Jab_r[1], rr (EFResidual::takeDataF / addPoint)
  float term[NS][SPL];
  float *J = nullptr;
  if (WRITE_J) J = (sel ? a.J1 : a.J0) + (size_t)r * SOSBA_JREC;
#pragma unroll
  for (int k = 0; k < SPL; k++) {
    const int idx = sub * SPL + k;
    const float3 hc = hit[k];
    const float residual = hc.x - (float)(affLL0 * color[k] + affLL1);
    const float drdA = (color[k] - b0);
    float w = sqrtf(a.outlierTHSum / (a.outlierTHSum + (hc.y * hc.y + hc.z * hc.z)));
    w = 0.5f * (w + weight[k]);
    float hw = fabsf(residual) < a.huberTH ? 1 : a.huberTH / fabsf(residual);
    term[0][k] = w * w * hw * residual * residual * (2 - hw);
    if (hw < 1) hw = sqrtf(hw);
    hw = hw * w;
    const float hx = hc.y * hw, hy = hc.z * hw;
    const float resF = residual * hw;
    float jab0 = drdA * hw, jab1 = hw;
    term[1][k] = hx * hx;
    term[2][k] = hy * hy;
    term[3][k] = hx * hy;
    term[4][k] = drdA * hw * hx;
    term[5][k] = drdA * hw * hy;
    term[6][k] = hw * hx;
    term[7][k] = hw * hy;
    term[8][k] = drdA * drdA * hw * hw;
    term[9][k] = drdA * hw * hw;
    term[10][k] = hw * hw;
    term[11][k] = hw * hw * (hx * hx + hy * hy);
    if (a.affModeA < 0) jab0 = 0;
    if (a.affModeB < 0) jab1 = 0;
    if (APPLY) {
      term[12 % NS][k] = resF * hx; term[13 % NS][k] = resF * hy;
      term[14 % NS][k] = resF * jab0; term[15 % NS][k] = resF * jab1;
      term[16 % NS][k] = resF * resF;
    }
    if (WRITE_J && reached) {
      J[JR_RES + idx] = resF;
      J[JR_JIDX0 + idx] = hx;
      J[JR_JIDX1 + idx] = hy;
      J[JR_JAB0 + idx] = jab0;
      J[JR_JAB1 + idx] = jab1;
      a.proj[(size_t)r * 16 + idx * 2] = Ku[k];
      a.proj[(size_t)r * 16 + idx * 2 + 1] = Kv[k];
    }
  }
  // in-order sums: lane 0 from zero, lane h continues lane h-1 (whole-warp shuffles at a convergent point)
  float S[NS];
#pragma unroll
  for (int q = 0; q < NS; q++) {
    float v = 0;
#pragma unroll
    for (int k = 0; k < SPL; k++) v += term[q][k];
    S[q] = v;
  }
#pragma unroll
  for (int hop = 1; hop < LANES; hop++) {
#pragma unroll
    for (int q = 0; q < NS; q++) {
      float v = __shfl_up_sync(0xffffffffu, S[q], 1);
#pragma unroll
      for (int k = 0; k < SPL; k++) v += term[q][k];
      if (sub == hop) S[q] = v;
    }
  }
  LIN_TS(4, S[0] + S[NS - 1]);
  if (writer && reached) {
    const float JIdxJIdx_00 = S[1], JIdxJIdx_11 = S[2], JIdxJIdx_10 = S[3];
    const float JabJIdx_00 = S[4], JabJIdx_01 = S[5], JabJIdx_10 = S[6], JabJIdx_11 = S[7];
    const float JabJab_00 = S[8], JabJab_01 = S[9], JabJab_11 = S[10], wJI2_sum = S[11];
    if (WRITE_J) {   // candidate record: PointFrameResidual::J
      float4 *G = (float4 *)(J + JR_GEO);
      G[0] = make_float4(d_xi_x[0], d_xi_x[1], d_xi_x[2], d_xi_x[3]);
      G[1] = make_float4(d_xi_x[4], d_xi_x[5], d_xi_y[0], d_xi_y[1]);
      G[2] = make_float4(d_xi_y[2], d_xi_y[3], d_xi_y[4], d_xi_y[5]);
      G[3] = make_float4(d_C_x[0], d_C_x[1], d_C_x[2], d_C_x[3]);
      G[4] = make_float4(d_C_y[0], d_C_y[1], d_C_y[2], d_C_y[3]);
      G[5] = make_float4(d_d_x, d_d_y, JIdxJIdx_00, JIdxJIdx_10);
      G[6] = make_float4(JIdxJIdx_10, JIdxJIdx_11, JabJIdx_00, JabJIdx_01);
      G[7] = make_float4(JabJIdx_10, JabJIdx_11, JabJab_00, JabJab_01);
      *(float2 *)(J + JR_JAB2 + 2) = make_float2(JabJab_01, JabJab_11);
    }
    float energyLeft = S[0];
    energyWO = energyLeft;
    if (energyLeft > th || wJI2_sum < 2) { energyLeft = th; outcome = SOSBA_RES_OUTLIER; }
    else outcome = SOSBA_RES_IN;
    ret_energy = energyLeft;
    if (APPLY && outcome == SOSBA_RES_IN) {   // EFResidual::takeDataF on the registers of this linearisation
      const float JI_r0 = S[12 % NS], JI_r1 = S[13 % NS], Jab_r0 = S[14 % NS], Jab_r1 = S[15 % NS], rr = S[16 % NS];
      float4 *rec = (float4 *)(a.rec + (size_t)r * SOSBA_CREC);
      const float v0 = JIdxJIdx_00 * d_d_x + JIdxJIdx_10 * d_d_y, v1 = JIdxJIdx_10 * d_d_x + JIdxJIdx_11 * d_d_y;
      float jp[8];
#pragma unroll
      for (int i = 0; i < 6; i++) jp[i] = d_xi_x[i] * v0 + d_xi_y[i] * v1;
      jp[6] = JabJIdx_00 * d_d_x + JabJIdx_01 * d_d_y;
      jp[7] = JabJIdx_10 * d_d_x + JabJIdx_11 * d_d_y;
      rec[0] = make_float4(d_C_x[0], d_C_x[1], d_C_x[2], d_C_x[3]);            // CR_X: Jpdc[0], Jpdxi[0]
      rec[1] = make_float4(d_xi_x[0], d_xi_x[1], d_xi_x[2], d_xi_x[3]);
      rec[2] = make_float4(d_xi_x[4], d_xi_x[5], d_C_y[0], d_C_y[1]);          // CR_Y = 10: Jpdc[1], Jpdxi[1]
      rec[3] = make_float4(d_C_y[2], d_C_y[3], d_xi_y[0], d_xi_y[1]);
      rec[4] = make_float4(d_xi_y[2], d_xi_y[3], d_xi_y[4], d_xi_y[5]);
      rec[5] = make_float4(JIdxJIdx_00, JIdxJIdx_10, JIdxJIdx_11, JabJIdx_00);  // CR_A = 20, CR_TR = 23
      rec[6] = make_float4(JabJIdx_01, JabJIdx_10, JabJIdx_11, JI_r0);
      rec[7] = make_float4(JI_r1, JabJab_00, JabJab_01, Jab_r0);               // CR_BR = 29
      rec[8] = make_float4(JabJab_11, Jab_r1, rr, d_d_x);                      // CR_JPDD = 35
      rec[9] = make_float4(d_d_y, 0.f, 0.f, 0.f);
      rec[10] = make_float4(jp[0], jp[1], jp[2], jp[3]);                       // CR_JPJDF = 40
      rec[11] = make_float4(jp[4], jp[5], jp[6], jp[7]);
    }
    a.r_new_energy[r] = energyLeft;
    a.center[(size_t)r * 3 + 0] = cKu; a.center[(size_t)r * 3 + 1] = cKv; a.center[(size_t)r * 3 + 2] = c_new_idepth;
  }
  if (live) {
    if (writer) {   // (r_sel was read by every lane of the group at the top, before the first vote)
      a.r_new_state[r] = (uint8_t)outcome;
      a.r_new_energy_wo[r] = energyWO;
      if (APPLY) {
        if (m_state != SOSBA_RES_OOB) {   // applyRes(true): "can never go back from OOB" (Residuals.cpp:306-309)
          const bool active = outcome == SOSBA_RES_IN;
          a.r_is_active[r] = active ? 1 : 0;
          a.r_state[r] = (uint8_t)outcome;
          // state_energy = state_NewEnergy, which a linearisation that left early (new state OOB) did not refresh
          a.r_energy[r] = outcome == SOSBA_RES_OOB ? a.r_new_energy[r] : ret_energy;
          if (WRITE_J && active) a.r_sel[r] ^= 1;   // std::swap(J, data->J)
          if (FIX) {
            if (active) {
              if (a.r_is_new[r]) {  // FullSystemOptimize.cpp:55-66
                const float *pc = a.precalc + (size_t)(host * a.nf + target) * SOSBA_PRECALC_FLOATS;
                const float *KRKi = pc + SOSBA_PC_KRKI, *Kt = pc + SOSBA_PC_KT;
                const float pu = a.p_u[pt], pv = a.p_v[pt], id = a.p_idepth[pt] * kSCALE_IDEPTH;
                float inf[3], ptp[3];
#pragma unroll
                for (int i = 0; i < 3; i++) inf[i] = (KRKi[3 * i] * pu + KRKi[3 * i + 1] * pv) + KRKi[3 * i + 2] * 1.0f;
#pragma unroll
                for (int i = 0; i < 3; i++) ptp[i] = inf[i] + Kt[i] * id;
                const float ex = inf[0] / inf[2] - ptp[0] / ptp[2], ey = inf[1] / inf[2] - ptp[1] / ptp[2];
                const float relBS = (float)(0.01 * (double)sqrtf(ex * ex + ey * ey));
                if (relBS > 0.f) atomicMax((int *)&a.p_maxRelBaseline[pt], __float_as_int(relBS));  // positive floats order like ints
                atomicAdd(&a.p_numGood[pt], 1);
              }
            } else { a.r_dropped[r] = 1; removed = 1; }  // toRemove -> ef->dropResidual (FullSystemOptimize.cpp:148-179)
          }
        } else if (FIX && !a.r_is_active[r]) { a.r_dropped[r] = 1; removed = 1; }
      }
    }
  }
  LIN_TS(5, ret_energy);
  {  // newest-frame energies (input of setNewFrameEnergyTH): one atomic per warp, unordered list
    const bool app = writer && reached && target == a.nf - 1;
    const unsigned bal = __ballot_sync(0xffffffffu, app);
    if (bal) {
      int base = 0;
      const int leader = __ffs(bal) - 1;
      if (lane == leader) base = atomicAdd(a.newE_count, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (app) a.newE[base + __popc(bal & ((1u << lane) - 1))] = energyWO;
    }
  }
  {  // energy and state histogram: warp sums -> CTA sums (fixed order) -> one atomic each
    const bool cnt = live && writer;
    double e = cnt ? (double)ret_energy : 0.0;
    int c0 = cnt && outcome == 0, c1 = cnt && outcome == 1, c2 = cnt && outcome == 2, c3 = removed;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e += __shfl_xor_sync(0xffffffffu, e, o);
      c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    }
    if (lane == 0) { const int w = threadIdx.x >> 5; s_we[w] = e; s_wc[w][0] = c0; s_wc[w][1] = c1; s_wc[w][2] = c2; s_wc[w][3] = c3; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0.0;
    int c[4] = {0, 0, 0, 0};
#pragma unroll
    for (int w = 0; w < LIN_T / 32; w++) { e += s_we[w]; c[0] += s_wc[w][0]; c[1] += s_wc[w][1]; c[2] += s_wc[w][2]; c[3] += s_wc[w][3]; }
    if (e != 0.0) atomicAdd(&a.stats[0], e);
    if (c[0]) atomicAdd(&a.counts[0], c[0]);
    if (c[1]) atomicAdd(&a.counts[1], c[1]);
    if (c[2]) atomicAdd(&a.counts[2], c[2]);
    if (FIX && c[3]) atomicAdd(&a.counts[3], c[3]);
  }
  LIN_TS(6, 0.f);
  TRACE_EXIT(a.trace);
}

__device__ __forceinline__ float bfly8(unsigned mask, float v) {
  v += __shfl_xor_sync(mask, v, 1, 8);
  v += __shfl_xor_sync(mask, v, 2, 8);
  v += __shfl_xor_sync(mask, v, 4, 8);
  return v;
}

// writes the 48-float commit record from a committed J record and the (JI_r, Jab_r, rr) of resApprox
__device__ __forceinline__ void write_commit_record(const float *__restrict__ J, float *__restrict__ rec, int idx, float JI_r0, float JI_r1,
                                                    float Jab_r0, float Jab_r1, float rr) {
  // lanes 0..5 copy x/y (20 floats, 4 per lane: lane k -> floats 4k..4k+3), lane 5 the 2x2 sums, lane 6/7 the rest
  const float *g = J + JR_GEO;  // Jpdxi0[6] Jpdxi1[6] Jpdc0[4] Jpdc1[4] Jpdd[2] JIdx2[4] JabJIdx[4] Jab2[4]
  const float Jpdd0 = g[20], Jpdd1 = g[21];
  const float a00 = g[22], a01 = g[23], a11 = g[25];
  if (idx == 0) { rec[CR_X + 0] = g[12]; rec[CR_X + 1] = g[13]; rec[CR_X + 2] = g[14]; rec[CR_X + 3] = g[15]; }          // Jpdc0
  if (idx == 1) { rec[CR_X + 4] = g[0]; rec[CR_X + 5] = g[1]; rec[CR_X + 6] = g[2]; rec[CR_X + 7] = g[3]; rec[CR_X + 8] = g[4]; rec[CR_X + 9] = g[5]; }
  if (idx == 2) { rec[CR_Y + 0] = g[16]; rec[CR_Y + 1] = g[17]; rec[CR_Y + 2] = g[18]; rec[CR_Y + 3] = g[19]; }          // Jpdc1
  if (idx == 3) { rec[CR_Y + 4] = g[6]; rec[CR_Y + 5] = g[7]; rec[CR_Y + 6] = g[8]; rec[CR_Y + 7] = g[9]; rec[CR_Y + 8] = g[10]; rec[CR_Y + 9] = g[11]; }
  if (idx == 4) {
    rec[CR_A + 0] = a00; rec[CR_A + 1] = a01; rec[CR_A + 2] = a11;
    rec[CR_TR + 0] = g[26]; rec[CR_TR + 1] = g[27]; rec[CR_TR + 2] = g[28]; rec[CR_TR + 3] = g[29]; rec[CR_TR + 4] = JI_r0; rec[CR_TR + 5] = JI_r1;
  }
  if (idx == 5) {
    rec[CR_BR + 0] = g[30]; rec[CR_BR + 1] = g[31]; rec[CR_BR + 2] = Jab_r0; rec[CR_BR + 3] = g[33]; rec[CR_BR + 4] = Jab_r1; rec[CR_BR + 5] = rr;
    rec[CR_JPDD + 0] = Jpdd0; rec[CR_JPDD + 1] = Jpdd1;
  }
  // EFResidual::takeDataF: JpJdF = [Jpdxi^T (JIdx2 Jpdd) ; JabJIdx Jpdd]
  const float v0 = a00 * Jpdd0 + a01 * Jpdd1, v1 = g[24] * Jpdd0 + a11 * Jpdd1;
  if (idx < 6) rec[CR_JPJDF + idx] = g[idx] * v0 + g[6 + idx] * v1;
  if (idx == 6) rec[CR_JPJDF + 6] = g[26] * Jpdd0 + g[27] * Jpdd1;
  if (idx == 7) rec[CR_JPJDF + 7] = g[28] * Jpdd0 + g[29] * Jpdd1;
}

// a5 (+ a4 bookkeeping when fix != 0): 8 lanes per residual
__global__ void __launch_bounds__(256) k_apply_res(LinArgs a, int fix) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = gid >> 3, idx = gid & 7;
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  if (r >= a.R) return;
  if (a.r_is_lin[r] | a.r_dropped[r]) return;
  if (a.r_state[r] == SOSBA_RES_OOB) {  // applyRes(true): "can never go back from OOB" — nothing changes
    if (fix && idx == 0 && !a.r_is_active[r]) { a.r_dropped[r] = 1; atomicAdd(&a.counts[3], 1); }
    return;
  }
  const int ns = a.r_new_state[r];
  bool active = false;
  if (ns == SOSBA_RES_IN) {
    active = true;
    const int sel = a.r_sel[r];
    const float *J = (sel ? a.J1 : a.J0) + (size_t)r * SOSBA_JREC;  // candidate becomes the committed record
    const float res = J[JR_RES + idx], jx = J[JR_JIDX0 + idx], jy = J[JR_JIDX1 + idx], ja = J[JR_JAB0 + idx], jb = J[JR_JAB1 + idx];
    const float JI_r0 = bfly8(gmask, res * jx), JI_r1 = bfly8(gmask, res * jy);
    const float Jab_r0 = bfly8(gmask, res * ja), Jab_r1 = bfly8(gmask, res * jb), rr = bfly8(gmask, res * res);
    write_commit_record(J, a.rec + (size_t)r * SOSBA_CREC, idx, JI_r0, JI_r1, Jab_r0, Jab_r1, rr);
    __syncwarp(gmask);
    if (idx == 0) a.r_sel[r] = (uint8_t)(sel ^ 1);  // std::swap(J, data->J)
  }
  if (idx == 0) {
    a.r_is_active[r] = active ? 1 : 0;
    a.r_state[r] = (uint8_t)ns;
    a.r_energy[r] = a.r_new_energy[r];
    if (fix) {
      if (active) {
        if (a.r_is_new[r]) {  // FullSystemOptimize.cpp:55-66
          const int pt = a.r_point[r];
          const float *pc = a.precalc + (size_t)(a.r_host[r] * a.nf + a.r_target[r]) * SOSBA_PRECALC_FLOATS;
          const float *KRKi = pc + SOSBA_PC_KRKI, *Kt = pc + SOSBA_PC_KT;
          const float pu = a.p_u[pt], pv = a.p_v[pt], id = a.p_idepth[pt] * kSCALE_IDEPTH;
          float inf[3], ptp[3];
          for (int i = 0; i < 3; i++) inf[i] = (KRKi[3 * i] * pu + KRKi[3 * i + 1] * pv) + KRKi[3 * i + 2] * 1.0f;
          for (int i = 0; i < 3; i++) ptp[i] = inf[i] + Kt[i] * id;
          const float ex = inf[0] / inf[2] - ptp[0] / ptp[2], ey = inf[1] / inf[2] - ptp[1] / ptp[2];
          const float relBS = (float)(0.01 * (double)sqrtf(ex * ex + ey * ey));
          if (relBS > 0.f) atomicMax((int *)&a.p_maxRelBaseline[pt], __float_as_int(relBS));  // positive floats order like ints
          atomicAdd(&a.p_numGood[pt], 1);
        }
      } else {
        a.r_dropped[r] = 1;  // toRemove -> ef->dropResidual (FullSystemOptimize.cpp:148-179)
        atomicAdd(&a.counts[3], 1);
      }
    }
  }
}

// PointFrameResidual::resetOOB (Residuals.h:81-86) over activeResiduals (FullSystemOptimize.cpp:316-329)
__global__ void k_reset_oob(LinArgs a, int *zero_words, int n_zero, int *zero_ctl) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0) {
    if (zero_words && (int)threadIdx.x < n_zero) zero_words[threadIdx.x] = 0;
    if (zero_ctl && threadIdx.x >= 32 && threadIdx.x < 36) zero_ctl[threadIdx.x - 32] = 0;
  }
  if (r >= a.R) return;
  if (a.r_is_lin[r] | a.r_dropped[r]) return;
  a.r_new_energy[r] = 0.f; a.r_energy[r] = 0.f;
  a.r_new_state[r] = SOSBA_RES_OUTLIER; a.r_state[r] = SOSBA_RES_IN;
}

// members of a freshly uploaded residual that follow from the uploaded ones (PointFrameResidual constructor + resetOOB
// defaults, Residuals.cpp:62-78): host frame of its point, state_NewEnergy = state_energy, state_NewEnergyWithOutlier = -1,
// state_NewState = OUTLIER, candidate record selector 0, not dropped
__global__ void k_residual_init(LinArgs a, const int *__restrict__ p_host) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.R) return;
  const_cast<int *>(a.r_host)[r] = p_host[a.r_point[r]];
  a.r_new_energy[r] = a.r_energy[r];
  a.r_new_energy_wo[r] = -1.f;
  a.r_new_state[r] = SOSBA_RES_OUTLIER;
  a.r_sel[r] = 0;
  a.r_dropped[r] = 0;
}

__device__ __forceinline__ float dot6(const float *x, const float *y) {
  float s = x[0] * y[0];
  for (int i = 1; i < 6; i++) s += x[i] * y[i];
  return s;
}
__device__ __forceinline__ float dot4(const float *x, const float *y) {
  float s = x[0] * y[0];
  for (int i = 1; i < 4; i++) s += x[i] * y[i];
  return s;
}

// a13  res_toZeroF = resF - J*delta ; isLinearized = true
__global__ void __launch_bounds__(256) k_fix_linearization(LinArgs a, const int *__restrict__ ids, int n) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = gid >> 3, idx = gid & 7;
  if (k >= n) return;
  const int r = ids[k];
  const float *J = (a.r_sel[r] ? a.J0 : a.J1) + (size_t)r * SOSBA_JREC;  // committed record = EFResidual::J
  const float *dp = a.adHTdeltaF + 8 * (size_t)(a.r_host[r] + a.nf * a.r_target[r]);
  const float *cD = a.calib + 6;
  const float deltaF = a.p_deltaF[a.r_point[r]];
  const float Jp_delta_x = dot6(J + JR_JPDXI0, dp) + dot4(J + JR_JPDC0, cD) + J[JR_JPDD] * deltaF;
  const float Jp_delta_y = dot6(J + JR_JPDXI1, dp) + dot4(J + JR_JPDC1, cD) + J[JR_JPDD + 1] * deltaF;
  float rtz = J[JR_RES + idx];
  rtz = rtz - J[JR_JIDX0 + idx] * Jp_delta_x;
  rtz = rtz - J[JR_JIDX1 + idx] * Jp_delta_y;
  rtz = rtz - J[JR_JAB0 + idx] * dp[6];
  rtz = rtz - J[JR_JAB1 + idx] * dp[7];
  a.rtz[(size_t)r * 8 + idx] = rtz;
  if (idx == 0) a.r_is_lin[r] = 1;
}

// commit records of linearised (mode 1) / to-be-marginalised (mode 2) residuals
__global__ void __launch_bounds__(256) k_prep_records(LinArgs a, int mode, const int *__restrict__ list, int n) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = gid >> 3, idx = gid & 7;
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  if (k >= n) return;
  const int r = list ? list[k] : k;
  if (a.r_dropped[r] || !a.r_is_active[r]) return;
  if (mode == 1 && !a.r_is_lin[r]) return;
  const float *J = (a.r_sel[r] ? a.J0 : a.J1) + (size_t)r * SOSBA_JREC;
  float res = a.rtz[(size_t)r * 8 + idx];
  const float jx = J[JR_JIDX0 + idx], jy = J[JR_JIDX1 + idx], ja = J[JR_JAB0 + idx], jb = J[JR_JAB1 + idx];
  if (mode == 1) {  // AccumulatedTopHessian.cpp:76-97
    const float *dp = a.adHTdeltaF + 8 * (size_t)(a.r_host[r] + a.nf * a.r_target[r]);
    const float *cD = a.calib + 6;
    const float dd = a.p_deltaF[a.r_point[r]];
    const float Jp_delta_x = dot6(J + JR_JPDXI0, dp) + dot4(J + JR_JPDC0, cD) + J[JR_JPDD] * dd;
    const float Jp_delta_y = dot6(J + JR_JPDXI1, dp) + dot4(J + JR_JPDC1, cD) + J[JR_JPDD + 1] * dd;
    res = res + jx * Jp_delta_x;
    res = res + jy * Jp_delta_y;
    res = res + ja * dp[6];
    res = res + jb * dp[7];
  }
  const float JI_r0 = bfly8(gmask, res * jx), JI_r1 = bfly8(gmask, res * jy);
  const float Jab_r0 = bfly8(gmask, res * ja), Jab_r1 = bfly8(gmask, res * jb), rr = bfly8(gmask, res * res);
  write_commit_record(J, a.rec + (size_t)r * SOSBA_CREC, idx, JI_r0, JI_r1, Jab_r0, Jab_r1, rr);
}

// ------------------------------------------------------------------------------------------------
// setNewFrameEnergyTH: energy_th.cuh
__global__ void __launch_bounds__(1024) k_energy_th(ThArgs a, const int *gate, int cache_words) {
  extern __shared__ unsigned s_cache[];   // lists beyond 8 energies per thread (point shards of many ranks) are staged here
  PDL_ENTER();
  if (gate && *gate) return;
  energy_th_body(a, s_cache, cache_words);
}

// ------------------------------------------------------------------------------------------------
// a14 / a17  one thread per reference point.  The warped buffers are written in place (index i) with
// weight 0 for rejected points instead of being compacted: calcGSSSE sums them, zero rows add nothing.
__device__ __forceinline__ float3 mul33(const float *M, float x, float y, float z) {
  return make_float3((M[0] * x + M[1] * y) + M[2] * z, (M[3] * x + M[4] * y) + M[5] * z, (M[6] * x + M[7] * y) + M[8] * z);
}

__global__ void __launch_bounds__(256) k_track_res(TrackResArgs a) {
  __shared__ float s_E, s_T, s_RT, s_N;
  __shared__ int s_c[3];
  if (threadIdx.x == 0) { s_E = s_T = s_RT = s_N = 0.f; s_c[0] = s_c[1] = s_c[2] = 0; }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float E = 0.f, sT = 0.f, sRT = 0.f, sN = 0.f;
  int inE = 0, inW = 0, sat = 0;
  if (i < a.n) {
    const float x = a.pc[i], y = a.pc[a.n + i], id = a.pc[2 * a.n + i], refColor = a.color[i];
    float3 pt;
    float rx0 = 0.f, rx1 = 0.f, rx2 = 0.f;
    if (a.kind == 2) {   // PoseEstimator::calcRes (LoopClosure/PoseEstimator.cpp:193-206): 3D point, R in a.RKi, id holds z
      pt = mul33(a.RKi, x, y, id);
    } else if (a.kind == 0) {
      pt = mul33(a.RKi, x, y, 1.f);
    } else {  // scale * RKi (ScaleOptimizer.cpp:296-297): the matrix entries are scaled first
      float sRKi[9];
      for (int k = 0; k < 9; k++) sRKi[k] = a.scale * a.RKi[k];
      pt = mul33(sRKi, x, y, 1.f);
      float3 rx = mul33(a.RKi, x, y, 1.f);
      rx0 = rx.x / id; rx1 = rx.y / id; rx2 = rx.z / id;
    }
    if (a.kind == 2) { pt.x = pt.x + a.t[0]; pt.y = pt.y + a.t[1]; pt.z = pt.z + a.t[2]; }
    else { pt.x = pt.x + a.t[0] * id; pt.y = pt.y + a.t[1] * id; pt.z = pt.z + a.t[2] * id; }
    const float u = pt.x / pt.z, v = pt.y / pt.z;
    const float Ku = a.fx * u + a.cx, Kv = a.fy * v + a.cy;
    const float new_idepth = a.kind == 2 ? 1 / pt.z : id / pt.z;
    if (a.kind == 2 && a.lvl == 0 && i % 32 == 0) {  // flow indicators of the loop-closure variant (PoseEstimator.cpp:206-238)
      const float Ku0 = a.fx * (x / id) + a.cx, Kv0 = a.fy * (y / id) + a.cy;
      const float3 ptT = make_float3(x + a.t[0], y + a.t[1], 1 + a.t[2]);       // sic: (x, y, 1), not the normalised point
      const float KuT = a.fx * (ptT.x / ptT.z) + a.cx, KvT = a.fy * (ptT.y / ptT.z) + a.cy;
      const float3 ptT2 = make_float3(x - a.t[0], y - a.t[1], 1 - a.t[2]);
      const float KuT2 = a.fx * (ptT2.x / ptT2.z) + a.cx, KvT2 = a.fy * (ptT2.y / ptT2.z) + a.cy;
      const float3 rp = mul33(a.RKi, x, y, 1.f);
      const float3 pt3 = make_float3(rp.x - a.t[0], rp.y - a.t[1], rp.z - a.t[2]);
      const float Ku3 = a.fx * (pt3.x / pt3.z) + a.cx, Kv3 = a.fy * (pt3.y / pt3.z) + a.cy;
      sT += (KuT - Ku0) * (KuT - Ku0) + (KvT - Kv0) * (KvT - Kv0);
      sT += (KuT2 - Ku0) * (KuT2 - Ku0) + (KvT2 - Kv0) * (KvT2 - Kv0);
      sRT += (Ku - Ku0) * (Ku - Ku0) + (Kv - Kv0) * (Kv - Kv0);
      sRT += (Ku3 - Ku0) * (Ku3 - Ku0) + (Kv3 - Kv0) * (Kv3 - Kv0);
      sN += 2;
    } else if (a.lvl == 0 && i % 32 == 0) {  // flow indicators (CoarseTracker.cpp:666-696)
      float sKi[9];
      for (int k = 0; k < 9; k++) sKi[k] = a.kind == 0 ? a.Ki[k] : a.scale * a.Ki[k];
      float3 kp = mul33(sKi, x, y, 1.f);
      float3 ptT = make_float3(kp.x + a.t[0] * id, kp.y + a.t[1] * id, kp.z + a.t[2] * id);
      float KuT = a.fx * (ptT.x / ptT.z) + a.cx, KvT = a.fy * (ptT.y / ptT.z) + a.cy;
      float3 ptT2 = make_float3(kp.x - a.t[0] * id, kp.y - a.t[1] * id, kp.z - a.t[2] * id);
      float KuT2 = a.fx * (ptT2.x / ptT2.z) + a.cx, KvT2 = a.fy * (ptT2.y / ptT2.z) + a.cy;
      float3 rp;
      if (a.kind == 0) rp = mul33(a.RKi, x, y, 1.f);
      else { float sRKi[9]; for (int k = 0; k < 9; k++) sRKi[k] = a.scale * a.RKi[k]; rp = mul33(sRKi, x, y, 1.f); }
      float3 pt3 = make_float3(rp.x - a.t[0] * id, rp.y - a.t[1] * id, rp.z - a.t[2] * id);
      float Ku3 = a.fx * (pt3.x / pt3.z) + a.cx, Kv3 = a.fy * (pt3.y / pt3.z) + a.cy;
      sT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sN += 2;
    }
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, o4 = 0.f, o5 = 0.f, o6 = 0.f, o7 = 0.f;
    if (Ku > 2 && Kv > 2 && Ku < a.w - 3 && Kv < a.h - 3 && new_idepth > 0) {
      float3 hit = interp33(a.img, Ku, Kv, a.w);
      if (isfinite(hit.x)) {
        const float residual = a.kind != 1 ? hit.x - (float)(a.aff0 * refColor + a.aff1) : hit.x - refColor;
        const float hw = fabsf(residual) < a.huberTH ? 1 : a.huberTH / fabsf(residual);
        if (fabsf(residual) > a.cutoffTH) {
          E += a.maxEnergy; inE++; sat++;
        } else {
          E += hw * residual * residual * (2 - hw);
          inE++; inW++;
          if (a.kind != 1) { o0 = new_idepth; o1 = u; o2 = v; }
          else { o0 = rx0; o1 = rx1; o2 = rx2; }
          o3 = hit.y; o4 = hit.z; o5 = residual; o6 = hw; o7 = refColor;
        }
      }
    }
    const size_t c = a.cap;
    a.warp[i] = o0; a.warp[c + i] = o1; a.warp[2 * c + i] = o2; a.warp[3 * c + i] = o3;
    a.warp[4 * c + i] = o4; a.warp[5 * c + i] = o5; a.warp[6 * c + i] = o6; a.warp[7 * c + i] = o7;
  }
  // block reduction
  for (int o = 16; o > 0; o >>= 1) {
    E += __shfl_xor_sync(0xffffffffu, E, o); sT += __shfl_xor_sync(0xffffffffu, sT, o);
    sRT += __shfl_xor_sync(0xffffffffu, sRT, o); sN += __shfl_xor_sync(0xffffffffu, sN, o);
    inE += __shfl_xor_sync(0xffffffffu, inE, o); inW += __shfl_xor_sync(0xffffffffu, inW, o); sat += __shfl_xor_sync(0xffffffffu, sat, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_E, E); atomicAdd(&s_T, sT); atomicAdd(&s_RT, sRT); atomicAdd(&s_N, sN);
    atomicAdd(&s_c[0], inE); atomicAdd(&s_c[1], inW); atomicAdd(&s_c[2], sat);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&a.acc[0], (double)s_E); atomicAdd(&a.acc[1], (double)s_T); atomicAdd(&a.acc[2], (double)s_RT); atomicAdd(&a.acc[3], (double)s_N);
    atomicAdd(&a.icnt[0], s_c[0]); atomicAdd(&a.icnt[1], s_c[1]); atomicAdd(&a.icnt[2], s_c[2]);
  }
}

}  // namespace

// ---- pre-pyramid image path ---------------------------------------------------------------------
//   PhotometricUndistorter::processFrame<T>   src/util/Undistort.cpp:194-227   (G[raw] * vignetteMapInv, or factor * raw)
//   Undistort::undistort<T>                   src/util/Undistort.cpp:361-458   (bilinear resampling through remapX / remapY)
// One thread per output pixel; the photometric correction is applied to the four source texels on the fly (raw bytes +
// the 1 KB response table instead of a float intermediate image), in the reference's float expression order.
template <class T> __device__ __forceinline__ float photometric(const UndistortArgs &a, const T *raw, int i) {
  const int v = raw[i];
  if (!a.G) return a.factor * v;
  float d = __ldg(a.G + v);
  if (a.vig) d *= __ldg(a.vig + i);
  return d;
}
template <class T> __global__ void __launch_bounds__(256) k_undistort(UndistortArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.w * a.h) return;
  const T *raw = (const T *)a.raw;
  if (!a.remap) { a.out[idx] = photometric(a, raw, idx); return; }
  const float2 r = __ldg(a.remap + idx);
  float xx = r.x, yy = r.y;
  if (xx < 0) { a.out[idx] = 0; return; }
  const int xxi = xx, yyi = yy;
  xx -= xxi; yy -= yyi;
  const float xxyy = xx * yy;
  const int s = xxi + yyi * a.wOrg;
  a.out[idx] = xxyy * photometric(a, raw, s + 1 + a.wOrg) + (yy - xxyy) * photometric(a, raw, s + a.wOrg) + (xx - xxyy) * photometric(a, raw, s + 1) +
               (1 - xx - yy + xxyy) * photometric(a, raw, s);
}
void launch_undistort(sosba *h, const UndistortArgs &a) {
  const int n = a.w * a.h;
  if (a.bits == 8) k_undistort<uint8_t><<<(n + 255) / 256, 256, 0, h->stream>>>(a);
  else k_undistort<uint16_t><<<(n + 255) / 256, 256, 0, h->stream>>>(a);
  h->launches++;
}

// ------------------------------------------------------------------------------------------------
void launch_make_images(sosba *h, int slot, const float *d_color, const float *d_B) {
  // all levels in one launch when the shape allows the tiling (SOSBA_PYR_FUSED=0: one launch per level): 3 to 5 levels (the halo
  // 2^(levels-1) must be a multiple of 4 texels for the 16-byte alignment of the bulk copies), width a multiple of 4, both
  // sides divisible by 2^(levels-1), a 16-byte aligned source
  static const bool fused = [] { const char *e = getenv("SOSBA_PYR_FUSED"); return !(e && e[0] == '0'); }();
  const int L = h->levels, w0 = h->cfg.w, h0 = h->cfg.h, m = (1 << (L - 1)) - 1;
  if (fused && L >= 3 && L <= PYR_MAXL && (w0 & 3) == 0 && (w0 & m) == 0 && (h0 & m) == 0 && ((uintptr_t)d_color & 15) == 0) {
    PyrArgs a;
    a.src = d_color; a.w = w0; a.h = h0; a.levels = L; a.gamma = h->cfg.gamma_weights_pixel_select; a.B = d_B;
    for (int l = 0; l < PYR_MAXL; l++) a.img[l] = l < L ? h->slot_img[slot] + h->lvl_off[l] : nullptr;
    const int H = 1 << (L - 1), SW = PYR_TW + 2 * H, SH = PYR_TH + 2 * H;
    size_t floats = 0;
    for (int l = 0; l < L; l++) floats += (size_t)(SW >> l) * (SH >> l);
    k_pyr_fused<<<dim3((w0 + PYR_TW - 1) / PYR_TW, (h0 + PYR_TH - 1) / PYR_TH), 256, floats * sizeof(float), h->stream>>>(a);
    h->launches++;
    return;
  }
  for (int l = 0; l < h->levels; l++) {
    const int w = h->wl[l], hh = h->hl[l], n = w * hh;
    const float *src = l == 0 ? d_color : h->slot_plane[slot] + h->lvl_off[l - 1];
    const int wm = l == 0 ? w : h->wl[l - 1];
    k_pyr<<<(n + 255) / 256, 256, 0, h->stream>>>(src, wm, h->slot_img[slot] + h->lvl_off[l], h->slot_plane[slot] + h->lvl_off[l], w, hh, l, d_B,
                                                  h->cfg.gamma_weights_pixel_select);
    h->launches++;
  }
}

// SOSBA_LIN_LANES = 8: the round-1 kernels (8 lanes per residual, butterfly-free ordered shuffles); 1, 2, 4: k_linearize_t with
// that many lanes per residual; unset: 4 lanes below 32k residuals (one partial wave: the launch lasts as long as one
// thread's chain), 2 lanes above (measured, fused loop launches, us: 13 672 residuals 13.3 / 9.2 / 8.8 for 1 / 2 / 4 lanes
// in the device timeline; 109 k residuals 38.0 / 32.7 / 43.9; 875 k residuals 235 / 187 / 275)
static int lin_lanes(int R) {
  static const int forced = [] { const char *e = getenv("SOSBA_LIN_LANES"); return e ? atoi(e) : 0; }();
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
  return R < 32768 ? 4 : 2;
}
template <bool APPLY, bool WRITE_J, bool FIX>
static void launch_lin_t(sosba *h, const LinArgs &a, int lanes, bool pdl) {
  const int blocks = (int)(((long long)a.R * lanes + LIN_T - 1) / LIN_T);
  auto go = [&](auto kern) {
    if (pdl) launch_pdl(kern, blocks, LIN_T, 0, h->stream, a);
    else kern<<<blocks, LIN_T, 0, h->stream>>>(a);
  };
  if (lanes == 1) go(k_linearize_t<1, APPLY, WRITE_J, FIX>);
  else if (lanes == 2) go(k_linearize_t<2, APPLY, WRITE_J, FIX>);
  else go(k_linearize_t<4, APPLY, WRITE_J, FIX>);
  h->launches++;
}
void launch_linearize(sosba *h, const LinArgs &a) {
  if (a.R == 0) return;
  const int lanes = lin_lanes(a.R);
  if (lanes == 8) { k_linearize<false, true, false><<<(a.R * 8 + 255) / 256, 256, 0, h->stream>>>(a); h->launches++; }
  else launch_lin_t<false, true, false>(h, a, lanes, false);
}
// linearizeAll(true): linearisation + applyRes(true) + the fixLinearization bookkeeping (FullSystemOptimize.cpp:52-74) in one launch
void launch_linearize_fix(sosba *h, const LinArgs &a) {
  if (a.R == 0) return;
  const int lanes = lin_lanes(a.R);
  if (lanes == 8) {
    k_linearize<false, true, false><<<(a.R * 8 + 255) / 256, 256, 0, h->stream>>>(a);
    k_apply_res<<<(a.R * 8 + 255) / 256, 256, 0, h->stream>>>(a, 1);
    h->launches += 2;
    return;
  }
  launch_lin_t<true, true, true>(h, a, lanes, true);
}
// linearizeAll(false) + setNewFrameEnergyTH + applyRes(true) in one launch (the loop body of FullSystem::optimize)
void launch_linearize_apply(sosba *h, const LinArgs &a, bool write_j, bool th_inline) {
  if (a.R == 0) { if (th_inline) { k_energy_th<<<1, 256, 0, h->stream>>>(a.th, a.gate, 0); h->launches++; } return; }
  const int lanes = lin_lanes(a.R);
  if (lanes != 8) {
    if (write_j) launch_lin_t<true, true, false>(h, a, lanes, true);
    else launch_lin_t<true, false, false>(h, a, lanes, true);
    if (th_inline) launch_energy_th(h, a.th, a.gate);   // rare path (no fused accumulation to carry the selection)
    return;
  }
  const int blocks = (a.R * 8 + 255) / 256;
  if (write_j) {
    if (th_inline) launch_pdl(k_linearize<true, true, true>, blocks, 256, 0, h->stream, a);
    else launch_pdl(k_linearize<true, true, false>, blocks, 256, 0, h->stream, a);
  } else {
    if (th_inline) launch_pdl(k_linearize<true, false, true>, blocks, 256, 0, h->stream, a);
    else launch_pdl(k_linearize<true, false, false>, blocks, 256, 0, h->stream, a);
  }
  h->launches++;
}
void launch_apply_res(sosba *h, const LinArgs &a, int fix) {
  if (a.R == 0) return;
  k_apply_res<<<(a.R * 8 + 255) / 256, 256, 0, h->stream>>>(a, fix);
  h->launches++;
}
void launch_residual_init(sosba *h, const LinArgs &a, const int *p_host) {
  if (a.R == 0) return;
  k_residual_init<<<(a.R + 255) / 256, 256, 0, h->stream>>>(a, p_host);
  h->launches++;
}
void launch_reset_oob(sosba *h, const LinArgs &a, int *zero_words, int n_zero, int *zero_ctl) {
  if (a.R == 0 && !zero_words && !zero_ctl) return;
  k_reset_oob<<<a.R > 0 ? (a.R + 255) / 256 : 1, 256, 0, h->stream>>>(a, zero_words, n_zero, zero_ctl);
  h->launches++;
}
void launch_fix_linearization(sosba *h, const LinArgs &a, const int *d_ids, int n) {
  if (n == 0) return;
  k_fix_linearization<<<(n * 8 + 255) / 256, 256, 0, h->stream>>>(a, d_ids, n);
  h->launches++;
}
void launch_prep_records(sosba *h, const LinArgs &a, int mode, const int *d_list, int n) {
  if (n == 0) return;
  k_prep_records<<<(n * 8 + 255) / 256, 256, 0, h->stream>>>(a, mode, d_list, n);
  h->launches++;
}
void launch_energy_th(sosba *h, const ThArgs &a, const int *gate) {
  // a single rank's list fits the registers (8 per thread); the gathered list of many point shards is staged in shared memory
  const int words = a.nseg > 1 ? 36 * 1024 : 0;
  if (words) cudaFuncSetAttribute(k_energy_th, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4);   // per device: set on every launch
  launch_pdl(k_energy_th, 1, 1024, (size_t)words * 4, h->stream, a, gate, words);
  h->launches++;
}
void launch_track_res(sosba *h, const TrackResArgs &a) {
  if (a.n == 0) return;
  k_track_res<<<(a.n + 255) / 256, 256, 0, h->stream>>>(a);
  h->launches++;
}
