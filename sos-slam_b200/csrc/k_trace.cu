// k_trace.cu — immature points, compiled with -fmad=false: every float expression follows the
// reference's order, the outcome (ImmaturePointStatus, the new inverse-depth interval) is bit-exact against the oracle.
//
//   immature_init   ImmaturePoint::ImmaturePoint       src/FullSystem/ImmaturePoint.cpp:28-60
//   trace_on        ImmaturePoint::traceOn             src/FullSystem/ImmaturePoint.cpp:70-415
//                   FullSystem::traceNewCoarse         src/FullSystem/FullSystem.cpp:311-361
//
// Eight lanes per immature point, one per pattern tap: the epipolar search is a sequential walk of up to 99 steps (the
// positions accumulate ptx += dx in float), each step 8 bilinear taps that the lanes fetch in parallel and sum in pattern
// order with shuffles; then at most 3 Gauss-Newton steps on the line.  Points of one host are contiguous, so neighbouring
// lane groups walk neighbouring epipolar segments of the same image (L1/L2 locality).
#include <math.h>

#include "kernels.h"

namespace {

__constant__ int t_pattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};  // settings.cpp:307-309

// globalFuncs.h:122-136 on the float4 image (channel 0)
__device__ __forceinline__ float interp31(const float4 *__restrict__ img, float x, float y, int width) {
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy;
  float dxdy = dx * dy;
  const float4 *bp = img + ix + iy * width;
  return dxdy * __ldg(&bp[1 + width].x) + (dy - dxdy) * __ldg(&bp[width].x) + (dx - dxdy) * __ldg(&bp[1].x) + (1 - dx - dy + dxdy) * __ldg(&bp[0].x);
}
// globalFuncs.h:68-82
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

__global__ void __launch_bounds__(128) k_immature_init(TraceArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.n) return;
  const int u = a.iu[p], v = a.iv[p];
  float G0 = 0, G1 = 0, G2 = 0, G3 = 0;
  bool bad = false;
  for (int idx = 0; idx < 8 && !bad; idx++) {
    // getInterpolatedElement33BiLin (globalFuncs.h:161-182) at the integer pixel (u + dx, v + dy)
    const float x = (float)(u + t_pattern[idx][0]), y = (float)(v + t_pattern[idx][1]);
    const int ix = (int)x, iy = (int)y;
    const float4 *bp = a.img + ix + iy * a.w;
    const float tl = __ldg(&bp[0].x), tr = __ldg(&bp[1].x), bl = __ldg(&bp[a.w].x), br = __ldg(&bp[a.w + 1].x);
    const float dx = x - ix, dy = y - iy;
    const float topInt = dx * tr + (1 - dx) * tl;
    const float botInt = dx * br + (1 - dx) * bl;
    const float leftInt = dy * bl + (1 - dy) * tl;
    const float rightInt = dy * br + (1 - dy) * tr;
    const float c0 = dx * rightInt + (1 - dx) * leftInt, c1 = rightInt - leftInt, c2 = botInt - topInt;
    a.color_out[8 * (size_t)p + idx] = c0;
    if (!isfinite(c0)) { bad = true; break; }
    G0 += c1 * c1; G1 += c1 * c2; G2 += c2 * c1; G3 += c2 * c2;
    a.weights_out[8 * (size_t)p + idx] = sqrtf(a.outlierTHSum / (a.outlierTHSum + (c1 * c1 + c2 * c2)));
  }
  a.gradH_out[4 * (size_t)p + 0] = G0; a.gradH_out[4 * (size_t)p + 1] = G1; a.gradH_out[4 * (size_t)p + 2] = G2; a.gradH_out[4 * (size_t)p + 3] = G3;
  if (bad) { a.energyTH_out[p] = nanf(""); return; }
  float th = 8 * 144.0f;   // patternNum * setting_outlierTH (settings.cpp:82)
  th *= a.overallWeight * a.overallWeight;
  a.energyTH_out[p] = th;
}

// Sum of the 8 per-tap terms of one point in pattern order, starting from `acc`: ((acc + t0) + t1) + ... -- the order
// of the reference's scalar loop.  `gmask` is the mask of the point's 8 lanes; every lane returns the same value.
__device__ __forceinline__ float ordered_sum8(unsigned gmask, float acc, float term) {
#pragma unroll
  for (int i = 0; i < 8; i++) acc += __shfl_sync(gmask, term, i, 8);
  return acc;
}
// same, but only the first `limit` taps contribute (a pattern aborted at tap `limit` keeps its partial sums)
__device__ __forceinline__ float ordered_sum8_upto(unsigned gmask, float acc, float term, int limit) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float t = __shfl_sync(gmask, term, i, 8);
    if (i < limit) acc += t;
  }
  return acc;
}

#define TRACE_THREADS 128
#define TRACE_PPB (TRACE_THREADS / 8)   // points per block

// 8 lanes per point, lane = pattern tap.  All 8 lanes evaluate the (identical) scalar set-up, so every branch is uniform
// within a point's lane group; the per-step energies are summed across the lanes in pattern order.
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_on(TraceArgs a) {
  __shared__ int s_cnt[6];
  __shared__ float s_err[TRACE_PPB][100];
  if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int grp = threadIdx.x >> 3, tap = threadIdx.x & 7;
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  const int k = blockIdx.x * TRACE_PPB + grp;
  if (k < a.n) {
    int status = a.status[k];
    if (status != SOSBA_IPS_OOB) {
      const int wG = a.w, hG = a.h;
      const float *KRKi = a.KRKi + 9 * a.host[k], *Kt = a.Kt + 3 * a.host[k], *aff = a.aff + 2 * a.host[k];
      const float pu = a.u[k], pv = a.v[k];
      float idepth_min = a.idepth_min[k], idepth_max = a.idepth_max[k];
      float uv0 = -1.f, uv1 = -1.f, pixint = 0.f;
      const float maxPixSearch = (wG + hG) * 0.027f;
      float pr[3], ptpMin[3];
#pragma unroll
      for (int i = 0; i < 3; i++) pr[i] = (KRKi[3 * i] * pu + KRKi[3 * i + 1] * pv) + KRKi[3 * i + 2] * 1.0f;
#pragma unroll
      for (int i = 0; i < 3; i++) ptpMin[i] = pr[i] + Kt[i] * idepth_min;
      const float uMin = ptpMin[0] / ptpMin[2], vMin = ptpMin[1] / ptpMin[2];
      bool done = false;
      if (!(uMin > 4 && vMin > 4 && uMin < wG - 5 && vMin < hG - 5)) { status = SOSBA_IPS_OOB; done = true; }
      float dist = 0.f, uMax = 0.f, vMax = 0.f;
      const bool finite_max = isfinite(idepth_max);
      if (!done) {
        if (finite_max) {
          float ptpMax[3];
#pragma unroll
          for (int i = 0; i < 3; i++) ptpMax[i] = pr[i] + Kt[i] * idepth_max;
          uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
          if (!(uMax > 4 && vMax > 4 && uMax < wG - 5 && vMax < hG - 5)) { status = SOSBA_IPS_OOB; done = true; }
          else {
            dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
            dist = sqrtf(dist);
            if (dist < 1.5f) {   // setting_trace_slackInterval
              uv0 = (uMax + uMin) * 0.5f; uv1 = (vMax + vMin) * 0.5f; pixint = dist;
              status = SOSBA_IPS_SKIPPED; done = true;
            }
          }
        } else {
          dist = maxPixSearch;
          float ptpMax[3];
#pragma unroll
          for (int i = 0; i < 3; i++) ptpMax[i] = pr[i] + Kt[i] * 0.01f;
          uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
          const float ddx = uMax - uMin, ddy = vMax - vMin;
          const float d = 1.0f / sqrtf(ddx * ddx + ddy * ddy);
          uMax = uMin + dist * ddx * d;
          vMax = vMin + dist * ddy * d;
          if (!(uMax > 4 && vMax > 4 && uMax < wG - 5 && vMax < hG - 5)) { status = SOSBA_IPS_OOB; done = true; }
        }
      }
      if (!done && !(idepth_min < 0 || (ptpMin[2] > 0.75f && ptpMin[2] < 1.5f))) { status = SOSBA_IPS_OOB; done = true; }
      float dx = 0.f, dy = 0.f, errorInPixel = 0.f;
      if (!done) {
        dx = 1.0f * (uMax - uMin); dy = 1.0f * (vMax - vMin);   // setting_trace_stepsize
        const float *G = a.gradH + 4 * (size_t)k;
        const float ea = (dx * G[0] + dy * G[2]) * dx + (dx * G[1] + dy * G[3]) * dy;
        const float eb = (dy * G[0] + (-dx) * G[2]) * dy + (dy * G[1] + (-dx) * G[3]) * (-dx);
        errorInPixel = 0.2f + 0.2f * (ea + eb) / ea;
        if (errorInPixel * 2.f > dist && finite_max) {   // setting_trace_minImprovementFactor
          uv0 = (uMax + uMin) * 0.5f; uv1 = (vMax + vMin) * 0.5f; pixint = dist;
          status = SOSBA_IPS_BADCONDITION; done = true;
        }
      }
      if (!done) {
        if (errorInPixel > 10) errorInPixel = 10;
        dx /= dist; dy /= dist;
        if (dist > maxPixSearch) { uMax = uMin + maxPixSearch * dx; vMax = vMin + maxPixSearch * dy; dist = maxPixSearch; }
        int numSteps = 1.9999f + dist / 1.0f;
        const float randShift = uMin * 1000 - floorf(uMin * 1000);
        float ptx = uMin - randShift * dx, pty = vMin - randShift * dy;
        // this lane's rotated pattern offset
        const float rot0 = KRKi[0] * t_pattern[tap][0] + KRKi[1] * t_pattern[tap][1];
        const float rot1 = KRKi[3] * t_pattern[tap][0] + KRKi[4] * t_pattern[tap][1];
        if (!isfinite(dx) || !isfinite(dy)) { status = SOSBA_IPS_OOB; done = true; }
        if (!done) {
          const float color = a.color[8 * (size_t)k + tap], weight = a.weights[8 * (size_t)k + tap];
          const float col = (float)(aff[0] * color + aff[1]);
          float *errors = s_err[grp];
          float bestU = 0, bestV = 0, bestEnergy = 1e10f;
          int bestIdx = -1;
          if (numSteps >= 100) numSteps = 99;
          for (int i = 0; i < numSteps; i++) {
            const float hit = interp31(a.img, (float)(ptx + rot0), (float)(pty + rot1), wG);
            float term = 1e5f;
            if (isfinite(hit)) {
              const float residual = hit - col;
              const float hw = fabsf(residual) < a.huberTH ? 1 : a.huberTH / fabsf(residual);
              term = hw * residual * residual * (2 - hw);
            }
            const float energy = ordered_sum8(gmask, 0.f, term);
            if (tap == (i & 7)) errors[i] = energy;
            if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
            ptx += dx; pty += dy;
          }
          __syncwarp(gmask);
          // min over the steps outside the test radius: each lane scans a stride-8 subset, then the group takes the min
          float secondBest = 1e10f;
          for (int i = tap; i < numSteps; i += 8)
            if ((i < bestIdx - 2 || i > bestIdx + 2) && errors[i] < secondBest) secondBest = errors[i];   // setting_minTraceTestRadius
#pragma unroll
          for (int o = 4; o >= 1; o >>= 1) secondBest = fminf(secondBest, __shfl_xor_sync(gmask, secondBest, o, 8));
          const float newQuality = secondBest / bestEnergy;
          float quality = a.quality[k];
          if (newQuality < quality || numSteps > 10) quality = newQuality;
          __syncwarp(gmask);
          if (tap == 0) a.quality[k] = quality;

          float uBak = bestU, vBak = bestV, stepBack = 0;
          bestEnergy = 1e5f;
          for (int it = 0; it < 3; it++) {   // setting_trace_GNIterations
            const float3 hit = interp33(a.img, (float)(bestU + rot0), (float)(bestV + rot1), wG);
            float tH = 0.f, tb = 0.f, tE = 1e5f;
            if (isfinite(hit.x)) {
              const float residual = hit.x - (aff[0] * color + aff[1]);
              const float dResdDist = dx * hit.y + dy * hit.z;
              const float hw = fabsf(residual) < a.huberTH ? 1 : a.huberTH / fabsf(residual);
              tH = hw * dResdDist * dResdDist;
              tb = hw * residual * dResdDist;
              tE = weight * weight * hw * residual * residual * (2 - hw);
            }
            // a non-finite tap adds nothing to H and b in the reference; adding +0.f leaves the sums bit-identical
            // unless a sum is -0.f, which cannot happen (H starts at 1; b = -0.f + 0.f = +0.f only if every term so far was -0.f,
            // and then b * anything compares and steps identically)
            const float H = ordered_sum8(gmask, 1.f, tH);
            const float bb = ordered_sum8(gmask, 0.f, tb);
            const float energy = ordered_sum8(gmask, 0.f, tE);
            if (energy > bestEnergy) {
              stepBack *= 0.5f;
              bestU = uBak + stepBack * dx;
              bestV = vBak + stepBack * dy;
            } else {
              float step = -1.0f * bb / H;
              if (step < -0.5f) step = -0.5f;
              else if (step > 0.5f) step = 0.5f;
              if (!isfinite(step)) step = 0;
              uBak = bestU; vBak = bestV; stepBack = step;
              bestU += step * dx; bestV += step * dy;
              bestEnergy = energy;
            }
            if (fabsf(stepBack) < 0.1f) break;   // setting_trace_GNThreshold
          }
          if (!(bestEnergy < a.energyTH[k] * 1.2f)) {   // setting_trace_extraSlackOnTH
            status = status == SOSBA_IPS_OUTLIER ? SOSBA_IPS_OOB : SOSBA_IPS_OUTLIER;
          } else {
            if (dx * dx > dy * dy) {
              idepth_min = (pr[2] * (bestU - errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU - errorInPixel * dx));
              idepth_max = (pr[2] * (bestU + errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU + errorInPixel * dx));
            } else {
              idepth_min = (pr[2] * (bestV - errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV - errorInPixel * dy));
              idepth_max = (pr[2] * (bestV + errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV + errorInPixel * dy));
            }
            if (idepth_min > idepth_max) { const float t = idepth_min; idepth_min = idepth_max; idepth_max = t; }
            if (!isfinite(idepth_min) || !isfinite(idepth_max) || (idepth_max < 0)) {
              status = SOSBA_IPS_OUTLIER;
            } else {
              pixint = 2 * errorInPixel; uv0 = bestU; uv1 = bestV;
              status = SOSBA_IPS_GOOD;
            }
            // the interval is rewritten before the NaN test (ImmaturePoint.cpp:385-405)
            if (tap == 0) { a.idepth_min[k] = idepth_min; a.idepth_max[k] = idepth_max; }
          }
        }
      }
      if (tap == 0) {
        a.status[k] = (uint8_t)status;
        a.uv[2 * (size_t)k] = uv0; a.uv[2 * (size_t)k + 1] = uv1; a.pixint[k] = pixint;
      }
    }
    if (tap == 0) atomicAdd(&s_cnt[status], 1);
  }
  __syncthreads();
  if (threadIdx.x < 6 && s_cnt[threadIdx.x]) atomicAdd(&a.counts[threadIdx.x], s_cnt[threadIdx.x]);
}

// ---- activation ---------------------------------------------------------------------------------
//   linearize_residual   ImmaturePoint::linearizeResidual      src/FullSystem/ImmaturePoint.cpp:475-545
//   k_optimize_immature  FullSystem::optimizeImmaturePoint     src/FullSystem/FullSystemOptPoint.cpp:47-192
//                        (activatePointsMT_Reductor, FullSystem.cpp:363-373: 8 lanes per immature point, one per pattern tap)
struct ActPoint {
  float u, v, energyTH;
  float color, w2;   // this lane's tap; w2 = weights[idx] * weights[idx], the only form the weights appear in
  int host;
};

// Eight lanes per point, lane = pattern tap.  Returns the residual's energy (a float value, or the stored state_energy);
// Hdd / bd accumulate tap by tap in pattern order, and the partial sums of a residual that leaves the image half-way through
// the pattern stay in (ImmaturePoint.cpp:497-534): the first failing tap `limit` cuts the ordered sums.
__device__ __forceinline__ float linearize_residual(const ActivateArgs &a, const ActPoint &p, unsigned gmask, int tap, float slack, int target, uint8_t state,
                                                    float state_energy, uint8_t &newState, float &newEnergy, float &Hdd, float &bd, float idepth) {
  if (state == SOSBA_RES_OOB) { newState = SOSBA_RES_OOB; return state_energy; }
  const size_t pair = (size_t)p.host * a.nf + target;
  const float *R = a.RTll + 9 * pair, *t = a.tTll + 3 * pair;
  const float aff0 = __ldg(a.aff + 2 * pair), aff1 = __ldg(a.aff + 2 * pair + 1);
  float Rr[9], tt[3];
#pragma unroll
  for (int i = 0; i < 9; i++) Rr[i] = __ldg(R + i);
#pragma unroll
  for (int i = 0; i < 3; i++) tt[i] = __ldg(t + i);
  const float fxli = 1.0f / a.fxl, fyli = 1.0f / a.fyl;
  const float wM3G = a.w - 3, hM3G = a.h - 3;
  const float4 *img = a.img[target];
  const int dx = t_pattern[tap][0], dy = t_pattern[tap][1];
  const float K0 = (p.u + dx - a.cxl) * fxli, K1 = (p.v + dy - a.cyl) * fyli;
  float ptp[3];
#pragma unroll
  for (int i = 0; i < 3; i++) ptp[i] = ((Rr[3 * i] * K0 + Rr[3 * i + 1] * K1) + Rr[3 * i + 2] * 1.0f) + tt[i] * idepth;
  const float drescale = 1.0f / ptp[2];
  bool ok = drescale > 0;
  const float u = ptp[0] * drescale, v = ptp[1] * drescale;
  const float Ku = u * a.fxl + a.cxl, Kv = v * a.fyl + a.cyl;
  ok = ok && (Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G);
  float tE = 0.f, tH = 0.f, tb = 0.f;
  if (ok) {
    const float3 hit = interp33(img, Ku, Kv, a.w);
    if (!isfinite(hit.x)) ok = false;
    else {
      const float residual = hit.x - (aff0 * p.color + aff1);
      float hw = fabsf(residual) < a.huberTH ? 1 : a.huberTH / fabsf(residual);
      tE = p.w2 * hw * residual * residual * (2 - hw);
      const float dxInterp = hit.y * a.fxl, dyInterp = hit.z * a.fyl;
      const float d_idepth = (dxInterp * drescale * (tt[0] - tt[2] * u) + dyInterp * drescale * (tt[1] - tt[2] * v)) * 1.0f;   // SCALE_IDEPTH
      hw *= p.w2;
      tH = (hw * d_idepth) * d_idepth;
      tb = (hw * residual) * d_idepth;
    }
  }
  const unsigned shift = (threadIdx.x & 31) & ~7;
  const unsigned fail = (__ballot_sync(gmask, !ok) >> shift) & 0xFFu;
  const int limit = fail ? __ffs(fail) - 1 : 8;
  Hdd = ordered_sum8_upto(gmask, Hdd, tH, limit);
  bd = ordered_sum8_upto(gmask, bd, tb, limit);
  if (fail) { newState = SOSBA_RES_OOB; return state_energy; }
  float energyLeft = ordered_sum8(gmask, 0.f, tE);
  if (energyLeft > p.energyTH * slack) { energyLeft = p.energyTH * slack; newState = SOSBA_RES_OUTLIER; }
  else newState = SOSBA_RES_IN;
  newEnergy = energyLeft;
  return energyLeft;
}

__global__ void __launch_bounds__(TRACE_THREADS) k_optimize_immature(ActivateArgs a) {
  const int grp = threadIdx.x >> 3, tap = threadIdx.x & 7;
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  const int k = blockIdx.x * TRACE_PPB + grp;
  if (k >= a.n) return;
  ActPoint p;
  p.u = a.u[k]; p.v = a.v[k]; p.energyTH = a.energyTH[k]; p.host = a.host[k];
  p.color = a.color[8 * (size_t)k + tap];
  { const float w = a.weights[8 * (size_t)k + tap]; p.w2 = w * w; }
  // ImmaturePointTemporaryResidual per target frame (slot of the host unused).  The 8 lanes of a point hold identical values
  // and all of them write, so every lane only ever reads back what it wrote itself: shared memory instead of four
  // dynamically indexed local arrays, no synchronisation needed.
  __shared__ uint8_t s_st[TRACE_PPB][2][SOSBA_ACT_MAXF];
  __shared__ float s_en[TRACE_PPB][2][SOSBA_ACT_MAXF];
  uint8_t *st = s_st[grp][0], *nst = s_st[grp][1];
  float *en = s_en[grp][0], *nen = s_en[grp][1];
  const int nf = a.nf;
  for (int f = 0; f < nf; f++) { st[f] = SOSBA_RES_IN; nst[f] = SOSBA_RES_OUTLIER; en[f] = nen[f] = 0.f; }
  float lastEnergy = 0, lastHdd = 0, lastbd = 0;
  float currentIdepth = (a.idepth_max[k] + a.idepth_min[k]) * 0.5f;
  for (int f = 0; f < nf; f++) {
    if (f == p.host) continue;
    const float e = linearize_residual(a, p, gmask, tap, 1000.f, f, st[f], en[f], nst[f], nen[f], lastHdd, lastbd, currentIdepth);
    lastEnergy = (float)((double)lastEnergy + (double)e);   // float += double (FullSystemOptPoint.cpp:70)
    __syncwarp(gmask);   // every lane has read the old state before any lane commits the new one
    st[f] = nst[f]; en[f] = nen[f];
  }
  int result = SOSBA_ACT_ACTIVATED;
  if (!isfinite(lastEnergy) || lastHdd < 100.f) result = SOSBA_ACT_SKIP;   // setting_minIdepthH_act
  if (result == SOSBA_ACT_ACTIVATED) {
    float lambda = 0.1f;
    for (int iteration = 0; iteration < 3; iteration++) {   // setting_GNItsOnPointActivation
      float H = lastHdd;
      H *= 1 + lambda;
      const float step = (float)((1.0 / (double)H) * (double)lastbd);
      const float newIdepth = currentIdepth - step;
      float newHdd = 0, newbd = 0, newEnergy = 0;
      for (int f = 0; f < nf; f++) {
        if (f == p.host) continue;
        const float e = linearize_residual(a, p, gmask, tap, 1.f, f, st[f], en[f], nst[f], nen[f], newHdd, newbd, newIdepth);
        newEnergy = (float)((double)newEnergy + (double)e);
      }
      if (!isfinite(lastEnergy) || newHdd < 100.f) { result = SOSBA_ACT_SKIP; break; }
      if (newEnergy < lastEnergy) {
        currentIdepth = newIdepth; lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
        __syncwarp(gmask);
        for (int f = 0; f < nf; f++) { st[f] = nst[f]; en[f] = nen[f]; }
        __syncwarp(gmask);
        lambda = (float)((double)lambda * 0.5);
      } else {
        lambda *= 5;
      }
      if ((double)fabsf(step) < 0.0001 * (double)currentIdepth) break;
    }
  }
  __syncwarp(gmask);
  if (tap != 0) return;
  int good = 0;
  for (int f = 0; f < nf; f++) {
    a.res_state[(size_t)k * nf + f] = f == p.host ? 255 : st[f];
    good += (f != p.host && st[f] == SOSBA_RES_IN);
  }
  if (result == SOSBA_ACT_ACTIVATED && (!isfinite(currentIdepth) || good < a.min_obs || !isfinite(p.energyTH))) result = SOSBA_ACT_DELETE;
  a.result[k] = (signed char)result;
  a.idepth[k] = currentIdepth;
}


// ---- initializer ----------------------------------------------------------------------------------
//   k_init_res   CoarseInitializer::calcResAndGS   src/FullSystem/CoarseInitializer.cpp:450-673
// One thread per Pnt: the 8 pattern taps in order (per-point outputs -- energy_new, isGood_new, maxstep, JbBuffer_new,
// lastHessian_new -- are bit-exact), its 45 + 45 upper-triangle products for acc9 / acc9SC in registers, then a shuffle tree
// and one fp64 red.global per entry per block (the reference's 4-lane float accumulators sum in a different order: H, b
// agree to float rounding).  alphaOpt is known before the launch: the reference's EAlpha never receives a term (:606-617
// add to E), so alphaEnergy = alphaW * |t|^2 * npts does not depend on the points.
template <int NV> __device__ __forceinline__ void block_sum_to(float (&v)[NV], double *__restrict__ dst) {
  __shared__ float s_part[4][NV];
#pragma unroll
  for (int q = 0; q < NV; q++)
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < NV; q++) s_part[warp][q] = v[q];
  __syncthreads();
  for (int q = threadIdx.x; q < NV; q += blockDim.x) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += (double)s_part[w][q];
    if (s != 0.0) atomicAdd(dst + q, s);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(128) k_init_res(InitArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[45];
#pragma unroll
  for (int q = 0; q < 45; q++) acc[q] = 0.f;
  float Jb[10];
#pragma unroll
  for (int q = 0; q < 10; q++) Jb[q] = 0.f;
  float E = 0.f;
  bool good_new = false;
  if (i < a.n) {
    float maxstep_pt = 1e10f;
    const float e0 = a.energy[2 * i], e1 = a.energy[2 * i + 1];
    bool isGood = a.isGood[i] != 0;
    float energy = 0.f;
    if (isGood) {
      const float pu = a.u[i], pv = a.v[i], idn = a.idepth_new[i];
      for (int idx = 0; idx < 8; idx++) {
        const int dx = t_pattern[idx][0], dy = t_pattern[idx][1];
        const float x = pu + dx, y = pv + dy;
        float pt[3];
#pragma unroll
        for (int k = 0; k < 3; k++) pt[k] = ((a.RKi[3 * k] * x + a.RKi[3 * k + 1] * y) + a.RKi[3 * k + 2] * 1.0f) + a.t[k] * idn;
        const float u = pt[0] / pt[2], v = pt[1] / pt[2];
        const float Ku = a.fx * u + a.cx, Kv = a.fy * v + a.cy;
        const float new_idepth = idn / pt[2];
        if (!(Ku > 1 && Kv > 1 && Ku < a.w - 2 && Kv < a.h - 2 && new_idepth > 0)) { isGood = false; break; }
        const float3 hit = interp33(a.imgNew, Ku, Kv, a.w);
        const float rlR = interp31(a.imgRef, x, y, a.w);
        if (!isfinite(rlR) || !isfinite(hit.x)) { isGood = false; break; }
        const float residual = hit.x - a.aff0 * rlR - a.aff1;
        float hw = fabsf(residual) < a.huberTH ? 1 : a.huberTH / fabsf(residual);
        energy += hw * residual * residual * (2 - hw);
        const float dxdd = (a.t[0] - a.t[2] * u) / pt[2];
        const float dydd = (a.t[1] - a.t[2] * v) / pt[2];
        if (hw < 1) hw = sqrtf(hw);
        const float dxInterp = hw * hit.y * a.fx, dyInterp = hw * hit.z * a.fy;
        float J[9];
        J[0] = new_idepth * dxInterp;
        J[1] = new_idepth * dyInterp;
        J[2] = -new_idepth * (u * dxInterp + v * dyInterp);
        J[3] = -u * v * dxInterp - (1 + v * v) * dyInterp;
        J[4] = (1 + u * u) * dxInterp + u * v * dyInterp;
        J[5] = -v * dxInterp + u * dyInterp;
        J[6] = -hw * a.aff0 * rlR;
        J[7] = -hw * 1;
        J[8] = hw * residual;
        const float dd = dxInterp * dxdd + dyInterp * dydd;
        const float mx = dxdd * a.fx, my = dydd * a.fy;
        const float maxstep = 1.0f / sqrtf(mx * mx + my * my);
        if (maxstep < maxstep_pt) maxstep_pt = maxstep;
#pragma unroll
        for (int k = 0; k < 8; k++) Jb[k] += J[k] * dd;
        Jb[8] += J[8] * dd;
        Jb[9] += dd * dd;
        int q = 0;
#pragma unroll
        for (int r = 0; r < 9; r++)
#pragma unroll
          for (int c = r; c < 9; c++) acc[q++] += J[r] * J[c];
      }
    }
    // (a point that was not good keeps a zero JbBuffer row only if it was good on entry: the reference zeroes the row after
    // the isGood test, :485-499)
    good_new = isGood && !(energy > a.outlierTH[i] * 20);
    a.maxstep[i] = maxstep_pt;
    a.isGood_new[i] = good_new ? 1 : 0;
    if (good_new) {
      E = energy;
      a.energy_new[2 * i] = energy;
      a.energy_new[2 * i + 1] = (a.idepth_new[i] - 1) * (a.idepth_new[i] - 1);
    } else {
      E = e0;
      a.energy_new[2 * i] = e0; a.energy_new[2 * i + 1] = e1;
#pragma unroll
      for (int q = 0; q < 45; q++) acc[q] = 0.f;
    }
    if (a.isGood[i] != 0) {   // JbBuffer_new[i] before the Schur step
      if (good_new) {
        a.lastHessian_new[i] = Jb[9];
        Jb[8] += a.alphaOpt * (a.idepth_new[i] - 1);
        Jb[9] += a.alphaOpt;
        if (a.alphaOpt == 0) {
          Jb[8] += a.couplingWeight * (a.idepth_new[i] - a.iR[i]);
          Jb[9] += a.couplingWeight;
        }
        Jb[9] = 1 / (1 + Jb[9]);
      }
#pragma unroll
      for (int q = 0; q < 10; q++) a.Jb[10 * (size_t)i + q] = Jb[q];
    }
  }
  block_sum_to<45>(acc, a.acc);
  // acc9SC.updateSingleWeighted (MatrixAccumulators.h:1543-1660): diagonal (Jr * Jr) * w, then Jr *= w and Jc * Jr
  {
    int q = 0;
    const float w = Jb[9];
#pragma unroll
    for (int r = 0; r < 9; r++) {
      acc[q++] = good_new ? Jb[r] * Jb[r] * w : 0.f;
      const float Jr = Jb[r] * w;
#pragma unroll
      for (int c = r + 1; c < 9; c++) acc[q++] = good_new ? Jb[c] * Jr : 0.f;
    }
  }
  block_sum_to<45>(acc, a.acc + 45);
  float e1v[1] = {E};
  block_sum_to<1>(e1v, a.acc + 90);
}

}  // namespace

void launch_init_res(sosba *h, const InitArgs &a) {
  if (a.n == 0) return;
  k_init_res<<<(a.n + 127) / 128, 128, 0, h->stream>>>(a);
  h->launches++;
}
void launch_optimize_immature(sosba *h, const ActivateArgs &a) {
  if (a.n == 0) return;
  k_optimize_immature<<<(a.n + TRACE_PPB - 1) / TRACE_PPB, TRACE_THREADS, 0, h->stream>>>(a);
  h->launches++;
}
void launch_immature_init(sosba *h, const TraceArgs &a) {
  if (a.n == 0) return;
  k_immature_init<<<(a.n + 127) / 128, 128, 0, h->stream>>>(a);
  h->launches++;
}
void launch_trace_on(sosba *h, const TraceArgs &a) {
  if (a.n == 0) return;
  k_trace_on<<<(a.n + TRACE_PPB - 1) / TRACE_PPB, TRACE_THREADS, 0, h->stream>>>(a);
  h->launches++;
}
