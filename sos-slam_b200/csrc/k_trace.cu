// k_trace.cu — immature points, compiled with -fmad=false: every float expression follows the
// reference's order, the outcome (ImmaturePointStatus, the new inverse-depth interval) is bit-exact against the oracle.
//
//   immature_init   ImmaturePoint::ImmaturePoint       src/FullSystem/ImmaturePoint.cpp:28-60
//   trace_on        ImmaturePoint::traceOn             src/FullSystem/ImmaturePoint.cpp:70-415
//                   FullSystem::traceNewCoarse         src/FullSystem/FullSystem.cpp:311-361
//
// One thread per immature point: the epipolar search is a sequential walk of up to 99 steps with 8 bilinear taps each,
// followed by at most 3 Gauss-Newton steps on the line; points of one host are contiguous, so neighbouring threads walk
// neighbouring epipolar segments of the same image (L1/L2 locality).
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

__global__ void __launch_bounds__(128) k_trace_on(TraceArgs a) {
  __shared__ int s_cnt[6];
  if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
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
        float rot[8][2];
#pragma unroll
        for (int idx = 0; idx < 8; idx++) {
          rot[idx][0] = KRKi[0] * t_pattern[idx][0] + KRKi[1] * t_pattern[idx][1];
          rot[idx][1] = KRKi[3] * t_pattern[idx][0] + KRKi[4] * t_pattern[idx][1];
        }
        if (!isfinite(dx) || !isfinite(dy)) { status = SOSBA_IPS_OOB; done = true; }
        if (!done) {
          const float *color = a.color + 8 * (size_t)k, *weights = a.weights + 8 * (size_t)k;
          float col[8];
#pragma unroll
          for (int idx = 0; idx < 8; idx++) col[idx] = (float)(aff[0] * color[idx] + aff[1]);
          float errors[100];
          float bestU = 0, bestV = 0, bestEnergy = 1e10f;
          int bestIdx = -1;
          if (numSteps >= 100) numSteps = 99;
          for (int i = 0; i < numSteps; i++) {
            float energy = 0;
#pragma unroll
            for (int idx = 0; idx < 8; idx++) {
              const float hit = interp31(a.img, (float)(ptx + rot[idx][0]), (float)(pty + rot[idx][1]), wG);
              if (!isfinite(hit)) { energy += 1e5f; continue; }
              const float residual = hit - col[idx];
              const float hw = fabsf(residual) < a.huberTH ? 1 : a.huberTH / fabsf(residual);
              energy += hw * residual * residual * (2 - hw);
            }
            errors[i] = energy;
            if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
            ptx += dx; pty += dy;
          }
          float secondBest = 1e10f;
          for (int i = 0; i < numSteps; i++)
            if ((i < bestIdx - 2 || i > bestIdx + 2) && errors[i] < secondBest) secondBest = errors[i];   // setting_minTraceTestRadius
          const float newQuality = secondBest / bestEnergy;
          float quality = a.quality[k];
          if (newQuality < quality || numSteps > 10) quality = newQuality;
          a.quality[k] = quality;

          float uBak = bestU, vBak = bestV, stepBack = 0;
          bestEnergy = 1e5f;
          for (int it = 0; it < 3; it++) {   // setting_trace_GNIterations
            float H = 1, bb = 0, energy = 0;
#pragma unroll
            for (int idx = 0; idx < 8; idx++) {
              const float3 hit = interp33(a.img, (float)(bestU + rot[idx][0]), (float)(bestV + rot[idx][1]), wG);
              if (!isfinite(hit.x)) { energy += 1e5f; continue; }
              const float residual = hit.x - (aff[0] * color[idx] + aff[1]);
              const float dResdDist = dx * hit.y + dy * hit.z;
              const float hw = fabsf(residual) < a.huberTH ? 1 : a.huberTH / fabsf(residual);
              H += hw * dResdDist * dResdDist;
              bb += hw * residual * dResdDist;
              energy += weights[idx] * weights[idx] * hw * residual * residual * (2 - hw);
            }
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
            a.idepth_min[k] = idepth_min; a.idepth_max[k] = idepth_max;   // the interval is rewritten before the NaN test (ImmaturePoint.cpp:385-405)
          }
        }
      }
      a.status[k] = (uint8_t)status;
      a.uv[2 * (size_t)k] = uv0; a.uv[2 * (size_t)k + 1] = uv1; a.pixint[k] = pixint;
    }
    atomicAdd(&s_cnt[status], 1);
  }
  __syncthreads();
  if (threadIdx.x < 6 && s_cnt[threadIdx.x]) atomicAdd(&a.counts[threadIdx.x], s_cnt[threadIdx.x]);
}

}  // namespace

void launch_immature_init(sosba *h, const TraceArgs &a) {
  if (a.n == 0) return;
  k_immature_init<<<(a.n + 127) / 128, 128, 0, h->stream>>>(a);
  h->launches++;
}
void launch_trace_on(sosba *h, const TraceArgs &a) {
  if (a.n == 0) return;
  k_trace_on<<<(a.n + 127) / 128, 128, 0, h->stream>>>(a);
  h->launches++;
}
