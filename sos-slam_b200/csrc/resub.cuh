// resub.cuh — EnergyFunctional::resubstituteFPt (EnergyFunctional.cpp:526-551) (+ the point part of backupState /
// doStepFromBackup when do_step) as a CTA-level device function: 8 lanes per point, lane q owns residual
// res_begin[p] + q; the subtraction chain runs in residual order like the reference.  Shared by k_resubstitute
// (k_accum.cu) and k_step (k_solve.cu: the loop-body launch that also carries the frame step in a spare CTA).
#pragma once
#include "kernels.h"

__device__ __forceinline__ void resubstitute_body(const ResubArgs &a, const int block, const int nblocks) {
  __shared__ double s_sum[3];
  if (a.zero_lin && block == 0 && threadIdx.x < 7) {   // energy | pad | counts[0..4] of the linearisation that follows
    if (threadIdx.x < 2) a.zero_lin[threadIdx.x] = 0.0;
    else ((int *)(a.zero_lin + 2))[threadIdx.x - 2] = 0;
  }
  if (a.zero_newE)
    for (int i = block * blockDim.x + threadIdx.x; i < a.zero_newE_n; i += nblocks * blockDim.x) a.zero_newE[i] = 0.f;
  if (threadIdx.x < 3) s_sum[threadIdx.x] = 0.0;
  __syncthreads();
  const int gid = block * blockDim.x + threadIdx.x;
  const int p = gid >> 3, sub = gid & 7;
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  float st2 = 0.f, absid = 0.f, one = 0.f;
  if (p < a.P) {
    const int rb = a.res_begin[p], re = a.res_begin[p + 1];
    const int host = a.p_host[p];
    const float *xc = a.xAd + (size_t)a.nf * a.nf * 8;
    float b = a.bdSumF[p];
    float dotc = 0.f;
    for (int i = 0; i < 4; i++) dotc += xc[i] * (a.HcdA[4 * p + i] + a.HcdL[4 * p + i]);
    b -= dotc;
    int ngood = 0;
    for (int base = rb; base < re; base += 8) {
      const int r = base + sub;
      const bool use = r < re && a.r_is_active[r] && !a.r_dropped[r];
      float s = 0.f;
      if (use) {
        const float4 *xa = (const float4 *)(a.xAd + 8 * (size_t)(host * a.nf + a.r_target[r]));
        const float4 *v = (const float4 *)(a.rec + (size_t)r * SOSBA_CREC + CR_JPJDF);
        const float4 x0 = xa[0], x1 = xa[1], v0 = v[0], v1 = v[1];
        s = x0.x * v0.x; s += x0.y * v0.y; s += x0.z * v0.z; s += x0.w * v0.w;
        s += x1.x * v1.x; s += x1.y * v1.y; s += x1.z * v1.z; s += x1.w * v1.w;
      }
      ngood += __popc(__ballot_sync(gmask, use) & gmask);
      const int cnt = min(8, re - base);
      for (int q = 0; q < cnt; q++) b -= __shfl_sync(gmask, s, q, 8);   // inactive lanes contribute 0
    }
    if (sub == 0) {
      const float step = ngood > 0 ? -b * a.HdiF[p] : 0.f;
      a.step[p] = step;
      if (a.do_step) {
        const float backup = a.idepth[p];       // backupState: idepth_backup = idepth
        a.idepth_backup[p] = backup;
        const float nid = backup + step;        // stepfacD = 1
        a.idepth[p] = nid; a.idepth_zero[p] = nid; a.deltaF[p] = 0.f;
        st2 = step * step; absid = fabsf(backup); one = 1.f;
      }
    }
  }
  if (a.do_step) {
    for (int o = 16; o > 0; o >>= 1) {
      st2 += __shfl_xor_sync(0xffffffffu, st2, o); absid += __shfl_xor_sync(0xffffffffu, absid, o); one += __shfl_xor_sync(0xffffffffu, one, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_sum[0], (double)st2); atomicAdd(&s_sum[1], (double)absid); atomicAdd(&s_sum[2], (double)one); }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(&a.stats[1 + threadIdx.x], s_sum[threadIdx.x]);
  }
}

