// energy_th.cuh — FullSystem::setNewFrameEnergyTH (FullSystemOptimize.cpp:84-124) as a CTA-wide device function:
// exact k-th smallest (std::nth_element) of the newest-frame residual energies by a 4-pass radix select on the bit
// patterns of the (non-negative) floats.  The values are read from global memory once (kept in registers when the list
// fits 8 per thread, else in the caller's shared-memory scratch when it fits there), histogram updates are aggregated per
// warp (most energies share their top byte).
// Included by k_exact.cu (stand-alone launch, last CTA of the linearisation) and k_accum.cu (spare CTA of the
// accumulation); the float formula uses explicit _rn intrinsics so both translation units round identically.
#pragma once
#include "kernels.h"

// One 256-bin histogram per warp (shared-memory atomics never cross warps; inside a warp the hardware serialises lanes that hit the
// same bin), summed into hist[] by the caller.  Up to ENERGY_TH_WARPS warps take part; further warps share the last copy.
constexpr int ENERGY_TH_WARPS = 8;
__device__ inline void energy_th_body(const ThArgs &a, unsigned *cache = nullptr, int cache_n = 0) {
  __shared__ unsigned hist[256];
  __shared__ unsigned whist[ENERGY_TH_WARPS][256];
  __shared__ unsigned s_prefix, s_k;
  __shared__ int s_off[17];   // prefix sums of the segment lengths
  if (threadIdx.x == 0) {
    int o = 0;
    for (int sgm = 0; sgm < a.nseg; sgm++) { s_off[sgm] = o; o += a.seg_counts[sgm]; }
    s_off[a.nseg] = o;
  }
  __syncthreads();
  const int n = s_off[a.nseg];
  const unsigned *v = (const unsigned *)a.newE;
  const int nt = blockDim.x;
  if (n == 0) {
    if (threadIdx.x == 0) { a.frameEnergyTH[a.nf - 1] = 12 * 12 * 8; a.thOut[0] = 12 * 12 * 8; }
    return;
  }
  auto elem = [&](int i) -> unsigned {   // i-th energy of the concatenated list
    int sgm = 0;
    while (sgm + 1 < a.nseg && i >= s_off[sgm + 1]) sgm++;
    return v[(size_t)sgm * a.seg_stride + (i - s_off[sgm])];
  };
  constexpr int KEEP = 8;
  unsigned mine[KEEP];
  const bool cached = n <= KEEP * nt;
  if (cached) {
#pragma unroll
    for (int q = 0; q < KEEP; q++) { const int i = threadIdx.x + q * nt; mine[q] = i < n ? elem(i) : 0u; }
  }
  const bool in_smem = !cached && cache && n <= cache_n;
  if (in_smem) {   // one pass over global memory, four loads in flight per thread
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * nt) {
      unsigned x[4];
#pragma unroll
      for (int q = 0; q < 4; q++) { const int i = i0 + q * nt; x[q] = i < n ? elem(i) : 0u; }
#pragma unroll
      for (int q = 0; q < 4; q++) { const int i = i0 + q * nt; if (i < n) cache[i] = x[q]; }
    }
  }
  if (threadIdx.x == 0) { s_prefix = 0; s_k = (unsigned)(int)(a.thN * n); }
  unsigned *myhist = whist[min((int)(threadIdx.x >> 5), ENERGY_TH_WARPS - 1)];
  for (int pass = 3; pass >= 0; pass--) {
    for (int i = threadIdx.x; i < ENERGY_TH_WARPS * 256; i += nt) (&whist[0][0])[i] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix, shift = 8 * pass;
    const unsigned himask = pass == 3 ? 0u : (0xFFFFFFFFu << (shift + 8));
    if (cached) {
#pragma unroll
      for (int q = 0; q < KEEP; q++) {
        const int i = threadIdx.x + q * nt;
        const unsigned x = mine[q];
        if (i < n && (x & himask) == prefix) atomicAdd(&myhist[(x >> shift) & 255u], 1u);
      }
    } else {
      for (int i0 = 0; i0 < n; i0 += nt) {
        const int i = i0 + threadIdx.x;
        const unsigned x = i < n ? (in_smem ? cache[i] : elem(i)) : 0u;
        if (i < n && (x & himask) == prefix) atomicAdd(&myhist[(x >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += nt) {
      unsigned s = 0;
#pragma unroll
      for (int w = 0; w < ENERGY_TH_WARPS; w++) s += whist[w][i];
      hist[i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // warp 0: find the bin holding rank s_k (8 bins per lane, inclusive scan over lanes)
      const unsigned lane = threadIdx.x;
      unsigned c[8], tot = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) { c[q] = hist[8 * lane + q]; tot += c[q]; }
      unsigned incl = tot;
      for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += u; }
      const unsigned excl = incl - tot, k = s_k;
      const bool own = k >= excl && k < incl;
      __syncwarp();
      if (own) {
        unsigned kk = k - excl, b = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { if (kk >= c[q] && b == (unsigned)q) { kk -= c[q]; b = q + 1; } }
        s_k = kk; s_prefix = prefix | ((8 * lane + b) << shift);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float nthElement = sqrtf(__uint_as_float(s_prefix));
    float th = __fmul_rn(nthElement, a.thFacMedian);
    th = __fadd_rn(__fmul_rn(26.0f, a.thConstWeight), __fmul_rn(th, __fsub_rn(1.0f, a.thConstWeight)));
    th = __fmul_rn(th, th);
    th = __fmul_rn(th, __fmul_rn(a.overallWeight, a.overallWeight));
    a.frameEnergyTH[a.nf - 1] = th;
    a.thOut[0] = th;
  }
}
