// k_select.cu — pixel selection, compiled with -fmad=false.
//
//   k_sel_hist / k_sel_smooth   PixelSelector::makeHists   src/FullSystem/PixelSelector2.cpp:69-145 (computeHistQuantil :59-67)
//   k_sel_blocks                PixelSelector::select      src/FullSystem/PixelSelector2.cpp:284-422
//   k_sel_scan / k_sel_count / k_sel_compact               raster-order list of the status map (what makeNewTraces walks)
//
// select() is sequential in the reference only through `randomPattern[n2]`: the direction of every block is indexed by the
// number of level-0 selections made so far in scan order.  Here one warp owns one (4*pot)^2 block (sel_block); the running
// count at the block's start comes from an exclusive scan of per-block counts taken in a first, direction-free pass
// (k_sel_count0).  Whether a pot-block selects a pixel depends on its direction only when every
// candidate gradient is exactly orthogonal to it, so the counts of the first pass are almost always already final; the
// second pass re-counts with the true directions and raises a flag if any block disagrees, and the host repeats the
// scan + pass until the counts are a fixed point (then they are the sequential result).
#include <math.h>

#include "kernels.h"

namespace {

__constant__ float c_dir[16][2] = {{0, 1.0000f},        {0.3827f, 0.9239f},  {0.1951f, 0.9808f},  {0.9239f, 0.3827f},
                                   {0.7071f, 0.7071f},  {0.3827f, -0.9239f}, {0.8315f, 0.5556f},  {0.8315f, -0.5556f},
                                   {0.5556f, -0.8315f}, {0.9808f, 0.1951f},  {0.9239f, -0.3827f}, {0.7071f, -0.7071f},
                                   {0.5556f, 0.8315f},  {0.9808f, -0.1951f}, {1.0000f, 0.0000f},  {0.1951f, -0.9808f}};   // PixelSelector2.cpp:297-303

// one CTA per 32x32 block: histogram of int(sqrt(absSquaredGrad)) (capped at 48), then the quantile
__global__ void __launch_bounds__(256) k_sel_hist(SelectArgs a) {
  __shared__ int hist[52];
  const int bx = blockIdx.x % a.w32, by = blockIdx.x / a.w32;
  if (threadIdx.x < 52) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < 1024; p += 256) {
    const int it = (p & 31) + 32 * bx, jt = (p >> 5) + 32 * by;
    if (it > a.w - 2 || jt > a.h - 2 || it < 1 || jt < 1) continue;
    int g = sqrtf(__ldg(&a.img0[it + jt * a.w].w));
    if (g > 48) g = 48;
    atomicAdd(&hist[g + 1], 1);
    atomicAdd(&hist[0], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int th = hist[0] * 0.5f + 0.5f;   // setting_minGradHistCut
    int q = 90;
    for (int i = 0; i < 90; i++) {
      th -= i + 1 < 50 ? hist[i + 1] : 0;
      if (th < 0) { q = i; break; }
    }
    a.ths[bx + by * a.w32] = q + 7.f;   // setting_minGradHistAdd
  }
}

__global__ void k_sel_smooth(SelectArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.w32 * a.h32) return;
  const int x = i % a.w32, y = i / a.w32, w32 = a.w32, h32 = a.h32;
  float sum = 0, num = 0;
  if (x > 0) {
    if (y > 0) { num++; sum += a.ths[x - 1 + (y - 1) * w32]; }
    if (y < h32 - 1) { num++; sum += a.ths[x - 1 + (y + 1) * w32]; }
    num++; sum += a.ths[x - 1 + y * w32];
  }
  if (x < w32 - 1) {
    if (y > 0) { num++; sum += a.ths[x + 1 + (y - 1) * w32]; }
    if (y < h32 - 1) { num++; sum += a.ths[x + 1 + (y + 1) * w32]; }
    num++; sum += a.ths[x + 1 + y * w32];
  }
  if (y > 0) { num++; sum += a.ths[x + (y - 1) * w32]; }
  if (y < h32 - 1) { num++; sum += a.ths[x + (y + 1) * w32]; }
  num++; sum += a.ths[x + y * w32];
  a.thsSmoothed[i] = (sum / num) * (sum / num);
}

// first arg-max across the warp: the largest value wins, ties go to the smaller scan position (= the pixel the
// reference's strict '>' scan would have kept)
__device__ __forceinline__ void warp_first_argmax(float &val, int &ord, int &idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, val, o);
    const int o2 = __shfl_xor_sync(0xffffffffu, ord, o), i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    if (v2 > val || (v2 == val && o2 < ord)) { val = v2; ord = o2; idx = i2; }
  }
}

// One warp owns one (4*pot)^2 block.  The reference scans its pixels one by one and interleaves three tests; what that scan
// computes is, per pot-block, the first arg-max of |g . dir2| over the pixels above the level-0 threshold (a hit kills the
// level-1 pick of its 2*pot-block and the level-2 pick of the 4*pot-block), per 2*pot-block the first arg-max of |g . dir3|
// over the level-1 candidates of the pot-blocks visited while nothing was hit, and the same once more for level 2.  The
// pot-blocks are visited in the reference's order (the running count n2 picks their direction); inside a pot-block the 32
// lanes split the pixels.  Returns (n2, n3, n4) of the block, identical in all lanes.
template <bool WRITE> __device__ __forceinline__ int3 sel_block(const SelectArgs &a, int B, int n2) {
  const int lane = threadIdx.x & 31;
  const int pot = a.pot, w = a.w, h = a.h, w1 = a.w1, w2 = a.w2;
  const int x4 = (B % a.nbx) * 4 * pot, y4 = (B / a.nbx) * 4 * pot;
  const float dw1 = 0.75f, dw2 = dw1 * dw1;   // setting_gradDownweightPerLevel
  const float thFactor = a.thFactor;
  const int n2_start = n2;
  int n3 = 0, n4 = 0;
  const int my3 = min(4 * pot, h - y4), mx3 = min(4 * pot, w - x4);
  bool kill4 = false;            // bestIdx4 == -2
  int bestIdx4 = -1;
  float bestVal4 = 0;
  const int d4 = __ldg(a.randomPattern + n2) & 0xF;
  for (int y3 = 0; y3 < my3; y3 += 2 * pot)
    for (int x3 = 0; x3 < mx3; x3 += 2 * pot) {
      const int x34 = x3 + x4, y34 = y3 + y4;
      const int my2 = min(2 * pot, h - y34), mx2 = min(2 * pot, w - x34);
      bool kill3 = false;        // bestIdx3 == -2
      int bestIdx3 = -1;
      float bestVal3 = 0;
      const int d3 = __ldg(a.randomPattern + n2) & 0xF;
      for (int y2 = 0; y2 < my2; y2 += pot)
        for (int x2 = 0; x2 < mx2; x2 += pot) {
          const int x234 = x2 + x34, y234 = y2 + y34;
          const int my1 = min(pot, h - y234), mx1 = min(pot, w - x234);
          const int d2 = __ldg(a.randomPattern + n2) & 0xF;
          const int npx = mx1 * my1;
          // level 0
          float v0 = 0.f; int o0 = 0x7fffffff, i0 = -1;
          for (int o = lane; o < npx; o += 32) {
            const int xf = o % mx1 + x234, yf = o / mx1 + y234, idx = xf + w * yf;
            if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
            const float pixelTH0 = __ldg(a.thsSmoothed + (xf >> 5) + (yf >> 5) * a.w32);
            const float4 px = __ldg(a.img0 + idx);   // {I, dx, dy, absSquaredGrad}
            if (px.w > pixelTH0 * thFactor) {
              const float dirNorm = fabsf(px.y * c_dir[d2][0] + px.z * c_dir[d2][1]);
              if (dirNorm > v0) { v0 = dirNorm; o0 = o; i0 = idx; }
            }
          }
          warp_first_argmax(v0, o0, i0);
          if (i0 > 0 && v0 > 0.f) {   // bestIdx2 > 0
            if (WRITE && lane == 0) a.map[i0] = 1;
            n2++;
            kill3 = true; kill4 = true;
            continue;
          }
          if (!WRITE) continue;   // the count-only pass needs the level-0 outcome only
          if (kill3) continue;
          // level 1 (reached only while nothing was hit in this 2*pot-block)
          float v1 = 0.f; int o1 = 0x7fffffff, i1 = -1;
          for (int o = lane; o < npx; o += 32) {
            const int xf = o % mx1 + x234, yf = o / mx1 + y234, idx = xf + w * yf;
            if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
            const float pixelTH1 = __ldg(a.thsSmoothed + (xf >> 5) + (yf >> 5) * a.w32) * dw1;
            const float ag1 = __ldg(&a.img1[(int)(xf * 0.5f + 0.25f) + (int)(yf * 0.5f + 0.25f) * w1].w);
            if (ag1 > pixelTH1 * thFactor) {
              const float4 px = __ldg(a.img0 + idx);
              const float dirNorm = fabsf(px.y * c_dir[d3][0] + px.z * c_dir[d3][1]);
              if (dirNorm > v1) { v1 = dirNorm; o1 = o; i1 = idx; }
            }
          }
          warp_first_argmax(v1, o1, i1);
          if (i1 > 0 && v1 > bestVal3) { bestVal3 = v1; bestIdx3 = i1; kill4 = true; }
          if (kill4) continue;
          // level 2 (reached only while nothing at all happened in this 4*pot-block)
          float v2 = 0.f; int o2 = 0x7fffffff, i2 = -1;
          for (int o = lane; o < npx; o += 32) {
            const int xf = o % mx1 + x234, yf = o / mx1 + y234, idx = xf + w * yf;
            if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
            const float pixelTH0 = __ldg(a.thsSmoothed + (xf >> 5) + (yf >> 5) * a.w32);
            const float pixelTH1 = pixelTH0 * dw1, pixelTH2 = pixelTH1 * dw2;
            const float ag2 = __ldg(&a.img2[(int)(xf * 0.25f + 0.125) + (int)(yf * 0.25f + 0.125) * w2].w);
            if (ag2 > pixelTH2 * thFactor) {
              const float4 px = __ldg(a.img0 + idx);
              const float dirNorm = fabsf(px.y * c_dir[d4][0] + px.z * c_dir[d4][1]);
              if (dirNorm > v2) { v2 = dirNorm; o2 = o; i2 = idx; }
            }
          }
          warp_first_argmax(v2, o2, i2);
          if (i2 > 0 && v2 > bestVal4) { bestVal4 = v2; bestIdx4 = i2; }
        }
      if (!kill3 && bestIdx3 > 0) { if (WRITE && lane == 0) a.map[bestIdx3] = 2; n3++; }
    }
  if (!kill4 && bestIdx4 > 0) { if (WRITE && lane == 0) a.map[bestIdx4] = 4; n4++; }
  return make_int3(n2 - n2_start, n3, n4);
}

// First pass: how many pot-blocks of each (4*pot)^2 block hold a pixel above the level-0 threshold.  That is the number of
// level-0 selections the block will make unless every candidate gradient of a pot-block is exactly orthogonal to its
// direction (the write pass checks).  No direction, no arg-max, no order: the lanes sweep the block's pixels with independent
// loads and OR one bit per pot-block.
__global__ void __launch_bounds__(128) k_sel_count0(SelectArgs a) {
  const int B = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (B >= a.nbx * a.nby) return;
  const int pot = a.pot, w = a.w, h = a.h;
  const int x4 = (B % a.nbx) * 4 * pot, y4 = (B / a.nbx) * 4 * pot;
  const int my3 = min(4 * pot, h - y4), mx3 = min(4 * pot, w - x4);
  unsigned mask = 0;
  for (int o = lane; o < mx3 * my3; o += 32) {
    const int x = o % mx3, y = o / mx3, xf = x4 + x, yf = y4 + y;
    if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
    const float pixelTH0 = __ldg(a.thsSmoothed + (xf >> 5) + (yf >> 5) * a.w32);
    const float ag0 = __ldg(&a.img0[xf + w * yf].w);
    if (ag0 > pixelTH0 * a.thFactor) mask |= 1u << ((y / pot) * 4 + x / pot);
  }
  mask = __reduce_or_sync(0xffffffffu, mask);
  if (lane == 0) a.cnt_out[B] = __popc(mask);
}

// One warp per (4*pot)^2 block (WRITE = false would only count: the first pass uses the cheaper k_sel_count0 instead).
template <bool WRITE> __global__ void __launch_bounds__(128) k_sel_blocks(SelectArgs a) {
  const int B = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (B >= a.nbx * a.nby) return;
  const int3 r = sel_block<WRITE>(a, B, WRITE ? a.base[B] : 0);
  if ((threadIdx.x & 31) != 0) return;
  const int cnt = r.x, n3 = r.y, n4 = r.z;
  if (WRITE) {
    if (cnt != a.cnt_in[B]) atomicExch(&a.totals[3], 1);   // the counts the bases were scanned from were not final
    a.cnt_out[B] = cnt;
    if (cnt) atomicAdd(&a.totals[0], cnt);
    if (n3) atomicAdd(&a.totals[1], n3);
    if (n4) atomicAdd(&a.totals[2], n4);
  } else {
    a.cnt_out[B] = cnt;
  }
}

// Fallback for inputs on which the counts do not settle (gradients exactly orthogonal to a direction over large areas, e.g.
// a synthetic image that varies along one axis only): the reference's order, one warp, blocks one after the other.
__global__ void __launch_bounds__(32) k_sel_serial(SelectArgs a) {
  int n2 = 0, n3 = 0, n4 = 0;
  for (int B = 0; B < a.nbx * a.nby; B++) {
    const int3 r = sel_block<true>(a, B, n2);
    n2 += r.x; n3 += r.y; n4 += r.z;
  }
  if (threadIdx.x == 0) { a.totals[0] = n2; a.totals[1] = n3; a.totals[2] = n4; a.totals[3] = 0; }
}

// exclusive scan of n ints by one CTA (n <= a few 10k): out[i] = sum in[0..i), *total = sum
__global__ void __launch_bounds__(1024) k_sel_scan(const int *in, int *out, int n, int *total) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int ws = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += y; }
      s_warp[lane] = ws;
    }
    __syncthreads();
    const int prefix = s_carry + (wid ? s_warp[wid - 1] : 0) + x - v;
    if (i < n) out[i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = s_carry;
}

// raster-order compaction of the status map: chunks of 1024 pixels
__global__ void __launch_bounds__(256) k_sel_count(const uint8_t *map, int n, int *chunk_cnt) {
  __shared__ int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  const int base = blockIdx.x * 1024 + threadIdx.x * 4;
  int c = 0;
  if (base + 3 < n) { const uchar4 m = *(const uchar4 *)(map + base); c = (m.x != 0) + (m.y != 0) + (m.z != 0) + (m.w != 0); }
  else for (int k = 0; k < 4; k++) if (base + k < n) c += map[base + k] != 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) chunk_cnt[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_sel_compact(const uint8_t *map, int n, const int *chunk_off, int cap, int2 *list) {
  __shared__ int s_w[8];
  const int base = blockIdx.x * 1024 + threadIdx.x * 4;
  uint8_t m[4] = {0, 0, 0, 0};
  if (base + 3 < n) { const uchar4 q = *(const uchar4 *)(map + base); m[0] = q.x; m[1] = q.y; m[2] = q.z; m[3] = q.w; }
  else for (int k = 0; k < 4; k++) if (base + k < n) m[k] = map[base + k];
  const int c = (m[0] != 0) + (m[1] != 0) + (m[2] != 0) + (m[3] != 0);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) s_w[wid] = x;
  __syncthreads();
  int off = chunk_off[blockIdx.x] + x - c;
  for (int k = 0; k < wid; k++) off += s_w[k];
  for (int k = 0; k < 4; k++)
    if (m[k]) { if (off < cap) list[off] = make_int2(base + k, m[k]); off++; }
}

}  // namespace

void launch_select_hists(sosba *h, const SelectArgs &a) {
  k_sel_hist<<<a.w32 * a.h32, 256, 0, h->stream>>>(a);
  k_sel_smooth<<<(a.w32 * a.h32 + 127) / 128, 128, 0, h->stream>>>(a);
  h->launches += 2;
}
void launch_select_blocks(sosba *h, const SelectArgs &a, bool write) {
  const int nb = a.nbx * a.nby;
  if (write) k_sel_blocks<true><<<(nb + 3) / 4, 128, 0, h->stream>>>(a);
  else k_sel_count0<<<(nb + 3) / 4, 128, 0, h->stream>>>(a);
  h->launches++;
}
void launch_select_serial(sosba *h, const SelectArgs &a) {
  k_sel_serial<<<1, 32, 0, h->stream>>>(a);
  h->launches++;
}
void launch_select_scan(sosba *h, const int *in, int *out, int n, int *total) {
  k_sel_scan<<<1, 1024, 0, h->stream>>>(in, out, n, total);
  h->launches++;
}
void launch_select_compact(sosba *h, const uint8_t *map, int n, int *chunk_cnt, int *chunk_off, int *total, int cap, int2 *list) {
  const int chunks = (n + 1023) / 1024;
  k_sel_count<<<chunks, 256, 0, h->stream>>>(map, n, chunk_cnt);
  k_sel_scan<<<1, 1024, 0, h->stream>>>(chunk_cnt, chunk_off, chunks, total);
  k_sel_compact<<<chunks, 256, 0, h->stream>>>(map, n, chunk_off, cap, list);
  h->launches += 3;
}
