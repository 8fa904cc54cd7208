// k_lm.cu — the control loops of the direct alignment, resident on the device (SURVEY.md §8 row a16 and the optimizeScale part
// of a17).  Built with -fmad=false: the per-point float expressions decide integer bookkeeping (numTermsInE, numSaturated, the
// cutoff doubling, accept / reject) and follow k_track_res / the reference operation by operation.
//
//   CoarseTracker::makeCoarseDepthL0     src/FullSystem/CoarseTracker.cpp:56-230    k_cd_* + sosba_tracker_make_coarse_depth
//   CoarseTracker::scaleCoarseDepthL0    :244-251                                   k_cd_scale
//   CoarseTracker::trackNewestCoarse     :366-552 (calcResPose :612-764, calcGSSSEPose :554-610 inside)   k_track_lm
//   ScaleOptimizer::optimizeScale        src/FullSystem/ScaleOptimizer.cpp:120-230 (calcResScale :273-437, calcGSSSEScale
//                                        :232-271 inside)                           k_scale_lm
//
// Re-design: the reference runs one hypothesis after the other and, per Levenberg-Marquardt iteration, one serial pass over the
// reference points for the residual, a second one for the normal equations, then an 8x8 solve on the same core.  Here ONE
// THREAD BLOCK OWNS ONE HYPOTHESIS for its whole life (all levels, all iterations): every iteration is a single pass in which
// each thread warps its points, taps the new frame and accumulates energy, counters AND the 45 normal-equation sums of the
// candidate (the reference recomputes them only after an accept; here they are simply kept when the step is accepted), a
// block reduction, and the damped LDL^T solve / SE3 update / accept test on one thread — no kernel boundary, no host round
// trip, no warped buffers in memory.  The ~110 hypotheses of FullSystem::trackNewCoarse (FullSystem.cpp:175-231) or the 7
// start scales of FullSystem::optimizeScale (:1133-1146) are one launch: a block each.
#include <limits.h>
#include <math.h>

#include <vector>

#include "kernels.h"
#include "lm_math.cuh"

int sosba_tracker_reserve(sosba *h, int lvl, int n);   // sosba_api.cu

namespace {

constexpr int LM_THREADS = 512;
constexpr int LM_WARPS = LM_THREADS / 32;

__device__ __forceinline__ float3 lm_mul33(const float *M, float x, float y, float z) {
  return make_float3((M[0] * x + M[1] * y) + M[2] * z, (M[3] * x + M[4] * y) + M[5] * z, (M[6] * x + M[7] * y) + M[8] * z);
}
// globalFuncs.h:68-82 on the float4 image {I, dx, dy, absSquaredGrad}: weights and summation order verbatim
__device__ __forceinline__ float3 lm_interp33(const float4 *__restrict__ img, float x, float y, int width) {
  const int ix = (int)x, iy = (int)y;
  const float dx = x - ix, dy = y - iy;
  const float dxdy = dx * dy;
  const float4 *bp = img + ix + iy * width;
  const float4 t11 = __ldg(bp + 1 + width), t01 = __ldg(bp + width), t10 = __ldg(bp + 1), t00 = __ldg(bp);
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  float3 o;
  o.x = w11 * t11.x + w01 * t01.x + w10 * t10.x + w00 * t00.x;
  o.y = w11 * t11.y + w01 * t01.y + w10 * t10.y + w00 * t00.y;
  o.z = w11 * t11.z + w01 * t01.z + w10 * t10.z + w00 * t00.z;
  return o;
}

// ---- makeCoarseDepthL0 -------------------------------------------------------------------------------------------
// (1) splat: idepth / weight sums of level 0 with float atomics.  A pixel hit by one or two points gets the reference's sum
//     exactly (0 + a, a + b are order-free); pixels hit by three or more are redone in point order by k_cd_fix.
__global__ void k_cd_splat(int n, const float *__restrict__ cpt, const float *__restrict__ HdiF, int w0, float *idepth, float *wsum, int *cnt, int *first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int u = cpt[3 * i] + 0.5f, v = cpt[3 * i + 1] + 0.5f;
  const float new_idepth = cpt[3 * i + 2];
  const float weight = sqrtf(1e-3 / (HdiF[i] + 1e-12));
  const int p = u + w0 * v;
  atomicAdd(idepth + p, new_idepth * weight);
  atomicAdd(wsum + p, weight);
  atomicAdd(cnt + p, 1);
  atomicMin(first + p, i);
}
__global__ void k_cd_fix(int n, const float *__restrict__ cpt, const float *__restrict__ HdiF, int w0, float *idepth, float *wsum, const int *cnt, const int *first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int u = cpt[3 * i] + 0.5f, v = cpt[3 * i + 1] + 0.5f;
  const int p = u + w0 * v;
  if (cnt[p] < 3 || first[p] != i) return;
  float si = 0.f, sw = 0.f;
  for (int j = i; j < n; j++) {
    const int uj = cpt[3 * j] + 0.5f, vj = cpt[3 * j + 1] + 0.5f;
    if (uj + w0 * vj != p) continue;
    const float weight = sqrtf(1e-3 / (HdiF[j] + 1e-12));
    si += cpt[3 * j + 2] * weight;
    sw += weight;
  }
  idepth[p] = si; wsum[p] = sw;
}
// (2) 2x2 sum-pool of both maps (:81-101)
__global__ void k_cd_pool(int wl, int hl, int wlm1, const float *__restrict__ idm, const float *__restrict__ wsm, float *idl, float *wsl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= wl * hl) return;
  const int x = i % wl, y = i / wl, b = 2 * x + 2 * y * wlm1;
  idl[i] = idm[b] + idm[b + 1] + idm[b + wlm1] + idm[b + wlm1 + 1];
  wsl[i] = wsm[b] + wsm[b + 1] + wsm[b + wlm1] + wsm[b + wlm1 + 1];
}
// (3) dilation by one: empty pixels take the mean of their non-empty neighbours (diagonal neighbours on levels 0 and 1 :104-146,
//     axis neighbours below :149-190).  Reads only pixels that are non-empty in the copy `bak`, writes only empty ones: no race.
//     Neighbours outside the map (the reference reads index -1 / w*h there) count as empty.
__global__ void k_cd_dilate(int wl, int N, int diag, const float *__restrict__ bak, float *idepth, float *wsum) {
  const int i = wl + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N - wl) return;
  if (!(bak[i] <= 0)) return;
  const int o0 = diag ? 1 + wl : 1, o1 = diag ? -1 - wl : -1, o2 = diag ? wl - 1 : wl, o3 = diag ? -wl + 1 : -wl;
  const int off[4] = {o0, o1, o2, o3};
  float sum = 0, num = 0, numn = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int j = i + off[k];
    if (j < 0 || j >= N) continue;
    const float b = bak[j];
    if (b > 0) { sum += idepth[j]; num += b; numn++; }
  }
  if (numn > 0) { idepth[i] = sum / numn; wsum[i] = num / numn; }
}
// (4) normalise and mark the pixels that enter the point list (:193-229)
__global__ void k_cd_mark(int wl, int hl, const float4 *__restrict__ img, float *idepth, float *wsum, uint8_t *mark) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= wl * hl) return;
  const int x = i % wl, y = i / wl;
  uint8_t m = 0;
  if (x >= 2 && x < wl - 2 && y >= 2 && y < hl - 2) {
    if (wsum[i] > 0) {
      const float id = idepth[i] / wsum[i];
      const float color = img[i].x;
      if (!isfinite(color) || !(id > 0)) idepth[i] = -1;   // (the reference `continue`s before resetting the weight)
      else { idepth[i] = id; m = 1; wsum[i] = 1; }
    } else { idepth[i] = -1; wsum[i] = 1; }
  }
  mark[i] = m;
}
// (5) the lists, in raster order: pc = u | v | idepth | color with stride n
__global__ void k_cd_gather(int n, int wl, const int2 *__restrict__ list, const float *__restrict__ idepth, const float4 *__restrict__ img, float *pc) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = list[k].x;
  pc[k] = (float)(i % wl); pc[n + k] = (float)(i / wl); pc[2 * (size_t)n + k] = idepth[i]; pc[3 * (size_t)n + k] = img[i].x;
}
__global__ void k_cd_scale(int n, float *idepth, float scale) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) idepth[k] /= scale;
}

// ---- CoarseDistanceMap::makeDistanceMap + growDistBFS (CoarseTracker.cpp:789-916) -----------------------------------------------
// The reference floods the level-1 map breadth first from the projected points with two explicit queues, 39 steps, alternating
// 8- and 4-neighbourhoods.  What that computes is a level-synchronous rule with no queue: at step k a pixel still above k takes k
// if one of its neighbours holds k - 1 and is not a border pixel (border pixels never spread, :836-838).  A step moves at most one
// pixel per axis, so a tile plus a 39-pixel halo is self-contained: ONE launch, every block floods its tile in shared memory
// (bytes, 255 = unreached), in place — a pixel set to k in this step is not k - 1, so it cannot feed another pixel of the same step.
constexpr int DM_TILE = 64, DM_HALO = 39, DM_SPAN = DM_TILE + 2 * DM_HALO;
__global__ void k_dm_seed(int n, const int *__restrict__ host, const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ idepth,
                          const float *__restrict__ KRKi, const float *__restrict__ Kt, int w1, int h1, uint8_t *seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *M = KRKi + 9 * host[i], *t = Kt + 3 * host[i];
  const float x = u[i], y = v[i], id = idepth[i];
  const float p0 = ((M[0] * x + M[1] * y) + M[2] * 1) + t[0] * id;
  const float p1 = ((M[3] * x + M[4] * y) + M[5] * 1) + t[1] * id;
  const float p2 = ((M[6] * x + M[7] * y) + M[8] * 1) + t[2] * id;
  const float fu = p0 / p2 + 0.5f, fv = p1 / p2 + 0.5f;
  if (!(fabsf(fu) < 1e9f) || !(fabsf(fv) < 1e9f)) return;   // (the reference's int conversion of such a value is undefined)
  const int uu = (int)fu, vv = (int)fv;
  if (!(uu > 0 && vv > 0 && uu < w1 && vv < h1)) return;
  seed[uu + w1 * vv] = 1;
}
__global__ void __launch_bounds__(1024) k_dm_flood(int w1, int h1, const uint8_t *__restrict__ seed, float *__restrict__ dist) {
  __shared__ uint8_t s[DM_SPAN * DM_SPAN];
  const int x0 = blockIdx.x * DM_TILE - DM_HALO, y0 = blockIdx.y * DM_TILE - DM_HALO;
  for (int e = threadIdx.x; e < DM_SPAN * DM_SPAN; e += blockDim.x) {
    const int x = x0 + e % DM_SPAN, y = y0 + e / DM_SPAN;
    s[e] = (x >= 0 && y >= 0 && x < w1 && y < h1 && seed[x + w1 * y]) ? 0 : 255;
  }
  __syncthreads();
  for (int k = 1; k < 40; k++) {
    for (int e = threadIdx.x; e < DM_SPAN * DM_SPAN; e += blockDim.x) {
      if (s[e] <= k) continue;
      const int lx = e % DM_SPAN, ly = e / DM_SPAN, x = x0 + lx, y = y0 + ly;
      if (x < 0 || y < 0 || x >= w1 || y >= h1) continue;
      bool hit = false;
      const int nn = (k & 1) ? 8 : 4;
      const int dx[8] = {1, -1, 0, 0, 1, -1, -1, 1}, dy[8] = {0, 0, 1, -1, 1, 1, -1, -1};
#pragma unroll
      for (int q = 0; q < 8; q++) {
        if (q >= nn) break;
        const int qx = lx - dx[q], qy = ly - dy[q];           // the neighbour that would have spread to this pixel
        if (qx < 0 || qy < 0 || qx >= DM_SPAN || qy >= DM_SPAN) continue;
        const int gx = x0 + qx, gy = y0 + qy;
        if (gx <= 0 || gy <= 0 || gx >= w1 - 1 || gy >= h1 - 1) continue;   // outside, or a border pixel: does not spread
        if (s[qy * DM_SPAN + qx] == k - 1) hit = true;
      }
      if (hit) s[e] = (uint8_t)k;
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < DM_TILE * DM_TILE; e += blockDim.x) {
    const int lx = DM_HALO + e % DM_TILE, ly = DM_HALO + e / DM_TILE, x = x0 + lx, y = y0 + ly;
    if (x < w1 && y < h1) { const uint8_t v = s[ly * DM_SPAN + lx]; dist[x + w1 * y] = v == 255 ? 1000.f : (float)v; }
  }
}

// ---- the per-iteration pass --------------------------------------------------------------------------------------
struct LmLevel {
  int w, h, n;
  const float *pc;       // u | v | idepth | color, stride n
  const float4 *img;     // level image of the new frame (pose) / of camera 1 (scale)
  float K[4];            // fx fy cx cy of the camera the points are projected into
  float Ki[9];           // inverse intrinsics of the reference camera (makeK)
};
struct LmArgs {
  int levels, coarsest;
  LmLevel lv[SOSBA_MAX_LEVELS];
  float huberTH, coarseCutoffTH, modeA, modeB;
  float ref_exp, new_exp;
  double ref_aff[2];
  float R10[9], t10[3];  // scale: tfmF0ToF1
  void *hyps;            // device array of sosba_track_hypothesis / sosba_scale_hypothesis
  float *terms;          // per hypothesis 2 x terms_stride floats: the energy terms of the accepted state and of the candidate
  int terms_stride;
};

// what one pass needs besides the level: the candidate
struct LmCand {
  float RKi[9], t[3];    // pose: R Ki, t      scale: R10 Ki (un-scaled), t10
  float aff0, aff1;      // pose: affLL
  float scale;           // scale
  float cutoffTH, maxEnergy;
};

constexpr int NSUM_POSE = 45 + 4, NSUM_SCALE = 3 + 4;   // normal-equation sums + E, flowT, flowRT, flowNum

// One pass over the reference points of a level for candidate `c`: calcRes + calcGSSSE fused.  All threads of the block.
// Results (block-wide sums) land in s_sum[0..NS) (doubles; [0..NA) normal equations, then E, sT, sRT, sN) and s_cnt[0..3)
// = numTermsInE, numTermsInWarped, numSaturated.
// terms[i] = the energy term of point i (0 for a skipped point): kept so that a near-tie between two candidates can be decided
// on the reference's own float sum, taken in point order (lm_decide).
template <int KIND>
__device__ __forceinline__ void lm_pass(const LmLevel &L, int lvl, const LmCand &c, float huberTH, float b0, float (*s_part)[52], int (*s_parti)[3],
                                        double *s_sum, int *s_cnt, float *__restrict__ terms) {
  constexpr int NA = KIND == 0 ? 45 : 3, NS = NA + 4;
  float acc[NA];
#pragma unroll
  for (int q = 0; q < NA; q++) acc[q] = 0.f;
  float E = 0.f, sT = 0.f, sRT = 0.f, sN = 0.f;
  int inE = 0, inW = 0, sat = 0;
  const float fx = L.K[0], fy = L.K[1], cx = L.K[2], cy = L.K[3];
  const int n = L.n;
  for (int i = threadIdx.x; i < n; i += LM_THREADS) {
    const float x = L.pc[i], y = L.pc[n + i], id = L.pc[2 * (size_t)n + i], refColor = L.pc[3 * (size_t)n + i];
    float3 pt;
    float rx0 = 0.f, rx1 = 0.f, rx2 = 0.f;
    if (KIND == 0) pt = lm_mul33(c.RKi, x, y, 1.f);
    else {   // scale * RKi (ScaleOptimizer.cpp:296-297): the matrix entries are scaled first
      float sRKi[9];
#pragma unroll
      for (int k = 0; k < 9; k++) sRKi[k] = c.scale * c.RKi[k];
      pt = lm_mul33(sRKi, x, y, 1.f);
      const float3 rx = lm_mul33(c.RKi, x, y, 1.f);
      rx0 = rx.x / id; rx1 = rx.y / id; rx2 = rx.z / id;
    }
    pt.x = pt.x + c.t[0] * id; pt.y = pt.y + c.t[1] * id; pt.z = pt.z + c.t[2] * id;
    const float u = pt.x / pt.z, v = pt.y / pt.z;
    const float Ku = fx * u + cx, Kv = fy * v + cy;
    const float new_idepth = id / pt.z;
    if (lvl == 0 && i % 32 == 0) {   // flow indicators (CoarseTracker.cpp:666-696, ScaleOptimizer.cpp:318-350)
      float sKi[9];
#pragma unroll
      for (int k = 0; k < 9; k++) sKi[k] = KIND == 0 ? L.Ki[k] : c.scale * L.Ki[k];
      const float3 kp = lm_mul33(sKi, x, y, 1.f);
      const float3 ptT = make_float3(kp.x + c.t[0] * id, kp.y + c.t[1] * id, kp.z + c.t[2] * id);
      const float KuT = fx * (ptT.x / ptT.z) + cx, KvT = fy * (ptT.y / ptT.z) + cy;
      const float3 ptT2 = make_float3(kp.x - c.t[0] * id, kp.y - c.t[1] * id, kp.z - c.t[2] * id);
      const float KuT2 = fx * (ptT2.x / ptT2.z) + cx, KvT2 = fy * (ptT2.y / ptT2.z) + cy;
      float3 rp;
      if (KIND == 0) rp = lm_mul33(c.RKi, x, y, 1.f);
      else {
        float sRKi[9];
#pragma unroll
        for (int k = 0; k < 9; k++) sRKi[k] = c.scale * c.RKi[k];
        rp = lm_mul33(sRKi, x, y, 1.f);
      }
      const float3 pt3 = make_float3(rp.x - c.t[0] * id, rp.y - c.t[1] * id, rp.z - c.t[2] * id);
      const float Ku3 = fx * (pt3.x / pt3.z) + cx, Kv3 = fy * (pt3.y / pt3.z) + cy;
      sT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sN += 2;
    }
    terms[i] = 0.f;
    if (!(Ku > 2 && Kv > 2 && Ku < L.w - 3 && Kv < L.h - 3 && new_idepth > 0)) continue;
    const float3 hit = lm_interp33(L.img, Ku, Kv, L.w);
    if (!isfinite(hit.x)) continue;
    const float residual = KIND == 0 ? hit.x - (float)(c.aff0 * refColor + c.aff1) : hit.x - refColor;
    const float hw = fabsf(residual) < huberTH ? 1 : huberTH / fabsf(residual);
    if (fabsf(residual) > c.cutoffTH) { E += c.maxEnergy; terms[i] = c.maxEnergy; inE++; sat++; continue; }
    const float term = hw * residual * residual * (2 - hw);
    E += term; terms[i] = term;
    inE++; inW++;
    if (KIND == 0) {   // calcGSSSEPose row (CoarseTracker.cpp:570-593) weighted by hw (Accumulator9::updateSSE_eighted)
      const float dx = hit.y * fx, dy = hit.z * fy;
      float J[9];
      J[0] = new_idepth * dx;
      J[1] = new_idepth * dy;
      J[2] = 0 - new_idepth * (u * dx + v * dy);
      J[3] = 0 - ((u * v) * dx + dy * (1 + v * v));
      J[4] = (u * v) * dy + dx * (1 + u * u);
      J[5] = u * dy - v * dx;
      J[6] = c.aff0 * (b0 - refColor);
      J[7] = -1;
      J[8] = residual;
      int q = 0;
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const float Jw = J[r] * hw;
#pragma unroll
        for (int cc = r; cc < 9; cc++) acc[q++] += Jw * J[cc];
      }
    } else {           // calcGSSSEScale row (ScaleOptimizer.cpp:245-262)
      const float dxfx = hit.y * fx, dyfy = hit.z * fy;
      const float deno_sqrt = c.scale * rx2 + c.t[2];
      const float deno = 1.0f / (deno_sqrt * deno_sqrt);
      const float xno = rx0 * c.t[2] - rx2 * c.t[0], yno = rx1 * c.t[2] - rx2 * c.t[1];
      const float J0 = dxfx * (deno * xno) + dyfy * (deno * yno), J1 = residual;
      const float J0w = J0 * hw, J1w = J1 * hw;
      acc[0] += J0w * J0; acc[1] += J0w * J1; acc[2] += J1w * J1;
    }
  }
  // block reduction: shuffle tree per warp, then one thread per sum adds the warps in double
#pragma unroll
  for (int q = 0; q < NA; q++)
    for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
  for (int o = 16; o > 0; o >>= 1) {
    E += __shfl_xor_sync(0xffffffffu, E, o); sT += __shfl_xor_sync(0xffffffffu, sT, o);
    sRT += __shfl_xor_sync(0xffffffffu, sRT, o); sN += __shfl_xor_sync(0xffffffffu, sN, o);
    inE += __shfl_xor_sync(0xffffffffu, inE, o); inW += __shfl_xor_sync(0xffffffffu, inW, o); sat += __shfl_xor_sync(0xffffffffu, sat, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // the previous pass' sums have been consumed
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NA; q++) s_part[warp][q] = acc[q];
    s_part[warp][NA] = E; s_part[warp][NA + 1] = sT; s_part[warp][NA + 2] = sRT; s_part[warp][NA + 3] = sN;
    s_parti[warp][0] = inE; s_parti[warp][1] = inW; s_parti[warp][2] = sat;
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    double s = 0;
    for (int w = 0; w < LM_WARPS; w++) s += (double)s_part[w][threadIdx.x];
    s_sum[threadIdx.x] = s;
  } else if (threadIdx.x >= 64 && threadIdx.x < 67) {
    int s = 0;
    for (int w = 0; w < LM_WARPS; w++) s += s_parti[w][threadIdx.x - 64];
    s_cnt[threadIdx.x - 64] = s;
  }
  __syncthreads();
}

// Vec6 of calcRes from the block sums (CoarseTracker.cpp:748-763): {E, numTermsInE, flowT, 0, flowRT, saturated ratio}
__device__ __forceinline__ void lm_res6(const double *s_sum, const int *s_cnt, int NA, double r[6]) {
  const float E = (float)s_sum[NA], sT = (float)s_sum[NA + 1], sRT = (float)s_sum[NA + 2], sN = (float)s_sum[NA + 3];
  r[0] = E; r[1] = s_cnt[0]; r[2] = sT / (sN + 0.1); r[3] = 0; r[4] = sRT / (sN + 0.1); r[5] = s_cnt[2] / (float)s_cnt[0];
}

// accept = mean energy of the candidate < mean energy of the current state (CoarseTracker.cpp:478, ScaleOptimizer.cpp:178).
// The reference adds its energies in one float, in point order; the block sums above are a tree.  Whenever the two means are
// closer than the rounding noise of such sums could explain (LM_TIE), both energies are re-added exactly like the reference
// does (two threads walk the stored terms in point order) and the comparison is made on those: same decision as the
// reference's loop, at the price of a serial pass that only happens on near-ties (typically the last iteration of a level).
constexpr double LM_TIE = 2e-5;
struct LmEnergy {
  float *terms[2];       // [accepted state, candidate] of this hypothesis
  int cur;               // which of the two holds the accepted state
  int old_exact;         // resOld[0] already is the point-order sum
};
__device__ __forceinline__ float lm_seq_sum(const float *__restrict__ t, int n) {
  float s = 0.f;
  int i = 0;
  for (; i + 8 <= n; i += 8) {   // loads in flight, additions strictly in order
    const float4 a = *(const float4 *)(t + i), b = *(const float4 *)(t + i + 4);
    s += a.x; s += a.y; s += a.z; s += a.w; s += b.x; s += b.y; s += b.z; s += b.w;
  }
  for (; i < n; i++) s += t[i];
  return s;
}
// block-wide; resOld / resNew in shared memory, s_x two floats of shared scratch.  Returns the decision to every thread.
__device__ __forceinline__ bool lm_decide(double *resOld, double *resNew, LmEnergy &en, int n, float *s_x, int *s_flag) {
  if (threadIdx.x == 0) {
    const double mo = resOld[0] / resOld[1], mn = resNew[0] / resNew[1];
    const bool tie = fabs(mn - mo) <= LM_TIE * fabs(mo) && resOld[1] > 0 && resNew[1] > 0;
    *s_flag = tie ? 1 : 0;
  }
  __syncthreads();
  if (*s_flag) {
    if (threadIdx.x == 0 && !en.old_exact) s_x[0] = lm_seq_sum(en.terms[en.cur], n);
    if (threadIdx.x == 32) s_x[1] = lm_seq_sum(en.terms[en.cur ^ 1], n);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (!en.old_exact) { resOld[0] = s_x[0]; en.old_exact = 1; }
      resNew[0] = s_x[1];
    }
  }
  __syncthreads();
  const bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
  __syncthreads();
  if (threadIdx.x == 0 && accept) { en.cur ^= 1; en.old_exact = *s_flag; }
  return accept;
}

// R.cast<float>() * Ki (3x3 float product, row by column with the summation order of the reference's expression)
__device__ __forceinline__ void lm_RKi(const double R[9], const float Ki[9], float out[9]) {
  float Rf[9];
  for (int i = 0; i < 9; i++) Rf[i] = (float)R[i];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out[3 * i + j] = (Rf[3 * i] * Ki[j] + Rf[3 * i + 1] * Ki[3 + j]) + Rf[3 * i + 2] * Ki[6 + j];
}

// AffLight::fromToVecExposure (NumType.h:157-168)
__device__ __forceinline__ void lm_affLL(float expF, float expT, const double g2F[2], const double g2T[2], double out[2]) {
  if (expF == 0 || expT == 0) expT = expF = 1;
  const double a = exp(g2T[0] - g2F[0]) * expT / expF;
  out[0] = a; out[1] = g2T[1] - a * g2F[1];
}

// ---- trackNewestCoarse: one block per hypothesis ---------------------------------------------------------------
// everything only the leading thread touches lives in shared memory, so that the registers of the per-point pass stay free
struct LeadPose {
  lm::Pose cur, cand;
  double aff_cur[2], aff_new[2];
  double H[64], b[8], resOld[6], resNew[6], inc[8];
  double Hl[64], W[64], nb[8];   // scratch of the damped solve
  float lambda, levelCutoffRepeat;
  int haveRepeated, pass, iteration;
  unsigned long long accept_mask, tie_mask;
};

// the candidate of the next pass: pose, affine, cutoff -> shared
__device__ __noinline__ void lead_publish(const LmArgs &a, int lvl, const lm::Pose &p, const double aff[2], float cutoff, LmCand &c) {
  double R[9], ll[2];
  lm::quat_matrix(p, R);
  lm_RKi(R, a.lv[lvl].Ki, c.RKi);
  for (int i = 0; i < 3; i++) c.t[i] = (float)p.t[i];
  lm_affLL(a.ref_exp, a.new_exp, a.ref_aff, aff, ll);
  c.aff0 = (float)ll[0]; c.aff1 = (float)ll[1];
  c.scale = 1.f;
  c.cutoffTH = cutoff; c.maxEnergy = 2 * a.huberTH * cutoff - a.huberTH * a.huberTH;
}
// H, b of calcGSSSEPose from the block sums (CoarseTracker.cpp:596-609); called by the first 45 threads, one sum each
__device__ __forceinline__ void take_gs_entry(const double *s_sum, const int *s_cnt, double *H, double *b) {
  const int q = threadIdx.x;
  if (q >= 45) return;
  int r = 0, base = 0;
  while (q >= base + (9 - r)) { base += 9 - r; r++; }
  const int cc = r + (q - base);
  const int n = (s_cnt[1] + 3) / 4 * 4;
  const float rn = 1.0f / n;
  const float sc[9] = {1.0f, 1.0f, 1.0f, 0.5f, 0.5f, 0.5f, 10.0f, 1000.0f, 1.0f};   // SCALE_XI_ROT x3, SCALE_XI_TRANS x3, SCALE_A, SCALE_B
  const double v = (double)(float)s_sum[q] * rn;
  if (cc < 8) { const double w = (v * sc[cc]) * sc[r]; H[8 * r + cc] = w; H[8 * cc + r] = w; }
  else if (r < 8) b[r] = v * sc[r];
}
// the damped solve, the step and the candidate pose of one iteration (CoarseTracker.cpp:419-463); leading thread
__device__ __noinline__ void lead_step(const LmArgs &a, LeadPose &S) {
  const float lambdaExtrapolationLimit = 0.001f;
  double *Hl = S.Hl, *nb = S.nb, *inc = S.inc;
#pragma unroll
  for (int i = 0; i < 64; i++) Hl[i] = S.H[i];
#pragma unroll
  for (int i = 0; i < 8; i++) { Hl[9 * i] *= (1 + S.lambda); nb[i] = -S.b[i]; }
  lm::ldlt_solve_fixed<8>(Hl, S.W, nb, inc);
  if (a.modeA < 0 && a.modeB < 0) {   // fix a, b
    double i6[6];
    for (int r = 0; r < 6; r++) for (int cc = 0; cc < 6; cc++) S.W[6 * r + cc] = Hl[8 * r + cc];
    lm::ldlt_solve<8>(S.W, 6, nb, i6);
    for (int i = 0; i < 6; i++) inc[i] = i6[i];
    inc[6] = inc[7] = 0;
  }
  if (!(a.modeA < 0) && a.modeB < 0) {   // fix b
    double i7[7];
    for (int r = 0; r < 7; r++) for (int cc = 0; cc < 7; cc++) S.W[7 * r + cc] = Hl[8 * r + cc];
    lm::ldlt_solve<8>(S.W, 7, nb, i7);
    for (int i = 0; i < 7; i++) inc[i] = i7[i];
    inc[7] = 0;
  }
  if (a.modeA < 0 && !(a.modeB < 0)) {   // fix a: column / row 7 take the place of 6 (:437-449)
    double bs[8], nb7[7], i7[7];
    for (int i = 0; i < 8; i++) bs[i] = S.b[i];
    for (int r = 0; r < 8; r++) Hl[8 * r + 6] = Hl[8 * r + 7];
    for (int cc = 0; cc < 8; cc++) Hl[8 * 6 + cc] = Hl[8 * 7 + cc];
    bs[6] = bs[7];
    for (int r = 0; r < 7; r++) { for (int cc = 0; cc < 7; cc++) S.W[7 * r + cc] = Hl[8 * r + cc]; nb7[r] = -bs[r]; }
    lm::ldlt_solve<8>(S.W, 7, nb7, i7);
    for (int i = 0; i < 6; i++) inc[i] = i7[i];
    inc[6] = 0; inc[7] = i7[6];
  }
  float extrapFac = 1;
  if (S.lambda < lambdaExtrapolationLimit) extrapFac = sqrt(sqrt(lambdaExtrapolationLimit / S.lambda));
  double incScaled[8];
  const float sc[8] = {1.0f, 1.0f, 1.0f, 0.5f, 0.5f, 0.5f, 10.0f, 1000.0f};
  double ssum = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { inc[i] *= extrapFac; incScaled[i] = inc[i] * sc[i]; ssum += incScaled[i]; }
  if (!isfinite(ssum)) for (int i = 0; i < 8; i++) incScaled[i] = 0;
  S.cand = lm::se3_mul(lm::se3_exp(incScaled), S.cur);
  S.aff_new[0] = S.aff_cur[0] + incScaled[6]; S.aff_new[1] = S.aff_cur[1] + incScaled[7];
}

__global__ void __launch_bounds__(LM_THREADS) k_track_lm(LmArgs a) {
  __shared__ float s_part[LM_WARPS][52];
  __shared__ int s_parti[LM_WARPS][3];
  __shared__ double s_sum[52];
  __shared__ int s_cnt[3];
  __shared__ LmCand s_c;
  __shared__ LeadPose S;
  __shared__ LmEnergy s_en;
  __shared__ float s_x[2];
  __shared__ int s_flag;
  __shared__ int s_go;       // control word written by thread 0: what the block does next
  sosba_track_hypothesis *hy = (sosba_track_hypothesis *)a.hyps + blockIdx.x;
  if (threadIdx.x == 0) {
    s_en.terms[0] = a.terms + (size_t)blockIdx.x * 2 * a.terms_stride; s_en.terms[1] = s_en.terms[0] + a.terms_stride;
    s_en.cur = 0; s_en.old_exact = 0;
  }
  const bool lead = threadIdx.x == 0;
  const int maxIterations[5] = {10, 20, 50, 50, 50};
  const float lambdaExtrapolationLimit = 0.001f;
  const float b0 = (float)a.ref_aff[1];
  if (lead) {
    S.cur.qx = hy->q[0]; S.cur.qy = hy->q[1]; S.cur.qz = hy->q[2]; S.cur.qw = hy->q[3];
    for (int i = 0; i < 3; i++) S.cur.t[i] = hy->t[i];
    S.aff_cur[0] = hy->aff_g2l[0]; S.aff_cur[1] = hy->aff_g2l[1];
    S.haveRepeated = 0;
    for (int i = 0; i < 5; i++) hy->last_residuals[i] = NAN;
    for (int i = 0; i < 3; i++) hy->flow_indicators[i] = 1000;
    hy->ok = 0; hy->n_passes = 0;
    for (int i = 0; i < SOSBA_TRACK_MAX_PASSES; i++) { hy->pass_lvl[i] = 0; hy->pass_iterations[i] = 0; hy->pass_accept[i] = 0; hy->pass_tie[i] = 0; hy->pass_residual[i] = 0; hy->pass_cutoff_repeat[i] = 0; }
  }
  for (int lvl = a.coarsest; lvl >= 0; lvl--) {
    const LmLevel &L = a.lv[lvl];
    if (lead) S.levelCutoffRepeat = 1;
    // resOld = calcResPose(current); double the cutoff while more than 60 % of the terms saturate (:385-399)
    for (;;) {
      if (lead) { lead_publish(a, lvl, S.cur, S.aff_cur, a.coarseCutoffTH * S.levelCutoffRepeat, s_c); s_en.old_exact = 0; }
      __syncthreads();
      lm_pass<0>(L, lvl, s_c, a.huberTH, b0, s_part, s_parti, s_sum, s_cnt, s_en.terms[s_en.cur]);
      if (lead) {
        lm_res6(s_sum, s_cnt, 45, S.resOld);
        s_go = (S.resOld[5] > 0.6 && S.levelCutoffRepeat < 50) ? 1 : 0;
        if (s_go) S.levelCutoffRepeat *= 2;
      }
      __syncthreads();
      if (!s_go) break;
    }
    take_gs_entry(s_sum, s_cnt, S.H, S.b);
    if (lead) {
      S.lambda = 0.01f; S.iteration = 0; S.accept_mask = 0; S.tie_mask = 0;
      S.pass = hy->n_passes < SOSBA_TRACK_MAX_PASSES ? hy->n_passes : SOSBA_TRACK_MAX_PASSES - 1;
    }
    for (int it = 0; it < maxIterations[lvl]; it++) {
      if (lead) {
        lead_step(a, S);
        lead_publish(a, lvl, S.cand, S.aff_new, a.coarseCutoffTH * S.levelCutoffRepeat, s_c);
      }
      __syncthreads();
      lm_pass<0>(L, lvl, s_c, a.huberTH, b0, s_part, s_parti, s_sum, s_cnt, s_en.terms[s_en.cur ^ 1]);
      if (lead) lm_res6(s_sum, s_cnt, 45, S.resNew);
      const bool accept = lm_decide(S.resOld, S.resNew, s_en, L.n, s_x, &s_flag);
      if (accept) take_gs_entry(s_sum, s_cnt, S.H, S.b);   // (read by the leading thread after the barrier below)
      if (lead) {
        if (s_flag && it < 64) S.tie_mask |= 1ull << it;
        if (accept) {
          for (int i = 0; i < 6; i++) S.resOld[i] = S.resNew[i];
          S.aff_cur[0] = S.aff_new[0]; S.aff_cur[1] = S.aff_new[1];
          S.cur = S.cand;
          S.lambda *= 0.5f;
          if (it < 64) S.accept_mask |= 1ull << it;
        } else {
          S.lambda *= 4;
          if (S.lambda < lambdaExtrapolationLimit) S.lambda = lambdaExtrapolationLimit;
        }
        double nrm = 0;
        for (int i = 0; i < 8; i++) nrm += S.inc[i] * S.inc[i];
        S.iteration = it + 1;
        s_go = sqrt(nrm) > 1e-3 ? 1 : 0;
      }
      __syncthreads();
      if (!s_go) break;
    }
    // end of the level (:509-520)
    if (lead) {
      const float lr = sqrtf((float)(S.resOld[0] / S.resOld[1]));
      const int pass = S.pass;
      hy->last_residuals[lvl] = lr;
      hy->flow_indicators[0] = S.resOld[2]; hy->flow_indicators[1] = S.resOld[3]; hy->flow_indicators[2] = S.resOld[4];
      hy->pass_lvl[pass] = lvl; hy->pass_iterations[pass] = S.iteration; hy->pass_accept[pass] = S.accept_mask; hy->pass_tie[pass] = S.tie_mask; hy->pass_residual[pass] = lr;
      hy->pass_cutoff_repeat[pass] = S.levelCutoffRepeat;
      hy->n_passes = hy->n_passes + 1;
      int go = 1;                                      // 1: next level, 2: repeat this level, 0: abort
      if (lr > 1.5 * hy->min_res_for_abort[lvl]) go = 0;
      else if (S.levelCutoffRepeat > 1 && !S.haveRepeated) { go = 2; S.haveRepeated = 1; }
      s_go = go;
    }
    __syncthreads();
    const int go = s_go;
    __syncthreads();
    if (go == 0) return;
    if (go == 2) lvl++;
  }
  if (lead) {   // "set!" and the final checks (:526-549)
    hy->q[0] = S.cur.qx; hy->q[1] = S.cur.qy; hy->q[2] = S.cur.qz; hy->q[3] = S.cur.qw;
    for (int i = 0; i < 3; i++) hy->t[i] = S.cur.t[i];
    hy->aff_g2l[0] = S.aff_cur[0]; hy->aff_g2l[1] = S.aff_cur[1];
    bool ok = true;
    if ((a.modeA != 0 && (fabsf((float)S.aff_cur[0]) > 1.2)) || (a.modeB != 0 && (fabsf((float)S.aff_cur[1]) > 200))) ok = false;
    if (ok) {
      double ll[2];
      lm_affLL(a.ref_exp, a.new_exp, a.ref_aff, S.aff_cur, ll);
      const float r0 = (float)ll[0], r1 = (float)ll[1];
      if ((a.modeA == 0 && (fabsf(logf(r0)) > 1.5)) || (a.modeB == 0 && (fabsf(r1) > 200))) ok = false;
    }
    if (ok) {
      if (a.modeA < 0) hy->aff_g2l[0] = 0;
      if (a.modeB < 0) hy->aff_g2l[1] = 0;
      hy->ok = 1;
    }
  }
}

// ---- optimizeScale: one block per start value --------------------------------------------------------------------
__global__ void __launch_bounds__(LM_THREADS) k_scale_lm(LmArgs a) {
  __shared__ float s_part[LM_WARPS][52];
  __shared__ int s_parti[LM_WARPS][3];
  __shared__ double s_sum[52];
  __shared__ int s_cnt[3];
  __shared__ LmCand s_c;
  __shared__ int s_go;
  __shared__ LmEnergy s_en;
  __shared__ float s_x[2];
  __shared__ int s_flag;
  __shared__ double resOld[6], resNew[6];
  sosba_scale_hypothesis *hy = (sosba_scale_hypothesis *)a.hyps + blockIdx.x;
  const bool lead = threadIdx.x == 0;
  const int maxIterations[5] = {10, 20, 50, 50, 50};
  const float lambdaExtrapolationLimit = 0.001f;
  float scale_current = 0.f, scale_new = 0.f, H = 0.f, b = 0.f;
  bool haveRepeated = false;
  if (threadIdx.x == 0) {
    s_en.terms[0] = a.terms + (size_t)blockIdx.x * 2 * a.terms_stride; s_en.terms[1] = s_en.terms[0] + a.terms_stride;
    s_en.cur = 0; s_en.old_exact = 0;
  }
  if (lead) {
    scale_current = hy->scale;
    for (int i = 0; i < 5; i++) hy->last_residuals[i] = NAN;
    hy->n_passes = 0;
    for (int i = 0; i < SOSBA_TRACK_MAX_PASSES; i++) { hy->pass_lvl[i] = 0; hy->pass_iterations[i] = 0; hy->pass_accept[i] = 0; hy->pass_tie[i] = 0; }
  }
  auto publish = [&](int lvl, float scale, float cutoff) {
    float Rf[9];
    for (int i = 0; i < 9; i++) Rf[i] = a.R10[i];
    const float *Ki = a.lv[lvl].Ki;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) s_c.RKi[3 * i + j] = (Rf[3 * i] * Ki[j] + Rf[3 * i + 1] * Ki[3 + j]) + Rf[3 * i + 2] * Ki[6 + j];
    for (int i = 0; i < 3; i++) s_c.t[i] = a.t10[i];
    s_c.aff0 = 1.f; s_c.aff1 = 0.f; s_c.scale = scale;
    s_c.cutoffTH = cutoff; s_c.maxEnergy = 2 * a.huberTH * cutoff - a.huberTH * a.huberTH;
  };
  auto take_gs = [&]() {   // ScaleOptimizer.cpp:266-270
    const int n = (s_cnt[1] + 3) / 4 * 4;
    H = (float)s_sum[0] * (1.0f / n);
    b = (float)s_sum[1] * (1.0f / n);
  };
  for (int lvl = a.coarsest; lvl >= 0; lvl--) {
    const LmLevel &L = a.lv[lvl];
    float levelCutoffRepeat = 1;
    for (;;) {
      if (lead) { publish(lvl, scale_current, a.coarseCutoffTH * levelCutoffRepeat); s_en.old_exact = 0; }
      __syncthreads();
      lm_pass<1>(L, lvl, s_c, a.huberTH, 0.f, s_part, s_parti, s_sum, s_cnt, s_en.terms[s_en.cur]);
      if (lead) {
        lm_res6(s_sum, s_cnt, 3, resOld);
        s_go = (resOld[5] > 0.6 && levelCutoffRepeat < 50) ? 1 : 0;
        if (s_go) levelCutoffRepeat *= 2;
      }
      __syncthreads();
      if (!s_go) break;
    }
    float lambda = 0.01f;
    int pass = 0, iteration = 0;
    unsigned long long accept_mask = 0, tie_mask = 0;
    if (lead) {
      take_gs();
      pass = hy->n_passes < SOSBA_TRACK_MAX_PASSES ? hy->n_passes : SOSBA_TRACK_MAX_PASSES - 1;
    }
    for (int it = 0; it < maxIterations[lvl]; it++) {
      float inc = 0.f;
      if (lead) {
        float Hl = H;
        Hl *= (1 + lambda);
        inc = -b / Hl;
        float extrapFac = 1;
        if (lambda < lambdaExtrapolationLimit) extrapFac = sqrt(sqrt(lambdaExtrapolationLimit / lambda));
        inc *= extrapFac;
        if (!isfinite(inc) || fabs(inc) > scale_current) inc = 0.0;
        scale_new = scale_current + inc;
        publish(lvl, scale_new, a.coarseCutoffTH * levelCutoffRepeat);
      }
      __syncthreads();
      lm_pass<1>(L, lvl, s_c, a.huberTH, 0.f, s_part, s_parti, s_sum, s_cnt, s_en.terms[s_en.cur ^ 1]);
      if (lead) lm_res6(s_sum, s_cnt, 3, resNew);
      const bool accept = lm_decide(resOld, resNew, s_en, L.n, s_x, &s_flag);
      if (lead) {
        if (s_flag && it < 64) tie_mask |= 1ull << it;
        if (accept) {
          take_gs();
          for (int i = 0; i < 6; i++) resOld[i] = resNew[i];
          scale_current = scale_new;
          lambda *= 0.5f;
          if (it < 64) accept_mask |= 1ull << it;
        } else {
          lambda *= 4;
          if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
        }
        iteration = it + 1;
        s_go = inc > 1e-3 ? 1 : 0;   // sic: the signed increment (ScaleOptimizer.cpp:200)
      }
      __syncthreads();
      if (!s_go) break;
    }
    if (lead) {
      hy->last_residuals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
      hy->pass_lvl[pass] = lvl; hy->pass_iterations[pass] = iteration; hy->pass_accept[pass] = accept_mask; hy->pass_tie[pass] = tie_mask;
      hy->n_passes = hy->n_passes + 1;
      int go = 1;
      if (levelCutoffRepeat > 1 && !haveRepeated) { go = 2; haveRepeated = true; }
      s_go = go;
    }
    __syncthreads();
    const int go = s_go;
    __syncthreads();
    if (go == 2) lvl++;
  }
  if (lead) { hy->scale = scale_current; hy->error = (float)hy->last_residuals[0]; }
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------------------------
#define API extern "C" __attribute__((visibility("default")))
#define LM_CHECK_H(h) do { if (!(h)) { sosba_set_error("null handle"); return SOSBA_E_ARG; } cudaSetDevice((h)->device); } while (0)

namespace {
template <class T> int lm_alloc(sosba *h, T **p, size_t n) {
  void *q = nullptr;
  if (cudaMalloc(&q, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); sosba_set_error("out of device memory (%zu bytes)", n * sizeof(T)); return SOSBA_E_CUDA; }
  h->lm_allocs.push_back(q);
  *p = (T *)q;
  return SOSBA_OK;
}
int lm_ensure_maps(sosba *h) {
  if (h->cd_idepth) return SOSBA_OK;
  const size_t T = h->lvl_off[h->levels], n0 = (size_t)h->wl[0] * h->hl[0];
  int rc;
  if ((rc = lm_alloc(h, &h->cd_idepth, 3 * T))) return rc;   // idepth | weight sums | their copy for the dilation
  h->cd_wsum = h->cd_idepth + T; h->cd_wbak = h->cd_wsum + T;
  if ((rc = lm_alloc(h, &h->cd_cnt, 2 * n0))) return rc;      // hits per pixel | lowest point index
  if ((rc = lm_alloc(h, &h->cd_mark, T))) return rc;
  if ((rc = lm_alloc(h, &h->cd_list, T))) return rc;
  if ((rc = lm_alloc(h, &h->cd_scan, 2 * ((n0 + 1023) / 1024) + 64))) return rc;   // chunk counts | offsets | per-level totals
  return SOSBA_OK;
}
void lm_make_Ki(const float K[4], float Ki[9]) {
  for (int i = 0; i < 9; i++) Ki[i] = 0.f;
  Ki[0] = 1.0f / K[0]; Ki[4] = 1.0f / K[1]; Ki[2] = -K[2] / K[0]; Ki[5] = -K[3] / K[1]; Ki[8] = 1.f;
}
}  // namespace

API int sosba_tracker_make_coarse_depth(sosba_t *h, int32_t ref_slot, int32_t n, const float *cpt, const float *HdiF, int32_t *pc_n_out) {
  LM_CHECK_H(h);
  if (ref_slot < 0 || ref_slot >= (int)h->slot_img.size() || !h->slot_valid[ref_slot] || n < 0 || (n > 0 && (!cpt || !HdiF))) { sosba_set_error("make_coarse_depth: bad slot / arrays"); return SOSBA_E_ARG; }
  for (int i = 0; i < n; i++) {
    const int u = cpt[3 * i] + 0.5f, v = cpt[3 * i + 1] + 0.5f;
    if (!(cpt[3 * i] >= 0) || !(cpt[3 * i + 1] >= 0) || u >= h->wl[0] || v >= h->hl[0]) { sosba_set_error("make_coarse_depth: point %d projects outside the image", i); return SOSBA_E_ARG; }
  }
  int rc = lm_ensure_maps(h);
  if (rc) return rc;
  if ((size_t)n * 4 * sizeof(float) > h->h_pinned_bytes) { sosba_set_error("make_coarse_depth: %d points exceed the staging buffer", n); return SOSBA_E_ARG; }
  const size_t T = h->lvl_off[h->levels], n0 = (size_t)h->wl[0] * h->hl[0];
  cudaStream_t st = h->stream;
  // inputs: one staged copy (centre projections | HdiF) into the tail of the list arena (free until the compaction)
  float *d_in = (float *)(h->cd_list + T / 2);
  if (n > 0) {
    SOSBA_CUDA(cudaStreamSynchronize(st));   // the pinned staging buffer may still feed an earlier copy
    memcpy(h->h_pinned, cpt, sizeof(float) * 3 * (size_t)n);
    memcpy(h->h_pinned + 3 * (size_t)n, HdiF, sizeof(float) * (size_t)n);
    SOSBA_CUDA(cudaMemcpyAsync(d_in, h->h_pinned, sizeof(float) * 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  }
  cudaMemsetAsync(h->cd_idepth, 0, sizeof(float) * 2 * T, st);
  cudaMemsetAsync(h->cd_cnt, 0, sizeof(int) * n0, st);
  cudaMemsetAsync(h->cd_cnt + n0, 0x7f, sizeof(int) * n0, st);
  if (n > 0) {
    k_cd_splat<<<(n + 255) / 256, 256, 0, st>>>(n, d_in, d_in + 3 * (size_t)n, h->wl[0], h->cd_idepth, h->cd_wsum, h->cd_cnt, h->cd_cnt + n0);
    k_cd_fix<<<(n + 255) / 256, 256, 0, st>>>(n, d_in, d_in + 3 * (size_t)n, h->wl[0], h->cd_idepth, h->cd_wsum, h->cd_cnt, h->cd_cnt + n0);
    h->launches += 2;
  }
  for (int l = 1; l < h->levels; l++) {
    const int N = h->wl[l] * h->hl[l];
    k_cd_pool<<<(N + 255) / 256, 256, 0, st>>>(h->wl[l], h->hl[l], h->wl[l - 1], h->cd_idepth + h->lvl_off[l - 1], h->cd_wsum + h->lvl_off[l - 1],
                                              h->cd_idepth + h->lvl_off[l], h->cd_wsum + h->lvl_off[l]);
    h->launches++;
  }
  cudaMemcpyAsync(h->cd_wbak, h->cd_wsum, sizeof(float) * T, cudaMemcpyDeviceToDevice, st);
  int *totals = h->cd_scan + 2 * ((n0 + 1023) / 1024);
  for (int l = 0; l < h->levels; l++) {
    const int N = h->wl[l] * h->hl[l];
    const size_t o = h->lvl_off[l];
    if (N > 2 * h->wl[l]) k_cd_dilate<<<(N - 2 * h->wl[l] + 255) / 256, 256, 0, st>>>(h->wl[l], N, l < 2 ? 1 : 0, h->cd_wbak + o, h->cd_idepth + o, h->cd_wsum + o);
    k_cd_mark<<<(N + 255) / 256, 256, 0, st>>>(h->wl[l], h->hl[l], h->slot_img[ref_slot] + o, h->cd_idepth + o, h->cd_wsum + o, h->cd_mark + o);
    h->launches += 2;
  }
  // the compaction of level l may overwrite the staged inputs only now (all readers are queued in front of it)
  for (int l = 0; l < h->levels; l++) {
    const int N = h->wl[l] * h->hl[l];
    const size_t o = h->lvl_off[l];
    launch_select_compact(h, h->cd_mark + o, N, h->cd_scan, h->cd_scan + (n0 + 1023) / 1024, totals + l, N, h->cd_list + o);
  }
  int counts[SOSBA_MAX_LEVELS];
  SOSBA_CUDA(cudaMemcpyAsync(counts, totals, sizeof(int) * h->levels, cudaMemcpyDeviceToHost, st));
  SOSBA_CUDA(cudaStreamSynchronize(st));   // the one synchronisation: the host needs the list lengths
  for (int l = 0; l < h->levels; l++) {
    const int m = counts[l];
    if ((rc = sosba_tracker_reserve(h, l, m))) return rc;
    h->t_n[l] = m;
    if (m > 0) {
      const size_t o = h->lvl_off[l];
      k_cd_gather<<<(m + 255) / 256, 256, 0, st>>>(m, h->wl[l], h->cd_list + o, h->cd_idepth + o, h->slot_img[ref_slot] + o, h->t_pc[l]);
      h->launches++;
    }
    if (pc_n_out) pc_n_out[l] = m;
  }
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

API int sosba_tracker_get_ref(sosba_t *h, int32_t lvl, int32_t *n, float *u, float *v, float *id, float *c) {
  LM_CHECK_H(h);
  if (lvl < 0 || lvl >= h->levels) return SOSBA_E_ARG;
  const int m = h->t_n[lvl];
  if (n) *n = m;
  float *dst[4] = {u, v, id, c};
  for (int k = 0; k < 4; k++)
    if (dst[k] && m > 0) SOSBA_CUDA(cudaMemcpyAsync(dst[k], h->t_pc[lvl] + (size_t)k * m, sizeof(float) * m, cudaMemcpyDeviceToHost, h->stream));
  SOSBA_CUDA(cudaStreamSynchronize(h->stream));
  return SOSBA_OK;
}

API int sosba_tracker_scale_coarse_depth(sosba_t *h, float scale) {
  LM_CHECK_H(h);
  for (int l = 0; l < h->levels; l++) {
    const int m = h->t_n[l];
    if (m > 0) { k_cd_scale<<<(m + 255) / 256, 256, 0, h->stream>>>(m, h->t_pc[l] + 2 * (size_t)m, scale); h->launches++; }
  }
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

namespace {
// levels, intrinsics, reference lists -> kernel arguments; K1 != nullptr: project into camera 1 (scale optimiser)
int lm_fill_args(sosba *h, int slot, int coarsest, bool stereo, LmArgs *a) {
  if (!h->t_haveK) { sosba_set_error("tracker_make_k first"); return SOSBA_E_STATE; }
  if (stereo && !h->t_haveStereo) { sosba_set_error("scale_set_stereo first"); return SOSBA_E_STATE; }
  if (slot < 0 || slot >= (int)h->slot_img.size() || !h->slot_valid[slot] || coarsest < 0 || coarsest >= h->levels || coarsest >= 5) { sosba_set_error("bad slot / level"); return SOSBA_E_ARG; }
  memset(a, 0, sizeof(*a));
  a->levels = h->levels; a->coarsest = coarsest;
  for (int l = 0; l <= coarsest; l++) {
    LmLevel &L = a->lv[l];
    L.w = h->wl[l]; L.h = h->hl[l]; L.n = h->t_n[l]; L.pc = h->t_pc[l];
    if (L.n > 0 && !L.pc) { sosba_set_error("no reference points on level %d", l); return SOSBA_E_STATE; }
    L.img = h->slot_img[slot] + h->lvl_off[l];
    for (int k = 0; k < 4; k++) L.K[k] = stereo ? h->t_K1[l][k] : h->t_K[l][k];
    lm_make_Ki(h->t_K[l], L.Ki);
  }
  a->huberTH = h->cfg.huber_th; a->coarseCutoffTH = h->cfg.coarse_cutoff_th; a->modeA = h->cfg.affine_opt_mode_a; a->modeB = h->cfg.affine_opt_mode_b;
  return SOSBA_OK;
}
int lm_run(sosba *h, bool stereo, LmArgs &a, void *hyps, size_t bytes, int n_hyp) {
  if (bytes > h->lm_hyp_cap) {
    void *p = nullptr;
    int rc = lm_alloc(h, (unsigned char **)&p, bytes + bytes / 2);
    if (rc) return rc;
    h->lm_hyp = p; h->lm_hyp_cap = bytes + bytes / 2;
  }
  if (bytes > h->h_pinned_bytes) { sosba_set_error("too many hypotheses for one call"); return SOSBA_E_ARG; }
  {  // energy terms of the accepted state and the candidate, per hypothesis (lm_decide)
    int nmax = 4;
    for (int l = 0; l <= a.coarsest; l++) nmax = a.lv[l].n > nmax ? a.lv[l].n : nmax;
    const int stride = (nmax + 3) / 4 * 4;
    const size_t need = (size_t)n_hyp * 2 * stride;
    if (need > h->lm_terms_cap) {
      float *p = nullptr;
      int rc = lm_alloc(h, &p, need + need / 4);
      if (rc) return rc;
      h->lm_terms = p; h->lm_terms_cap = need + need / 4;
    }
    a.terms = h->lm_terms; a.terms_stride = stride;
  }
  cudaStream_t st = h->stream;
  SOSBA_CUDA(cudaStreamSynchronize(st));
  memcpy(h->h_pinned, hyps, bytes);
  SOSBA_CUDA(cudaMemcpyAsync(h->lm_hyp, h->h_pinned, bytes, cudaMemcpyHostToDevice, st));
  a.hyps = h->lm_hyp;
  if (stereo) k_scale_lm<<<n_hyp, LM_THREADS, 0, st>>>(a);
  else k_track_lm<<<n_hyp, LM_THREADS, 0, st>>>(a);
  h->launches++;
  SOSBA_CUDA(cudaGetLastError());
  SOSBA_CUDA(cudaMemcpyAsync(h->h_pinned, h->lm_hyp, bytes, cudaMemcpyDeviceToHost, st));
  SOSBA_CUDA(cudaStreamSynchronize(st));   // the one synchronisation of the call
  memcpy(hyps, h->h_pinned, bytes);
  return SOSBA_OK;
}
}  // namespace

API int sosba_tracker_track(sosba_t *h, int32_t new_slot, float ref_ab_exposure, float new_ab_exposure, const double ref_aff_g2l[2], int32_t coarsest_lvl,
                            int32_t n_hyp, sosba_track_hypothesis *hyps) {
  LM_CHECK_H(h);
  if (n_hyp < 0 || (n_hyp > 0 && !hyps) || !ref_aff_g2l) return SOSBA_E_ARG;
  if (n_hyp == 0) return SOSBA_OK;
  LmArgs a;
  int rc = lm_fill_args(h, new_slot, coarsest_lvl, false, &a);
  if (rc) return rc;
  a.ref_exp = ref_ab_exposure; a.new_exp = new_ab_exposure; a.ref_aff[0] = ref_aff_g2l[0]; a.ref_aff[1] = ref_aff_g2l[1];
  return lm_run(h, false, a, hyps, sizeof(sosba_track_hypothesis) * (size_t)n_hyp, n_hyp);
}

API int sosba_scale_optimize(sosba_t *h, int32_t stereo_slot, int32_t coarsest_lvl, int32_t n_hyp, sosba_scale_hypothesis *hyps) {
  LM_CHECK_H(h);
  if (n_hyp < 0 || (n_hyp > 0 && !hyps)) return SOSBA_E_ARG;
  if (n_hyp == 0) return SOSBA_OK;
  LmArgs a;
  int rc = lm_fill_args(h, stereo_slot, coarsest_lvl, true, &a);
  if (rc) return rc;
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) a.R10[3 * i + j] = (float)h->t_T10[4 * i + j]; a.t10[i] = (float)h->t_T10[4 * i + 3]; }
  return lm_run(h, true, a, hyps, sizeof(sosba_scale_hypothesis) * (size_t)n_hyp, n_hyp);
}

API int sosba_distance_map(sosba_t *h, int32_t nhosts, const float *KRKi, const float *Kt, int32_t n, const int32_t *host, const float *u, const float *v,
                           const float *idepth, float *dist_out) {
  LM_CHECK_H(h);
  if (h->levels < 2 || nhosts < 0 || n < 0 || !dist_out || (n > 0 && (!KRKi || !Kt || !host || !u || !v || !idepth))) { sosba_set_error("distance_map: bad arguments"); return SOSBA_E_ARG; }
  for (int i = 0; i < n; i++) if (host[i] < 0 || host[i] >= nhosts) { sosba_set_error("distance_map: host %d of point %d", host[i], i); return SOSBA_E_ARG; }
  const int w1 = h->wl[1], h1 = h->hl[1], N = w1 * h1;
  const size_t in_floats = 4 * (size_t)n + 12 * (size_t)nhosts;
  if (in_floats * 4 > h->h_pinned_bytes || (size_t)N * 4 > h->h_pinned_bytes) { sosba_set_error("distance_map: %d points exceed the staging buffer", n); return SOSBA_E_ARG; }
  if (in_floats + N > h->dm_cap) {
    float *p = nullptr;
    int rc = lm_alloc(h, &p, (in_floats + N) * 2 + (size_t)N);
    if (rc) return rc;
    h->dm_buf = p; h->dm_cap = (in_floats + N) * 2;
  }
  cudaStream_t st = h->stream;
  SOSBA_CUDA(cudaStreamSynchronize(st));
  float *pin = h->h_pinned;
  memcpy(pin, host, 4 * (size_t)n); memcpy(pin + n, u, 4 * (size_t)n); memcpy(pin + 2 * (size_t)n, v, 4 * (size_t)n); memcpy(pin + 3 * (size_t)n, idepth, 4 * (size_t)n);
  memcpy(pin + 4 * (size_t)n, KRKi, 36 * (size_t)nhosts); memcpy(pin + 4 * (size_t)n + 9 * (size_t)nhosts, Kt, 12 * (size_t)nhosts);
  float *d_in = h->dm_buf, *d_dist = h->dm_buf + in_floats;
  uint8_t *d_seed = (uint8_t *)(d_dist + N);
  if (in_floats) SOSBA_CUDA(cudaMemcpyAsync(d_in, pin, in_floats * 4, cudaMemcpyHostToDevice, st));
  cudaMemsetAsync(d_seed, 0, N, st);
  if (n > 0) {
    k_dm_seed<<<(n + 255) / 256, 256, 0, st>>>(n, (const int *)d_in, d_in + n, d_in + 2 * (size_t)n, d_in + 3 * (size_t)n, d_in + 4 * (size_t)n,
                                              d_in + 4 * (size_t)n + 9 * (size_t)nhosts, w1, h1, d_seed);
    h->launches++;
  }
  k_dm_flood<<<dim3((w1 + DM_TILE - 1) / DM_TILE, (h1 + DM_TILE - 1) / DM_TILE), 1024, 0, st>>>(w1, h1, d_seed, d_dist);
  h->launches++;
  SOSBA_CUDA(cudaGetLastError());
  SOSBA_CUDA(cudaMemcpyAsync(pin, d_dist, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  SOSBA_CUDA(cudaStreamSynchronize(st));
  memcpy(dist_out, pin, (size_t)N * 4);
  return SOSBA_OK;
}
