// k_tracker.cu — normal equations of the direct-alignment steps.
//
//   pose   CoarseTracker::calcGSSSEPose + Accumulator9::updateSSE_eighted   CoarseTracker.cpp:554-610, MatrixAccumulators.h:1314-1432
//   scale  ScaleOptimizer::calcGSSSEScale + ScaleAccumulator::updateSSE_oneed   ScaleOptimizer.cpp:232-271, ScaleAccumulator.h:60-77
//
// One thread per warped point (rows rejected by calcRes carry weight 0), 45 (resp. 3) upper-triangle products per
// thread, warp-shuffle tree, one fp64 red.global per entry per block.
#include "kernels.h"

namespace {

template <int NV>
__device__ __forceinline__ void block_reduce_to(float (&v)[NV], double *__restrict__ dst) {
  __shared__ float s_part[8][NV];
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
}

__global__ void __launch_bounds__(256) k_track_gs_pose(TrackGSArgs a) {
  float acc[45];
#pragma unroll
  for (int q = 0; q < 45; q++) acc[q] = 0.f;
  const size_t c = a.cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const float w = a.warp[6 * c + i];
    if (w == 0.f) continue;
    const float id = a.warp[i], u = a.warp[c + i], v = a.warp[2 * c + i];
    const float dx = a.warp[3 * c + i] * a.fx, dy = a.warp[4 * c + i] * a.fy;
    float J[9];
    J[0] = id * dx;
    J[1] = id * dy;
    J[2] = 0 - id * (u * dx + v * dy);
    J[3] = 0 - ((u * v) * dx + dy * (1 + v * v));
    J[4] = (u * v) * dy + dx * (1 + u * u);
    J[5] = u * dy - v * dx;
    J[6] = a.a * (a.b0 - a.warp[7 * c + i]);
    J[7] = -1;
    J[8] = a.warp[5 * c + i];
    int q = 0;
#pragma unroll
    for (int r = 0; r < 9; r++) {
      const float Jw = J[r] * w;
#pragma unroll
      for (int cc = r; cc < 9; cc++) acc[q++] += Jw * J[cc];
    }
  }
  block_reduce_to<45>(acc, a.acc + 8);
}

__global__ void __launch_bounds__(256) k_track_gs_scale(TrackGSArgs a) {
  float acc[3] = {0.f, 0.f, 0.f};
  const size_t c = a.cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const float w = a.warp[6 * c + i];
    if (w == 0.f) continue;
    const float rx1 = a.warp[i], rx2 = a.warp[c + i], rx3 = a.warp[2 * c + i];
    const float dxfx = a.warp[3 * c + i] * a.fx, dyfy = a.warp[4 * c + i] * a.fy;
    const float deno_sqrt = a.scale * rx3 + a.tz;
    const float deno = 1.0f / (deno_sqrt * deno_sqrt);
    const float xno = rx1 * a.tz - rx3 * a.tx, yno = rx2 * a.tz - rx3 * a.ty;
    const float J0 = dxfx * (deno * xno) + dyfy * (deno * yno), J1 = a.warp[5 * c + i];
    const float J0w = J0 * w, J1w = J1 * w;
    acc[0] += J0w * J0; acc[1] += J0w * J1; acc[2] += J1w * J1;
  }
  block_reduce_to<3>(acc, a.acc + 8);
}

}  // namespace

void launch_track_gs(sosba *h, const TrackGSArgs &a) {
  if (a.n == 0) return;
  int blocks = (a.n + 255) / 256;
  if (blocks > 2 * h->sm_count) blocks = 2 * h->sm_count;
  if (a.kind == 0) k_track_gs_pose<<<blocks, 256, 0, h->stream>>>(a);
  else k_track_gs_scale<<<blocks, 256, 0, h->stream>>>(a);
  h->launches++;
}
