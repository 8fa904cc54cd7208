// fp64 pipe probe: DFMA latency (dependent chain, 1 warp) and throughput (independent chains, many warps), one SM and full chip.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *o, int n, long long *cyc) {
  double a = o[0], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = fma(a, b, c);
  long long t1 = clock64();
  o[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void thr(double *o, int n, long long *cyc) {
  double a0 = o[0], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
  __syncthreads();
  long long t1 = clock64();
  o[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7; if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void thr32(float *o, int n, long long *cyc) {
  float a0 = o[0], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, b = 1.0000001f, c = 1e-9f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c); a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c); }
  __syncthreads();
  long long t1 = clock64();
  o[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7; if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dmma(double *o, int n, long long *cyc) {
  double a = o[0], b = 1.0000001, c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(f0), "+d"(f1) : "d"(a), "d"(b));
  }
  __syncthreads();
  long long t1 = clock64();
  o[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1; if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *o; long long *c, h; float *of;
  cudaMalloc(&o, 1 << 24); cudaMalloc(&of, 1 << 24); cudaMalloc(&c, 8); cudaMemset(o, 0, 1 << 24);
  const int n = 4096;
  lat<<<1, 32>>>(o, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DFMA latency: %.2f cycles\n", (double)h / n);
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    thr<<<1, 32 * warps>>>(o, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 1 SM, %2d warps: %.2f lane-FMA/clk/SM\n", warps, (double)n * 8 * 32 * warps / h);
  }
  thr32<<<1, 1024>>>(of, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("FFMA 1 SM, 32 warps: %.2f lane-FMA/clk/SM\n", (double)n * 8 * 1024 / h);
  for (int warps : {1, 4, 8, 16}) {
    dmma<<<1, 32 * warps>>>(o, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DMMA m8n8k4 1 SM, %2d warps: %.2f FMA/clk/SM (%.1f cycles per mma per warp-slot)\n", warps, (double)n * 4 * 256 * warps / h, (double)h / (n * 4));
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  thr<<<148 * 4, 512>>>(o, n, c); cudaEventRecord(e0); thr<<<148 * 4, 512>>>(o, n, c); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); printf("chip DFMA: %.2f TFLOP/s\n", 2.0 * n * 8 * 148 * 4 * 512 / (ms * 1e-3) / 1e12);
  dmma<<<148 * 4, 512>>>(o, n, c); cudaEventRecord(e0); dmma<<<148 * 4, 512>>>(o, n, c); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1); printf("chip DMMA: %.2f TFLOP/s\n", 2.0 * n * 4 * 256 * 148 * 4 * 16 / (ms * 1e-3) / 1e12);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
