// latency probes for reciprocal seeds on sm_100a: MUFU.RCP64H, fp32 MUFU.RCP + conversions, __drcp_rn, 1.0/x, integer-seed Newton
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int MODE>
__global__ void lat(double *o, long long *cyc) {
  double a = o[0] + 1.5;
  long long t0 = clock64();
  for (int i = 0; i < N; i++) {
    if (MODE == 0) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r; }
    if (MODE == 1) { a = __drcp_rn(a); }
    if (MODE == 2) { a = 1.0 / a; }
    if (MODE == 3) { float f = (float)a; float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f)); a = (double)r; }
    if (MODE == 4) { long long b = __double_as_longlong(a); double r = __longlong_as_double(0x7FDE6238502484BAll - b); 
                     double e = fma(-a, r, 1.0); r = fma(r, e, r); e = fma(-a, r, 1.0); r = fma(r, e, r); e = fma(-a, r, 1.0); r = fma(r, e, r);
                     e = fma(-a, r, 1.0); r = fma(r, e, r); e = fma(-a, r, 1.0); r = fma(r, e, r); a = r; }
    if (MODE == 5) { a = fma(a, 1.0000001, 1e-9); }
    if (MODE == 6) { a = sqrt(a) + 1.0; }
    if (MODE == 7) { float f = (float)a; a = (double)f; }
    if (MODE == 8) { a = (a > 1.0) ? a * 0.5 : a * 2.0; }   // DSETP + select chain
  }
  long long t1 = clock64();
  o[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *o; long long *c, h; cudaMalloc(&o, 4096); cudaMalloc(&c, 8); cudaMemset(o, 0, 4096);
  const char *names[] = {"rcp.approx.ftz.f64 (MUFU.RCP64H)", "__drcp_rn", "1.0/x", "f64->f32, MUFU.RCP, f32->f64", "integer seed + 5 Newton", "DFMA", "sqrt(x)+1", "f64->f32->f64", "DSETP+select+DMUL"};
#define RUN(M) lat<M><<<1, 32>>>(o, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-36s %.1f cycles/iter\n", names[M], (double)h / N);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
