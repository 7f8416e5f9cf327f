// Does fp64 arithmetic on B200 slow down on denormal operands / results?  One warp-wide dependent chain per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o scripts/denormal_bench scripts/denormal_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, double x0, double m, double a, int n) {
  double x = x0 + threadIdx.x * 0.0, y = x0 * 0.5;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    x = x * m + a;   // DMUL + DADD (no contraction)
    y = y * m + a;
    x = x * m - a;
    y = y * m - a;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}
int main() {
  double* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct { const char* name; double x0, m, a; } cases[] = {
    {"normal operands       ", 1.0, 1.0000001, 1e-3},
    {"denormal x, normal m  ", 1e-310, 1.0, 0.0},
    {"denormal x and a      ", 1e-310, 1.0, 3e-320},
    {"tiny normal x (1e-300)", 1e-300, 1.0, 0.0},
    {"zeros                 ", 0.0, 1.0, 0.0}};
  const int n = 200000;
  for (auto& c : cases) {
    k<<<148 * 8, 256>>>(d, c.x0, c.m, c.a, 1000);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<148 * 8, 256>>>(d, c.x0, c.m, c.a, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%s: %.3f ms  (%.2f Gop/s per SM)\n", c.name, ms, 8.0 * n * 8 * 256 / ms / 1e6);
  }
  return 0;
}
