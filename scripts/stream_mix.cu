// stream_mix.cu -- development probe: HBM throughput of a pure streaming kernel as a function of the number of
// read and write streams (NR reads + NW writes of distinct 134 MB arrays, NRW read-modify-write arrays), to
// calibrate what the 5R:2W mix of the acoustic adjoint step (and the 10-18 stream elastic passes) can reach.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/stream_mix scripts/stream_mix.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define MAXS 20
struct Ptrs { double* p[MAXS]; };

// each CTA walks `rows` rows of a 512-column tile (like the marching CTAs): thread = double2 column pair
template <int NR, int NW, int NRW>
__global__ void __launch_bounds__(256) k_mix(Ptrs a, int ld, int H, int rb, int nct) {
  const int ct = blockIdx.x % nct, tr = blockIdx.x / nct;
  const int j = ct * 512 + threadIdx.x * 2;
  const int r0 = tr * rb, r1 = min(H, r0 + rb);
  if (j >= ld) return;
  for (int r = r0; r < r1; r++) {
    const size_t o = (size_t)r * ld + j;
    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < NR; k++) {
      const double2 v = *reinterpret_cast<const double2*>(a.p[k] + o);
      acc.x += v.x; acc.y += v.y;
    }
#pragma unroll
    for (int k = 0; k < NRW; k++) {
      double2 v = *reinterpret_cast<const double2*>(a.p[NR + k] + o);
      v.x += acc.x; v.y += acc.y;
      *reinterpret_cast<double2*>(a.p[NR + k] + o) = v;
    }
#pragma unroll
    for (int k = 0; k < NW; k++) *reinterpret_cast<double2*>(a.p[NR + NRW + k] + o) = acc;
  }
}

template <int NR, int NW, int NRW>
static void run(Ptrs a, int ld, int H, int rb) {
  const int nct = (ld + 511) / 512, ntr = (H + rb - 1) / rb;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int k = 0; k < 3; k++) k_mix<NR, NW, NRW><<<nct * ntr, 256>>>(a, ld, H, rb, nct);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int k = 0; k < reps; k++) k_mix<NR, NW, NRW><<<nct * ntr, 256>>>(a, ld, H, rb, nct);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)(NR + NW + 2 * NRW) * ld * H * 8.0;
  printf("R=%d W=%d RW=%d rb=%3d : %8.1f us  %7.1f GB/s  (%s)\n", NR, NW, NRW, rb, ms / reps * 1e3,
         bytes / (ms / reps * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int ld = 4112, H = 4098;
  Ptrs a;
  for (int k = 0; k < MAXS; k++) {
    if (cudaMalloc(&a.p[k], (size_t)ld * H * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(a.p[k], 0, (size_t)ld * H * 8);
  }
  for (int rb : {28, 56}) {
    run<1, 1, 0>(a, ld, H, rb);
    run<2, 1, 0>(a, ld, H, rb);
    run<3, 1, 0>(a, ld, H, rb);   // acoustic forward
    run<4, 1, 0>(a, ld, H, rb);
    run<5, 2, 0>(a, ld, H, rb);   // acoustic adjoint as 7 streams
    run<4, 1, 1>(a, ld, H, rb);   // acoustic adjoint: G as read-modify-write
    run<3, 0, 2>(a, ld, H, rb);   // ... and ubar[s+1] -> ubar[s-1] in place
    run<8, 0, 0>(a, ld, H, rb);
    run<0, 4, 0>(a, ld, H, rb);
    run<8, 3, 0>(a, ld, H, rb);   // elastic sigma pass
    run<5, 0, 3>(a, ld, H, rb);   // elastic sigma pass in place
    run<7, 2, 0>(a, ld, H, rb);   // elastic velocity pass
    run<5, 0, 2>(a, ld, H, rb);
    run<10, 0, 5>(a, ld, H, rb);
  }
  return 0;
}
