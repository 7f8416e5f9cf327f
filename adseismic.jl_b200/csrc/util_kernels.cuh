// util_kernels.cuh -- small kernels shared by the acoustic and elastic plans (one copy per translation unit).
#pragma once
#include "common.cuh"

// res = 2 (rcvv - obs) for the receivers in `mask` (owned), 0 elsewhere; loss = sum (rcvv-obs)^2 (owned only).
// Two passes with a FIXED decomposition (RL_BLOCKS contiguous chunks, fixed thread striding and tree order), so the
// sum is deterministic and independent of the GPU; `loss` points to 1 + RL_BLOCKS doubles (loss, partial sums).
// (A single-CTA version of this took 14 ms for the 5001 x 4058 traces of the C4 workload: 6 % of an 8-GPU gradient.)
#define RL_BLOCKS 1024
static __global__ void __launch_bounds__(256)
k_residual_partial(const double* __restrict__ rcvv, const double* __restrict__ obs,
                   const unsigned char* __restrict__ owned, int ncol, int col_is_fast, i64 n,
                   double* __restrict__ res, double* __restrict__ loss) {
  __shared__ double sh[256];
  const i64 per = (n + RL_BLOCKS - 1) / RL_BLOCKS;
  const i64 k0 = (i64)blockIdx.x * per, k1 = (k0 + per < n) ? k0 + per : n;
  const i64 nrow = n / ncol;  // elements per receiver when the receiver index is the slow one
  double acc = 0.0;
  for (i64 k = k0 + threadIdx.x; k < k1; k += 256) {
    int r;
    if (n < 2147483647LL) r = col_is_fast ? (int)((unsigned)k % (unsigned)ncol) : (int)((unsigned)k / (unsigned)nrow);
    else r = col_is_fast ? (int)(k % ncol) : (int)(k / nrow);
    double d = 0.0;
    if (owned[r]) d = rcvv[k] - obs[k];
    res[k] = 2.0 * d;
    acc += d * d;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[1 + blockIdx.x] = sh[0];
}
static __global__ void __launch_bounds__(RL_BLOCKS) k_residual_final(double* __restrict__ loss) {
  __shared__ double sh[RL_BLOCKS];
  sh[threadIdx.x] = loss[1 + threadIdx.x];
  __syncthreads();
  for (int s = RL_BLOCKS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = sh[0];
}
