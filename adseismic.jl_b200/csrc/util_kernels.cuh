// util_kernels.cuh -- small kernels shared by the acoustic and elastic plans (one copy per translation unit).
#pragma once
#include "common.cuh"

// res = 2 (rcvv - obs) for the receivers in `mask` (owned), 0 elsewhere; loss = sum (rcvv-obs)^2 (owned only).
// One CTA, fixed summation order -> deterministic.
static __global__ void k_residual_loss(const double* __restrict__ rcvv, const double* __restrict__ obs,
                                const unsigned char* __restrict__ owned, int ncol, int col_is_fast, i64 n,
                                double* __restrict__ res, double* __restrict__ loss) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (i64 k = threadIdx.x; k < n; k += blockDim.x) {
    const int r = col_is_fast ? (int)(k % ncol) : (int)(k / (n / ncol));
    double d = 0.0;
    if (owned[r]) d = rcvv[k] - obs[k];
    res[k] = 2.0 * d;
    acc += d * d;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = sh[0];
}

