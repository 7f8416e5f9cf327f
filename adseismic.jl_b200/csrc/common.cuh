// common.cuh -- shared host/device helpers of libadseis_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/adseis.h"

#define ADSEIS_API extern "C" __attribute__((visibility("default")))

typedef long long i64;

// ---------------------------------------------------------------------------------------------------------
// error handling: every CUDA call is checked; the text is kept per thread for adseis_last_error()
// ---------------------------------------------------------------------------------------------------------
void adseis_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      adseis_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
      return (_e == cudaErrorMemoryAllocation) ? ADSEIS_ENOMEM : ADSEIS_ECUDA;                     \
    }                                                                                             \
  } while (0)

#define TRY(expr)                 \
  do {                            \
    int _r = (expr);              \
    if (_r != ADSEIS_OK) return _r; \
  } while (0)

#define REQUIRE(cond, ...)              \
  do {                                  \
    if (!(cond)) {                      \
      adseis_set_error(__VA_ARGS__);    \
      return ADSEIS_EINVAL;             \
    }                                   \
  } while (0)

struct adseis_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  i64 launches;
};

// RAII-free tiny device buffer helper (plans free explicitly in destroy)
template <typename T>
static inline int dev_alloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CUDA_TRY(cudaMalloc((void**)p, n * sizeof(T)));
  return ADSEIS_OK;
}
template <typename T>
static inline int dev_alloc_zero(T** p, size_t n, cudaStream_t s) {
  TRY(dev_alloc(p, n));
  CUDA_TRY(cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(T), s));
  return ADSEIS_OK;
}
template <typename T>
static inline int dev_upload(T** p, const std::vector<T>& h, cudaStream_t s) {
  TRY(dev_alloc(p, h.size()));
  if (!h.empty()) CUDA_TRY(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return ADSEIS_OK;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------
// Point sets: sources and receivers grouped by grid cell so that the time-step kernels can inject / sample
// them in their tile epilogue without atomics and in the reference's sequential order
// (ScatterAddOps.h:3-6, AddSource.cpp:57-85 add in source order; duplicates on one cell are legal).
//   key[k]   = coltile * plane + cell   (sorted ascending; one entry per UNIQUE (cell[,field]) )
//   cell[k]  = local flat offset row*ld + col
//   start[k]..start[k+1] = range into perm[] of the points that sit on that cell, in original order
//   perm[m]  = original (global) point index
// ---------------------------------------------------------------------------------------------------------
struct PointSetHost {
  std::vector<i64> key;
  std::vector<int> cell, start, perm, field;
};
struct PointSetDev {
  int nu;            // unique cells
  int npts;          // points
  const i64* key;
  const int* cell;
  const int* start;
  const int* perm;
  const int* field;  // elastic: field id per unique entry (else null)
};

// rows/cols are local row and column of each kept point, gid its global index; field may be empty.
static inline void build_point_set(const std::vector<int>& rows, const std::vector<int>& cols,
                                   const std::vector<int>& gid, const std::vector<int>& field, int ld, i64 plane,
                                   int tile_cols, PointSetHost* out) {
  size_t n = rows.size();
  std::vector<size_t> order(n);
  for (size_t k = 0; k < n; k++) order[k] = k;
  auto keyof = [&](size_t k) -> i64 {
    i64 cell = (i64)rows[k] * ld + cols[k];
    i64 f = field.empty() ? 0 : field[k];
    return ((i64)(cols[k] / tile_cols) * plane + cell) * 8 + f;
  };
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keyof(a) < keyof(b); });
  out->key.clear(); out->cell.clear(); out->start.clear(); out->perm.clear(); out->field.clear();
  for (size_t m = 0; m < n; m++) {
    size_t k = order[m];
    i64 key = keyof(k);
    if (out->key.empty() || out->key.back() != key) {
      out->key.push_back(key);
      out->cell.push_back(rows[k] * ld + cols[k]);
      out->field.push_back(field.empty() ? 0 : field[k]);
      out->start.push_back((int)m);
    }
    out->perm.push_back(gid[k]);
  }
  out->start.push_back((int)n);
}

struct PointSetStorage {
  i64* key = nullptr;
  int *cell = nullptr, *start = nullptr, *perm = nullptr, *field = nullptr;
  PointSetDev dev{};
};
static inline int upload_point_set(const PointSetHost& h, PointSetStorage* st, cudaStream_t s) {
  TRY(dev_upload(&st->key, h.key, s));
  TRY(dev_upload(&st->cell, h.cell, s));
  TRY(dev_upload(&st->start, h.start, s));
  TRY(dev_upload(&st->perm, h.perm, s));
  TRY(dev_upload(&st->field, h.field, s));
  st->dev.nu = (int)h.key.size();
  st->dev.npts = (int)h.perm.size();
  st->dev.key = st->key; st->dev.cell = st->cell; st->dev.start = st->start; st->dev.perm = st->perm;
  st->dev.field = st->field;
  return ADSEIS_OK;
}
static inline void free_point_set(PointSetStorage* st) {
  cudaFree(st->key); cudaFree(st->cell); cudaFree(st->start); cudaFree(st->perm); cudaFree(st->field);
  *st = PointSetStorage();
}

#ifdef __CUDACC__
// first index k in [0,n) with key[k] >= v
__device__ __forceinline__ int ps_lower_bound(const i64* __restrict__ key, int n, i64 v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (key[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// entries of `ps` whose cell lies in rows [r0,r1) of column tile `ct`  ->  [*a,*b)
__device__ __forceinline__ void ps_range(const PointSetDev& ps, int ct, int r0, int r1, int ld, i64 plane, int* a,
                                         int* b) {
  if (ps.nu == 0) { *a = 0; *b = 0; return; }
  i64 base = (i64)ct * plane;
  *a = ps_lower_bound(ps.key, ps.nu, (base + (i64)r0 * ld) * 8);
  *b = ps_lower_bound(ps.key, ps.nu, (base + (i64)r1 * ld) * 8);
}

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
// streaming (evict-first) 16-byte accesses for data touched once per time step
__device__ __forceinline__ double2 ld2_stream(const double* p) { return __ldcs(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st2_stream(double* p, double2 v) { __stcs(reinterpret_cast<double2*>(p), v); }
#endif
