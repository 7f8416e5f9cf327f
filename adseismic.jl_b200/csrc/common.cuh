// common.cuh -- shared host/device helpers of libadseis_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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
  int plans;    // live plans bound to this context
  bool zombie;  // adseis_ctx_destroy was called while plans were alive: destruction happens with the last plan
};
void adseis_ctx_release_plan(adseis_ctx* ctx);  // ctx.cu: called by every plan destructor (last statement)

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

// Programmatic dependent launch (sm_90+): the time-step kernels call griddepcontrol.launch_dependents on entry, so
// the next step's CTAs are scheduled while the current step drains, run their prologue (CTA table, mbarrier init)
// and block in griddepcontrol.wait until the current step's memory operations are complete and visible.  Hides the
// launch gap between the thousands of dependent step launches of a sweep.  ADSEIS_PDL=0 disables it.  Slab plans
// launch without it (`pdl` = false): measured on 2 and 8 B200s, early-resident CTAs of the next step make the
// single-wave slab launches 10-25 % slower (forward 16.5 -> 22.2 us on a 512 x 4096 slab).
static inline bool adseis_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ADSEIS_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}
static inline bool adseis_pdl_slab() {  // ADSEIS_PDL_SLAB=1: tuning experiments (PDL on slab plans too)
  static int on = -1;
  if (on < 0) on = getenv("ADSEIS_PDL_SLAB") != nullptr ? 1 : 0;
  return on != 0;
}
// ADSEIS_PDL_TBSLAB=1: PDL for the two-step path on slab plans too (experiments).  Measured on 2 B200s, C4: 107.7
// Gcell-upd/s with, 118.4 without -- CTAs of the next launch that become resident early spin on the neighbour's flags.
static inline int adseis_pdl_tb_slab() {  // 0 off, 1 frame-only launches, 2 box-pair launches too
  static int on = -1;
  if (on < 0) { const char* e = getenv("ADSEIS_PDL_TBSLAB"); on = (e && e[0] >= '1' && e[0] <= '2') ? e[0] - '0' : 0; }
  return on;
}
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_step(bool pdl, void (*kernel)(KArgs...), int grid, int block, size_t smem,
                                      cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (pdl && adseis_pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------------------
// Point sets: sources and receivers grouped by owning CTA and grid cell, so that the time-step kernels can
// inject / sample them in their CTA epilogue without atomics and in the reference's sequential order
// (ScatterAddOps.h:3-6, AddSource.cpp:57-85 add in source order; duplicates on one cell are legal).
//   blk[b]..blk[b+1]       = unique-cell entries owned by CTA b
//   cell[k]                = local flat offset row*ld + col of entry k   (elastic: field id in `field[k]`)
//   start[k]..start[k+1]   = range into perm[] of the points that sit on that cell, in original order
//   perm[m]                = original (global) point index
// ---------------------------------------------------------------------------------------------------------
struct PointSetHost {
  std::vector<int> blk, cell, start, perm, field;
};

// owner/cells/gid/field describe the kept points (field may be empty); nblocks = CTAs of the launch.
static inline void build_point_set(const std::vector<int>& owner, const std::vector<int>& cells,
                                   const std::vector<int>& gid, const std::vector<int>& field, int nblocks,
                                   PointSetHost* out) {
  const size_t n = owner.size();
  std::vector<size_t> order(n);
  for (size_t k = 0; k < n; k++) order[k] = k;
  auto less = [&](size_t a, size_t b) {
    if (owner[a] != owner[b]) return owner[a] < owner[b];
    if (cells[a] != cells[b]) return cells[a] < cells[b];
    const int fa = field.empty() ? 0 : field[a], fb = field.empty() ? 0 : field[b];
    return fa < fb;
  };
  std::stable_sort(order.begin(), order.end(), less);
  out->blk.assign((size_t)nblocks + 1, 0);
  out->cell.clear(); out->start.clear(); out->perm.clear(); out->field.clear();
  std::vector<int> ent_owner;
  for (size_t m = 0; m < n; m++) {
    const size_t k = order[m];
    const int f = field.empty() ? 0 : field[k];
    if (out->cell.empty() || ent_owner.back() != owner[k] || out->cell.back() != cells[k] || out->field.back() != f) {
      out->cell.push_back(cells[k]);
      out->field.push_back(f);
      out->start.push_back((int)m);
      ent_owner.push_back(owner[k]);
    }
    out->perm.push_back(gid[k]);
  }
  out->start.push_back((int)n);
  for (int o : ent_owner) out->blk[(size_t)o + 1]++;
  for (int b = 0; b < nblocks; b++) out->blk[(size_t)b + 1] += out->blk[(size_t)b];
}

struct PointSetStorage {
  int *blk = nullptr, *cell = nullptr, *start = nullptr, *perm = nullptr, *field = nullptr;
  int nu = 0, npts = 0;
  size_t cap[5] = {0, 0, 0, 0, 0};  // allocated elements: a re-upload of the same size keeps the device pointers
};
static inline int upload_ints_reuse(int** p, size_t* cap, const std::vector<int>& h, cudaStream_t s) {
  const size_t n = h.empty() ? 1 : h.size();
  if (*p == nullptr || *cap < n) {
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    TRY(dev_alloc(p, n));
    *cap = n;
  }
  if (!h.empty()) CUDA_TRY(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  return ADSEIS_OK;
}
// (re)upload: device buffers are reused when they are large enough, so a plan that is re-pointed at the next shot
// (same counts) keeps every device address -- kernel arguments, and a captured CUDA graph, stay valid
static inline int upload_point_set(const PointSetHost& h, PointSetStorage* st, cudaStream_t s) {
  TRY(upload_ints_reuse(&st->blk, &st->cap[0], h.blk, s));
  TRY(upload_ints_reuse(&st->cell, &st->cap[1], h.cell, s));
  TRY(upload_ints_reuse(&st->start, &st->cap[2], h.start, s));
  TRY(upload_ints_reuse(&st->perm, &st->cap[3], h.perm, s));
  TRY(upload_ints_reuse(&st->field, &st->cap[4], h.field, s));
  st->nu = (int)h.cell.size();
  st->npts = (int)h.perm.size();
  return ADSEIS_OK;
}
static inline void free_point_set(PointSetStorage* st) {
  cudaFree(st->blk); cudaFree(st->cell); cudaFree(st->start); cudaFree(st->perm); cudaFree(st->field);
  *st = PointSetStorage();
}

#ifdef __CUDACC__
// Packed halo words ("LL"): an 8-byte word = 32 data bits + the 32-bit epoch of the launch that sent it; two words per
// double.  8-byte stores are single-copy atomic, so a word whose epoch matches carries valid data -- no fence, no flag.
__device__ __forceinline__ void ll_put(ulonglong2* row, int j, double v, unsigned ep) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), e = (unsigned long long)ep << 32;
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(row + j), "l"(e | (b & 0xffffffffULL)), "l"(e | (b >> 32))
               : "memory");
}
__device__ __forceinline__ double ll_get(const ulonglong2* row, int j, unsigned ep, unsigned long long* my_flags) {
  unsigned long long x, y, spins = 0;
  while (true) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(row + j) : "memory");
    if ((unsigned)(x >> 32) == ep && (unsigned)(y >> 32) == ep) break;
    if (++spins > (1ULL << 27)) { my_flags[2] = 1ULL; break; }  // neighbour lost (minutes): report, do not hang
  }
  return __longlong_as_double((long long)((x & 0xffffffffULL) | (y << 32)));
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
// streaming (evict-first) 16-byte accesses for data touched once per time step
__device__ __forceinline__ double2 ld2_stream(const double* p) { return __ldcs(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st2_stream(double* p, double2 v) { __stcs(reinterpret_cast<double2*>(p), v); }

// RN(x / h) from rh = RN(1/h) computed once on the host: q0 = x*rh is within 2 ulp, one FMA correction makes it a
// faithful quotient and a second one the correctly rounded quotient (Markstein's theorem; the remainders
// r = x - q*h are exact in an FMA).  Five issue slots instead of the ~40-instruction divide sequence, whose slow
// path is also taken for every zero or denormal-range wavefield value.
__device__ __forceinline__ double div_markstein(double x, double h, double rh) {
  double q = x * rh;
  double r = fma(-q, h, x);
  q = fma(r, rh, q);
  r = fma(-q, h, x);
  return fma(r, rh, q);
}

// Numerators near the denormal range (|x| < 1e-280: the numerical precursor that runs ahead of every wavefront is a
// band tens of cells wide of such values) would make the remainders inexact.  They are scaled by 2^400 (exact), divided
// in the normal range, and scaled back.  Scaling back is exact while the quotient stays normal.  For a subnormal
// result the scaled quotient `q` = RN(2^400 x / h) is adjusted with the sign of its exact remainder before the final
// multiplication rounds it onto the coarser subnormal grid: made ROUND-TO-ODD when >= 2 bits are lost (then the second
// rounding cannot double-round), and stepped off the rounding boundary towards the true quotient when exactly one bit
// is lost.  Bit-identical to x / h for every input (host model of this routine checked against x / h on 4.4e8 random
// tiny numerators, 11 divisors: scripts/div_tiny_check.c); no divide instruction, ~15 inline instructions that only
// warps holding such a numerator execute.  Round 1 fell back to the divide sequence for these numerators, whose slow
// path made the CTAs that cross the precursor band several times slower than the rest (a long tail on every launch:
// elastic forward 117 -> 281 us per step at 2000^2 once the band had spread, profiles/r02c_elastic_division.md).
__device__ __forceinline__ double div_fix_tiny(double x, double h, double q) {
  const double aq = fabs(q);
  if (aq < 0x1p-622) {                                // subnormal result
    const double r = fma(-q, h, x * 0x1p+400);        // exact remainder: its sign tells on which side x/h lies
    long long b = __double_as_longlong(q);
    const bool one_bit = aq >= 0x1p-623;
    if (r != 0.0 && (((b & 1LL) == 0) != one_bit)) {
      const bool up = (r > 0.0) == (h > 0.0);         // true quotient > q ?
      b += ((q > 0.0) == up) ? 1LL : -1LL;
      q = __longlong_as_double(b);
    }
  }
  return q * 0x1p-400;
}

__device__ __forceinline__ double div_exact(double x, double h, double rh) {
  const bool tiny = (x != 0.0) & (fabs(x) < 1e-280);
  const double q = div_markstein(tiny ? x * 0x1p+400 : x, h, rh);
  return tiny ? div_fix_tiny(x, h, q) : q;
}

// Same routine for hot loops (the `tiny` flag of round 1, which made the caller redo the iteration with true
// divisions, is kept in the signature but no longer raised).
__device__ __forceinline__ double div_core(double x, double h, double rh, bool& tiny) {
  (void)tiny;
  return div_exact(x, h, rh);
}

// ---------------------------------------------------------------------------------------------------------
// TMA row staging (sm_90+/sm_100a): 1-D bulk asynchronous copies global -> shared (`cp.async.bulk`, SASS UBLKCP)
// completing on an mbarrier.  The marching kernels keep a ring of row stages per CTA: a producer warp arms the
// stage's `full` mbarrier with the byte count and issues one multi-KB bulk copy per streamed array several rows
// ahead; the consumer warps wait on the barrier's phase parity, read the stage and release it through an `empty`
// mbarrier.  Memory-level parallelism is then set by the ring depth (shared memory), not by registers/occupancy.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// Cross-proxy write-after-read fence of the TMA rings.  The consumers read a ring stage with ordinary (generic-proxy)
// shared-memory loads and release it through the `empty` mbarrier; the producer then refills the stage with
// cp.async.bulk, which writes through the ASYNC proxy.  mbarrier release/acquire orders generic-proxy accesses only,
// so the refill needs a proxy fence between the producer's acquire of `empty` and the bulk copies -- without it a
// refill can overtake a consumer's still-pending loads of the previous row.  Measured on B200 (round 2,
// profiles/r02_ring_race.md): without the fence the forward sweep at 2 CTAs/SM is non-deterministic in ~10 % of
// 120-step runs at 4096^2 (ring depth 8, and depth 4 padded to the same shared-memory footprint); with it, clean.
// -DADSEIS_NO_RING_FENCE rebuilds the unfenced round-1 behaviour for that A/B experiment.
__device__ __forceinline__ void ring_refill_fence() {
#ifndef ADSEIS_NO_RING_FENCE
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
// bytes: multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif
