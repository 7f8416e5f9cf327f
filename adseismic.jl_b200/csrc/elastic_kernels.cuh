// elastic_kernels.cuh -- fused elastic (velocity-stress, 4th-order staggered, CPML) time-step kernels, sm_100a.
//
// The reference runs one elastic step as ~60 gather/scatter custom-op launches plus TF element-wise kernels
// (src/Core.jl:96-228) and differentiates through them with tf.gradients.  Here a step is TWO fused launches
//   forward : el_sigma_fwd  = fw1 + fw2   (sigma_xx, sigma_yy, sigma_xy += dt * C : grad v   + CPML mem1..4)
//             el_vel_fwd    = fw3 + fw4   (vx, vy += dt / rho * div sigma                    + CPML mem5..8)
//                             + source injection (AddSource.cpp:57-85) + receiver sampling (GetReceive.cpp:10-46)
//   adjoint : el_vel_adj    = fw4^T + fw3^T  (sigma_bar += D^T dbar5..8, rho gradients, mem_bar5..8)
//             el_sigma_adj  = fw2^T + fw1^T  (v_bar += D^T dbar1..4, lambda/mu gradients, mem_bar1..4)
//                             + receiver-residual injection + grad_srcv sampling
// (the transposes are in GATHER form so that every output cell is written by one thread.)
//
// Both variants of the reference share the kernels through ElGeom: "S" = src/Core.jl (padded (NX+2)x(NY+2),
// averaged materials, a different update region per sub-step), "M" = src/MPIElastic.jl on the global grid
// ((NX+4)x(NY+4) with zero ghost cells, no averaging, every cell updated).
//
// Work decomposition (HBM-bound fp64 stencils, no tensor cores).  The grid of one launch holds two kinds of CTAs,
// described by a host-built table (ElCta):
//   * MARCHING CTAs tile the BOX: the cells that are inside the update region of all four sub-steps, carry no CPML
//     profile, and whose stencil neighbours (2 cells each way) all share those properties.  A CTA owns up to 512
//     columns and marches down its rows.  A producer warp streams one row of every input plane per iteration into
//     shared-memory rings with TMA bulk copies (one multi-KB `cp.async.bulk` per plane and row, mbarrier
//     full/empty handshake, EL_PF row bundles in flight); eight consumer warps (64 columns each, one double2 per
//     lane) keep the x-direction stencil windows in registers, take the y-direction neighbours of the centre row
//     from the ring, and store whole 128-byte lines.  Every plane element is requested from DRAM once per pass.
//   * GENERIC CTAs (64x16, 32x32 or 16x64 cell tiles, four cells per thread) cover the rest -- CPML strips, region edges,
//     ring / ghost cells -- with the reference's full expressions and the compact CPML memories.
// Sources / receivers are injected / sampled by the CTA that owns their cell, through per-CTA lists (no atomics;
// duplicates accumulate in the reference's order).
//
// State layout.  A wavefield SLOT is 5 pitched planes (vx, vy, sxx, syy, sxy) + the 8 CPML memories stored
// COMPACTLY: x-memories only on the rows where the x-profile is non-zero (all columns), y-memories only on the
// columns where the y-profile is non-zero (all rows) -- ~4 % of a slot at 2000^2.  Stress sources are injected
// LAZILY: a slot holds v AFTER and sigma BEFORE the injection of its step, and el_sigma_fwd adds the pending stress
// sources when it loads sigma (same `+=` arithmetic and order as the reference).  That keeps in the history exactly
// the quantities the adjoint needs (fw3/fw4 see pre-injection stresses, fw1/fw2 post-injection velocities) and
// avoids a separate injection launch racing with stencil reads.
// Forward arithmetic follows the reference's evaluation order (bit-identical with -fmad=false; the divisions by
// 24*dx, 24*dy of the marching CTAs use div_exact, which returns the correctly rounded quotient).
#pragma once
#include <type_traits>

#include "common.cuh"

#define EL_BX 64                    // generic tile: columns
#define EL_BY 4                     // generic tile: thread rows
#define EL_ROWS 16                  // generic tile: rows
#define EL_MW 8                     // marching CTA: consumer warps (32 columns each, one column per lane)
#define EL_NT (EL_MW * 32 + 32)     // threads per CTA: consumers + one producer warp (generic tiles use the first 256)
#define EL_TCOLS (EL_MW * 32)       // marching CTA: columns
// ring row stride: tile columns + a 2-column (16-byte) halo on each side, padded to a multiple of 128 bytes so that
// bulk copies into neighbouring ring rows never share a 128-byte shared-memory line (see AC_HPAD)
#ifdef ADSEIS_NO_SMEM_PAD
#define EL_RC (EL_TCOLS + 4)
#else
#define EL_RC ((EL_TCOLS + 4 + 15) / 16 * 16)
#endif
#ifndef EL_PF
#define EL_PF 2                     // row bundles in flight beyond the one being consumed
#endif
#define EL_NB (EL_PF + 1)           // bundle barriers
#ifndef EL_LATE_RELEASE
#define EL_LATE_RELEASE 1           // adjoint marching loops: release a ring bundle after the row's arithmetic (the loads need not
#endif                              // all be live at once: no spills at 96 registers) instead of right after its loads
#define EL_HALO 2                   // slab decomposition: halo rows per interior side
#ifndef EL_MINB_FWD
#define EL_MINB_FWD 3                // forward kernels: CTAs per SM the register allocation aims at
#endif

struct ElGeom {
  int H, W;       // global array rows / columns (incl. ring or ghost cells)
  int Hl, ld;     // local rows (incl. 2 halo rows per interior side) and pitch
  int goff;       // global row of local row 0
  int NX, NY;
  int p0[4], p1[4], q0[4], q1[4];  // inclusive GLOBAL update regions of fw1..fw4
  int cx, cy;     // CPML coefficient index = p - cx, q - cy
  int xlo, xhi, ylo, yhi;  // coefficient indices k < lo or k >= hi carry a non-zero CPML profile
  int nxr, ycp;   // compact x-memory rows (= xlo + NX - xhi), compact y-memory pitch
  double dt, dx, dy;
  double h24x, h24y, r24x, r24y;  // 24*dx, 24*dy and their correctly rounded reciprocals (host)
  i64 plane;      // Hl*ld
  i64 xm_sz, ym_sz;  // doubles per compact x-/y-memory array
  int own0, own1; // owned local rows [own0, own1)
};

// one CTA of a launch (host-built): kind 0 = marching tile of the box, 1 = generic tile
struct ElCta { int kind, r0, r1, c0, c1, ltw, pad1, pad2; };  // ltw: log2 of a generic tile's thread columns (4..6)

struct ElSlot {     // one wavefield slot (or the adjoint state)
  double *vx, *vy, *sxx, *syy, *sxy;
  double* xm;       // 4 compact x-memories: mem1 (x half), mem3 (x int), mem5 (x int), mem7 (x half)
  double* ym;       // 4 compact y-memories: mem2 (y int), mem4 (y half), mem6 (y int), mem8 (y half)
};

struct ElMat {      // materials as the kernels consume them (pitched planes)
  const double *lamb, *lmb;  // fw1: lambda_bar, lambda_bar + 2 mu_bar        (M: lambda, lambda + 2 mu)
  const double* mub2;        // fw2: mu_bar                                   (M: mu)
  const double *rho, *rhob;  // fw3: rho ; fw4: rho_bar                       (M: rho, rho)
  const double *rinv, *rbinv;  // reciprocals (adjoint only)
};

struct ElCoef { const double *ax, *bx, *ay, *by; };  // [2*NX], [2*NY]: row 0 integer grid, row 1 half grid

// point lists (sources / receivers) per CTA; see PointSet in common.cuh.  `field` 0..4 = vx,vy,sxx,syy,sxy.
struct ElPoints {
  const int *blk, *cell, *field, *start, *perm;
  const int *xstart, *xperm;  // per unique entry: points of the OTHER set on the same (cell, field)
};

// Slab decomposition (rows are split over GPUs, halo = 2 rows per interior side), fused into the step kernels; all
// null / zero on a single GPU.  The CTAs that hold one of my first / last EL_HALO owned rows next to a neighbour are
// launched first (perm), wait until the neighbour's PREVIOUS launch has delivered the halo rows they read, and --
// after their epilogue -- store their columns of those edge rows of every field this launch produces straight into
// the neighbour's halo rows over NVLink, then publish with a system-scope fence + atomic on the neighbour's flag.
// Replaces the 18 mpi_halo_exchange2 ops per step of src/MPIElastic.jl:411-436, 515-516, 551, 582, 618.
struct ElFuse {
  const int* perm;              // launch order -> logical CTA id, or null
  int has_lo, has_hi;
  int nf;                       // fields pushed by this launch (0..3)
  const double* src[3];         // my planes
  double* lo[3];                // rank-1: its upper halo rows (local rows Hl'-2, Hl'-1) of the same plane
  double* hi[3];                // rank+1: its lower halo rows (local rows 0, 1)
  unsigned long long *sig_lo, *sig_hi;   // neighbour flags to bump (peer pointers)
  unsigned long long* my_flags;          // [3] bumped by rank-1's step kernels, [4] by rank+1's, [2] error
  unsigned long long expect_lo, expect_hi;
  // packed halo rows (see common.cuh: ll_put / ll_get and csrc/acoustic_kernels.cuh AcFuse): the sender stores its
  // EL_HALO edge rows of the planes it produces into the neighbour's packed rows [field][row][column] and is done; the
  // receiving CTA of the neighbour's NEXT launch unpacks the columns it owns into the halo rows of the planes it reads
  // (= the planes the previous launch produced).
  int ll;
  unsigned ep_send, ep_recv;             // epoch stamped on my rows; epoch to wait for (0: nothing to receive)
  ulonglong2 *tx_lo, *tx_hi;             // the neighbours' packed rows (peer pointers)
  const ulonglong2 *rx_lo, *rx_hi;       // mine, filled by the neighbours' previous fused launch
  int nf_in;
  double* in[3];                         // my planes whose halo rows this launch reads
};

__device__ __forceinline__ int el_bid(const ElFuse& f) { return f.perm ? f.perm[blockIdx.x] : blockIdx.x; }

__device__ __forceinline__ ElCta el_cta(const ElCta* __restrict__ ctas, int bid) {
  const int4* p = reinterpret_cast<const int4*>(ctas + bid);
  const int4 a = p[0], b = p[1];
  ElCta d;
  d.kind = a.x; d.r0 = a.y; d.r1 = a.z; d.c0 = a.w; d.c1 = b.x; d.ltw = b.y; d.pad1 = d.pad2 = 0;
  return d;
}

// A CTA is an edge CTA when it holds one of my first / last EL_HALO owned rows next to a neighbour: those are the
// rows it must push, and (a superset of) the CTAs whose stencils read halo rows.
__device__ __forceinline__ void el_cta_edges(const ElGeom& g, const ElFuse& f, const ElCta& d, bool* t_lo, bool* t_hi) {
  *t_lo = f.has_lo && d.r0 < g.own0 + EL_HALO;
  *t_hi = f.has_hi && d.r1 > g.own1 - EL_HALO;
}

// out of line, plain values (keeps the step kernels' register budget and their parameter block out of local memory)
__device__ __noinline__ void el_ll_recv(const ulonglong2* rx_lo, const ulonglong2* rx_hi, double* in0, double* in1, double* in2,
                                        int nf_in, unsigned ep, unsigned long long* my_flags, int own0, int own1, int ld,
                                        int c0, int c1, bool t_lo, bool t_hi) {
  double* in[3] = {in0, in1, in2};
  for (int k = 0; k < nf_in; k++) {
#pragma unroll
    for (int r = 0; r < EL_HALO; r++) {
      const ulonglong2* lo = rx_lo + (i64)(k * EL_HALO + r) * ld;
      const ulonglong2* hi = rx_hi + (i64)(k * EL_HALO + r) * ld;
      for (int q = c0 + threadIdx.x; q < c1; q += EL_NT) {
        if (t_lo) in[k][(i64)(own0 - EL_HALO + r) * ld + q] = ll_get(lo, q, ep, my_flags);
        if (t_hi) in[k][(i64)(own1 + r) * ld + q] = ll_get(hi, q, ep, my_flags);
      }
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");   // the halo rows are read by bulk copies of this CTA
}
__device__ __noinline__ void el_ll_send(ulonglong2* tx_lo, ulonglong2* tx_hi, const double* s0, const double* s1, const double* s2,
                                        int nf, unsigned ep, int own0, int own1, int ld, int r0, int r1, int c0, int c1,
                                        bool t_lo, bool t_hi) {
  const double* src[3] = {s0, s1, s2};
  for (int k = 0; k < nf; k++) {
#pragma unroll
    for (int r = 0; r < EL_HALO; r++) {
      const int lo_row = own0 + r, hi_row = own1 - EL_HALO + r;
      if (t_lo && lo_row >= r0 && lo_row < r1)
        for (int q = c0 + threadIdx.x; q < c1; q += EL_NT) ll_put(tx_lo + (i64)(k * EL_HALO + r) * ld, q, src[k][(i64)lo_row * ld + q], ep);
      if (t_hi && hi_row >= r0 && hi_row < r1)
        for (int q = c0 + threadIdx.x; q < c1; q += EL_NT) ll_put(tx_hi + (i64)(k * EL_HALO + r) * ld, q, src[k][(i64)hi_row * ld + q], ep);
    }
  }
}

__device__ __forceinline__ void el_fuse_wait(const ElGeom& g, const ElFuse& f, const ElCta& d, bool t_lo, bool t_hi) {
  pdl_wait();  // (programmatic dependent launch) everything the previous launch wrote is visible from here on
  if (!(t_lo || t_hi)) return;  // CTA-uniform
  if (f.ll) {
    if (f.ep_recv != 0u) {
      el_ll_recv(f.rx_lo, f.rx_hi, f.in[0], f.in[1], f.in[2], f.nf_in, f.ep_recv, f.my_flags, g.own0, g.own1, g.ld, d.c0, d.c1,
                 t_lo, t_hi);
      __syncthreads();
    }
    return;
  }
  if (threadIdx.x == 0) {
    volatile unsigned long long* fl = f.my_flags;
    unsigned long long spins = 0;
    while ((t_lo && fl[3] < f.expect_lo) || (t_hi && fl[4] < f.expect_hi)) {
      if (++spins > (1ULL << 26)) { fl[2] = 1ULL; break; }  // neighbour lost: report, do not hang
    }
    __threadfence_system();
  }
  __syncthreads();
}

// push this CTA's columns of my edge rows, then signal (called by all threads of an edge CTA)
__device__ __forceinline__ void el_fuse_push(const ElGeom& g, const ElFuse& f, const ElCta& d, bool t_lo, bool t_hi) {
  if (!(t_lo || t_hi)) return;
  __syncthreads();  // all cells (and point injections) of this CTA are written
  if (f.ll) {
    el_ll_send(f.tx_lo, f.tx_hi, f.src[0], f.src[1], f.src[2], f.nf, f.ep_send, g.own0, g.own1, g.ld, d.r0, d.r1, d.c0, d.c1, t_lo, t_hi);
    return;
  }
  for (int k = 0; k < f.nf; k++) {
#pragma unroll
    for (int r = 0; r < EL_HALO; r++) {
      const int lo_row = g.own0 + r, hi_row = g.own1 - EL_HALO + r;
      if (t_lo && lo_row >= d.r0 && lo_row < d.r1)
        for (int q = d.c0 + threadIdx.x; q < d.c1; q += EL_NT) f.lo[k][(i64)r * g.ld + q] = f.src[k][(i64)lo_row * g.ld + q];
      if (t_hi && hi_row >= d.r0 && hi_row < d.r1)
        for (int q = d.c0 + threadIdx.x; q < d.c1; q += EL_NT) f.hi[k][(i64)r * g.ld + q] = f.src[k][(i64)hi_row * g.ld + q];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (t_lo) atomicAdd_system(f.sig_lo, 1ULL);
    if (t_hi) atomicAdd_system(f.sig_hi, 1ULL);
  }
}

__device__ __forceinline__ bool el_in(const ElGeom& g, int k, int gp, int q) {
  return gp >= g.p0[k] && gp <= g.p1[k] && q >= g.q0[k] && q <= g.q1[k];
}
__device__ __forceinline__ bool el_xpml(const ElGeom& g, int kx) { return kx < g.xlo || kx >= g.xhi; }
__device__ __forceinline__ bool el_ypml(const ElGeom& g, int ky) { return ky < g.ylo || ky >= g.yhi; }
__device__ __forceinline__ i64 el_xidx(const ElGeom& g, int kx, int q) {
  return (i64)(kx < g.xlo ? kx : kx - g.xhi + g.xlo) * g.ld + q;
}
__device__ __forceinline__ i64 el_yidx(const ElGeom& g, int li, int ky) {
  return (i64)li * g.ycp + (ky < g.ylo ? ky : ky - g.yhi + g.ylo);
}

// Add to `v` the values of the points of `ps` owned by this CTA that sit on (cell, field), in original order.
__device__ __forceinline__ double el_apply_points(double v, const ElPoints& ps, int a, int b, int cell, int field,
                                                  const double* __restrict__ val) {
  for (int k = a; k < b; k++)
    if (ps.cell[k] == cell && ps.field[k] == field)
      for (int m = ps.start[k]; m < ps.start[k + 1]; m++) v += val[ps.perm[m]];
  return v;
}

__device__ __forceinline__ double* el_field(const ElSlot& s, int f) {
  switch (f) {
    case 0: return s.vx;
    case 1: return s.vy;
    case 2: return s.sxx;
    case 3: return s.syy;
    default: return s.sxy;
  }
}

// ------------------------------------------------------------------------------------------------------------
// TMA row rings of the marching CTAs.
// A kernel streams NS planes; plane s has a LEAD L(s): iteration `it` (row r0+it of the tile) needs rows up to
// r0+it+L of it.  Bundle `it` = {row r0+it+L(s) of every plane s}, armed on full[it % EL_NB]; the ring of a lead-L
// plane has L+1+EL_PF rows (rows it..it+L live, EL_PF bundles in flight), slot = (row - r0) % depth.  The consumers
// release bundle barrier it % EL_NB after iteration `it`, which frees exactly the slots that bundle it+EL_NB
// overwrites (row r0+it of every plane).  A prologue bundle brings rows r0 .. r0+L-1.
// Every copy has the same shape: columns [c0-2, c1+2) of one row, (c1-c0+4)*8 bytes.
// bars: [0] prologue, [1..EL_NB] full, [1+EL_NB..2*EL_NB] empty.
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int el_depth(int L) { return L + 1 + EL_PF; }
template <class T>
__host__ __device__ constexpr int el_ring_off(int s) {  // first ring row of plane s
  int o = 0;
  for (int k = 0; k < s; k++) o += el_depth(T::lead(k));
  return o;
}
template <class T>
__host__ __device__ constexpr int el_ring_bytes() { return el_ring_off<T>(T::NS) * EL_RC * 8 + 64; }

#define EL_RING(T, s, slot) (ring + (size_t)(el_ring_off<T>(s) + (slot)) * EL_RC)

template <class T>
__device__ __forceinline__ void el_bars_init(unsigned long long* bars) {
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
#pragma unroll
    for (int k = 0; k < EL_NB; k++) { mbar_init(bars + 1 + k, 1); mbar_init(bars + 1 + EL_NB + k, EL_MW); }
    mbar_init_fence();
  }
  __syncthreads();
}

template <class T>
__device__ __forceinline__ void el_produce(const double* const* sp, double* ring, unsigned long long* bars, int r0,
                                           int nrows, int ld, int c0, int c1) {
  const unsigned bytes = (unsigned)(c1 - c0 + 4) * 8u;
  int npro = 0;
#pragma unroll
  for (int s = 0; s < T::NS; s++) npro += T::lead(s);
  if (npro > 0) {
    mbar_arrive_expect_tx(bars, (unsigned)npro * bytes);
#pragma unroll
    for (int s = 0; s < T::NS; s++)
#pragma unroll
      for (int k = 0; k < T::lead(s); k++)
        bulk_g2s(EL_RING(T, s, k), sp[s] + (i64)(r0 + k) * ld + c0 - 2, bytes, bars);
  } else {
    mbar_arrive(bars);
  }
  int sl[3] = {0, 1 % el_depth(1), 2 % el_depth(2)};  // slot of row it+L for L = 0,1,2
  for (int it = 0; it < nrows; it++) {
    const int b = it % EL_NB;
    if (it >= EL_NB) { mbar_wait(bars + 1 + EL_NB + b, (unsigned)(it / EL_NB - 1) & 1u); ring_refill_fence(); }
    mbar_arrive_expect_tx(bars + 1 + b, (unsigned)T::NS * bytes);
#pragma unroll
    for (int s = 0; s < T::NS; s++) {
      const int L = T::lead(s);
      bulk_g2s(EL_RING(T, s, sl[L]), sp[s] + (i64)(r0 + it + L) * ld + c0 - 2, bytes, bars + 1 + b);
    }
#pragma unroll
    for (int L = 0; L < 3; L++) sl[L] = (sl[L] + 1 == el_depth(L)) ? 0 : sl[L] + 1;
  }
}

// consumer-side ring cursor: slots of rows it (centre) and it+L (newest) for the three leads
struct ElCursor {
  int c[3];  // slot of row `it` in a lead-L ring
  int n[3];  // slot of row `it+L`
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int L = 0; L < 3; L++) { c[L] = 0; n[L] = L % el_depth(L); }
  }
  __device__ __forceinline__ void next() {
#pragma unroll
    for (int L = 0; L < 3; L++) {
      c[L] = (c[L] + 1 == el_depth(L)) ? 0 : c[L] + 1;
      n[L] = (n[L] + 1 == el_depth(L)) ? 0 : n[L] + 1;
    }
  }
};

// Adjoint arithmetic (compared with the reference at 1e-10, not bit for bit) uses fused multiply-adds explicitly -- the
// library is compiled with -fmad=false for the forward kernels -- and the 4th-order stencil as 27 (a - b) + (c - d).
__device__ __forceinline__ double el_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double el_st4(double a, double b, double c, double d) { return __fma_rn(27.0, a - b, c - d); }

// x / r with the zero numerators of the quiet zone kept off the divide's slow path
// (a select, not a branch: the compiler if-converts `x == 0 ? x : x / r` and then calls the divide's slow-path
// subroutine for every zero numerator -- 35 % of all instructions of el_vel_fwd in the quiet zone; dividing 1.0
// instead keeps the inline fast path and the result is discarded)
__device__ __forceinline__ double el_div_var(double x, double r) {
  const bool z = (x == 0.0);
  const bool tiny = !z & (fabs(x) < 1e-280);       // precursor band: keep the divide on its inline fast path too
  double xs = z ? 1.0 : (tiny ? x * 0x1p+400 : x);
  asm volatile("" : "+d"(xs));  // opaque: otherwise the compiler folds the selects back into x / r
  double q = xs / r;
  if (tiny) q = div_fix_tiny(x, r, q);             // exact scaling back (see common.cuh)
  return z ? x : q;
}

// ------------------------------------------------------------------------------------------------------------
// forward sigma pass: fw1 + fw2   (src/Core.jl:96-155, src/MPIElastic.jl:483-555)
//   in : slot s-1 (v post-injection, sigma pre-injection, memories)      out: sigma and mem1..4 of slot s
//   `src`/`srcv_prev`: stress sources of step s-1 that are still pending (null for s == 1)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void el_sigma_fwd_cell(const ElGeom& g, int li, int q, const ElSlot& in, const ElSlot& out,
                                                  const ElMat& mt, const ElCoef& cf, const ElPoints& src, int sa, int sb,
                                                  const double* __restrict__ srcv_prev) {
  const double* __restrict__ vx = in.vx;
  const double* __restrict__ vy = in.vy;
  const int ld = g.ld, NX = g.NX, NY = g.NY;
  const double dt = g.dt;
  const int gp = g.goff + li;
  const i64 c = (i64)li * ld + q;
  double sxx = in.sxx[c], syy = in.syy[c], sxy = in.sxy[c];
  if (sb > sa) {  // pending stress injection of the previous step (AddSource.cpp:69-84)
    sxx = el_apply_points(sxx, src, sa, sb, (int)c, 2, srcv_prev);
    syy = el_apply_points(syy, src, sa, sb, (int)c, 3, srcv_prev);
    sxy = el_apply_points(sxy, src, sa, sb, (int)c, 4, srcv_prev);
  }
  const int kx = gp - g.cx, ky = q - g.cy;
  if (el_in(g, 0, gp, q)) {  // fw1
    const double l_ = mt.lamb[c], lm = mt.lmb[c];
    double d1 = div_exact(27 * vx[c + ld] - 27 * vx[c] - vx[c + 2 * ld] + vx[c - ld], g.h24x, g.r24x);
    double d2 = div_exact(27 * vy[c] - 27 * vy[c - 1] - vy[c + 1] + vy[c - 2], g.h24y, g.r24y);
    if (el_xpml(g, kx)) {
      const i64 m = el_xidx(g, kx, q);
      const double n1 = cf.bx[NX + kx] * in.xm[m] + cf.ax[NX + kx] * d1;
      out.xm[m] = n1;
      d1 = d1 + n1;
    }
    if (el_ypml(g, ky)) {
      const i64 m = el_yidx(g, li, ky);
      const double n2 = cf.by[ky] * in.ym[m] + cf.ay[ky] * d2;
      out.ym[m] = n2;
      d2 = d2 + n2;
    }
    sxx += (lm * d1 + l_ * d2) * dt;
    syy += (lm * d2 + l_ * d1) * dt;
  }
  if (el_in(g, 1, gp, q)) {  // fw2
    const double m_ = mt.mub2[c];
    double d3 = div_exact(27 * vy[c] - 27 * vy[c - ld] - vy[c + ld] + vy[c - 2 * ld], g.h24x, g.r24x);
    double d4 = div_exact(27 * vx[c + 1] - 27 * vx[c] - vx[c + 2] + vx[c - 1], g.h24y, g.r24y);
    if (el_xpml(g, kx)) {
      const i64 m = g.xm_sz + el_xidx(g, kx, q);
      const double n1 = cf.bx[kx] * in.xm[m] + cf.ax[kx] * d3;
      out.xm[m] = n1;
      d3 = d3 + n1;
    }
    if (el_ypml(g, ky)) {
      const i64 m = g.ym_sz + el_yidx(g, li, ky);
      const double n2 = cf.by[NY + ky] * in.ym[m] + cf.ay[NY + ky] * d4;
      out.ym[m] = n2;
      d4 = d4 + n2;
    }
    sxy += m_ * (d3 + d4) * dt;
  }
  out.sxx[c] = sxx; out.syy[c] = syy; out.sxy[c] = sxy;
}

struct ElSigFwdT {  // planes: vx (lead 2), vy (lead 1), sxx, syy, sxy, lamb, lmb, mub2
  static constexpr int NS = 8;
  __host__ __device__ static constexpr int lead(int s) { return s == 0 ? 2 : (s == 1 ? 1 : 0); }
};

__device__ __forceinline__ void el_sigma_fwd_march(const ElGeom& g, const ElCta& d, const ElSlot& in, const ElSlot& out,
                                                   const ElMat& mt, const ElPoints& src, int sa, int sb,
                                                   const double* __restrict__ srcv_prev, double* ring,
                                                   unsigned long long* bars) {
  typedef ElSigFwdT T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = g.ld, nrows = d.r1 - d.r0;
  el_bars_init<T>(bars);
  if (warp == EL_MW) {
    if (lane == 0) {
      const double* sp[T::NS] = {in.vx, in.vy, in.sxx, in.syy, in.sxy, mt.lamb, mt.lmb, mt.mub2};
      el_produce<T>(sp, ring, bars, d.r0, nrows, ld, d.c0, d.c1);
    }
    return;
  }
  const int so = 2 + warp * 32 + lane;  // this lane's column inside a ring row
  const int q = d.c0 + warp * 32 + lane;
  const bool act = q < d.c1;
  const double dt = g.dt, hx = g.h24x, hy = g.h24y, rx = g.r24x, ry = g.r24y;
  // x-direction windows (own column): vx rows li-1, li, li+1 (+ li+2 new), vy rows li-2, li-1, li (+ li+1 new)
  double vxm1 = 0.0, vxc, vxp1, vym2 = 0.0, vym1 = 0.0, vyc;
  if (act) {
    vxm1 = in.vx[(i64)(d.r0 - 1) * ld + q];
    vym2 = in.vy[(i64)(d.r0 - 2) * ld + q];
    vym1 = in.vy[(i64)(d.r0 - 1) * ld + q];
  }
  mbar_wait(bars, 0);
  vxc = EL_RING(T, 0, 0)[so];
  vxp1 = EL_RING(T, 0, 1)[so];
  vyc = EL_RING(T, 1, 0)[so];
  ElCursor cu;
  cu.init();
  for (int it = 0; it < nrows; it++) {
    const int li = d.r0 + it, b = it % EL_NB;
    mbar_wait(bars + 1 + b, (unsigned)(it / EL_NB) & 1u);
    const double vxp2 = EL_RING(T, 0, cu.n[2])[so];
    const double vyp1 = EL_RING(T, 1, cu.n[1])[so];
    const double* cx_ = EL_RING(T, 0, cu.c[2]) + so;  // centre row of vx
    const double vx_m1 = cx_[-1], vx_p1 = cx_[1], vx_p2 = cx_[2];
    const double* cy_ = EL_RING(T, 1, cu.c[1]) + so;  // centre row of vy
    const double vy_m2 = cy_[-2], vy_m1 = cy_[-1], vy_p1 = cy_[1];
    double sxx = EL_RING(T, 2, cu.c[0])[so], syy = EL_RING(T, 3, cu.c[0])[so], sxy = EL_RING(T, 4, cu.c[0])[so];
    const double l_ = EL_RING(T, 5, cu.c[0])[so], lm = EL_RING(T, 6, cu.c[0])[so], m_ = EL_RING(T, 7, cu.c[0])[so];
    __syncwarp();  // every lane has read its ring rows
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + b);
    if (act) {
      const i64 c = (i64)li * ld + q;
      if (sb > sa) {  // pending stress injection of the previous step (AddSource.cpp:69-84)
        sxx = el_apply_points(sxx, src, sa, sb, (int)c, 2, srcv_prev);
        syy = el_apply_points(syy, src, sa, sb, (int)c, 3, srcv_prev);
        sxy = el_apply_points(sxy, src, sa, sb, (int)c, 4, srcv_prev);
      }
      // EX: true divisions (redo of a cell in which a numerator was near the denormal range); else exact
      // reciprocal-based quotients without branches
      auto calc = [&](auto EX) -> bool {
        constexpr bool E = decltype(EX)::value;
        bool tiny = false;
        auto DX = [&](double x) { return E ? x / hx : div_core(x, hx, rx, tiny); };
        auto DY = [&](double x) { return E ? x / hy : div_core(x, hy, ry, tiny); };
        const double d1 = DX(27 * vxp1 - 27 * vxc - vxp2 + vxm1);
        const double d2 = DY(27 * vyc - 27 * vy_m1 - vy_p1 + vy_m2);
        const double d3 = DX(27 * vyc - 27 * vym1 - vyp1 + vym2);
        const double d4 = DY(27 * vx_p1 - 27 * vxc - vx_p2 + vx_m1);
        out.sxx[c] = sxx + (lm * d1 + l_ * d2) * dt;
        out.syy[c] = syy + (lm * d2 + l_ * d1) * dt;
        out.sxy[c] = sxy + m_ * (d3 + d4) * dt;
        return tiny;
      };
      if (calc(std::false_type{})) calc(std::true_type{});
    }
    vxm1 = vxc; vxc = vxp1; vxp1 = vxp2;
    vym2 = vym1; vym1 = vyc; vyc = vyp1;
    cu.next();
  }
}

__global__ void __launch_bounds__(EL_NT, EL_MINB_FWD)
el_sigma_fwd(ElGeom g, const ElCta* __restrict__ ctas, ElSlot in, ElSlot out, ElMat mt, ElCoef cf, ElPoints src,
             const double* __restrict__ srcv_prev, ElFuse f) {
  extern __shared__ __align__(128) unsigned char el_smem[];
  pdl_launch_dependents();
  const int bid = el_bid(f);
  const ElCta d = el_cta(ctas, bid);
#ifdef EL_DEBUG_SKIP_GENERIC  // timing experiments only (wrong results)
  if (d.kind != 0) return;
#endif
#ifdef EL_DEBUG_SKIP_MARCH
  if (d.kind == 0) return;
#endif
  bool t_lo, t_hi;
  el_cta_edges(g, f, d, &t_lo, &t_hi);
  el_fuse_wait(g, f, d, t_lo, t_hi);
  int sa = 0, sb = 0;
  if (src.blk != nullptr && srcv_prev != nullptr) { sa = src.blk[bid]; sb = src.blk[bid + 1]; }
  if (d.kind == 0) {
    double* ring = reinterpret_cast<double*>(el_smem + 64);
    el_sigma_fwd_march(g, d, in, out, mt, src, sa, sb, srcv_prev, ring, reinterpret_cast<unsigned long long*>(el_smem));
  } else if (threadIdx.x < EL_BX * EL_BY) {
    const int q = d.c0 + (threadIdx.x & ((1 << d.ltw) - 1));
    if (q < d.c1)
      for (int li = d.r0 + (threadIdx.x >> d.ltw); li < d.r1; li += (EL_BX * EL_BY) >> d.ltw)
        el_sigma_fwd_cell(g, li, q, in, out, mt, cf, src, sa, sb, srcv_prev);
  }
  el_fuse_push(g, f, d, t_lo, t_hi);
}

// ------------------------------------------------------------------------------------------------------------
// forward velocity pass: fw3 + fw4 (src/Core.jl:158-213), then add_source (velocity types now, stress types lazily)
// and get_receive for this slot.
//   in : v of slot s-1, sigma (already updated) of slot s       out: v and mem5..8 of slot s
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void el_vel_fwd_cell(const ElGeom& g, int li, int q, const ElSlot& in, const ElSlot& out,
                                                const ElMat& mt, const ElCoef& cf) {
  const double* __restrict__ sxx = out.sxx;
  const double* __restrict__ syy = out.syy;
  const double* __restrict__ sxy = out.sxy;
  const int ld = g.ld, NX = g.NX, NY = g.NY;
  const double dt = g.dt;
  const int gp = g.goff + li;
  const i64 c = (i64)li * ld + q;
  double vx = in.vx[c], vy = in.vy[c];
  const int kx = gp - g.cx, ky = q - g.cy;
  if (el_in(g, 2, gp, q)) {  // fw3
    double d5 = div_exact(27 * sxx[c] - 27 * sxx[c - ld] - sxx[c + ld] + sxx[c - 2 * ld], g.h24x, g.r24x);
    double d6 = div_exact(27 * sxy[c] - 27 * sxy[c - 1] - sxy[c + 1] + sxy[c - 2], g.h24y, g.r24y);
    if (el_xpml(g, kx)) {
      const i64 m = 2 * g.xm_sz + el_xidx(g, kx, q);
      const double n1 = cf.bx[kx] * in.xm[m] + cf.ax[kx] * d5;
      out.xm[m] = n1;
      d5 = d5 + n1;
    }
    if (el_ypml(g, ky)) {
      const i64 m = 2 * g.ym_sz + el_yidx(g, li, ky);
      const double n2 = cf.by[ky] * in.ym[m] + cf.ay[ky] * d6;
      out.ym[m] = n2;
      d6 = d6 + n2;
    }
    vx += el_div_var((d5 + d6) * dt, mt.rho[c]);
  }
  if (el_in(g, 3, gp, q)) {  // fw4
    const double r_ = mt.rhob[c];
    double d7 = div_exact(27 * sxy[c + ld] - 27 * sxy[c] - sxy[c + 2 * ld] + sxy[c - ld], g.h24x, g.r24x);
    double d8 = div_exact(27 * syy[c + 1] - 27 * syy[c] - syy[c + 2] + syy[c - 1], g.h24y, g.r24y);
    if (el_xpml(g, kx)) {
      const i64 m = 3 * g.xm_sz + el_xidx(g, kx, q);
      const double n1 = cf.bx[NX + kx] * in.xm[m] + cf.ax[NX + kx] * d7;
      out.xm[m] = n1;
      d7 = d7 + n1;
    }
    if (el_ypml(g, ky)) {
      const i64 m = 3 * g.ym_sz + el_yidx(g, li, ky);
      const double n2 = cf.by[NY + ky] * in.ym[m] + cf.ay[NY + ky] * d8;
      out.ym[m] = n2;
      d8 = d8 + n2;
    }
    vy += el_div_var((d7 + d8) * dt, r_);
  }
  out.vx[c] = vx; out.vy[c] = vy;
}

struct ElVelFwdT {  // planes: sxy (lead 2), sxx (lead 1), syy, vx, vy, rho, rhob
  static constexpr int NS = 7;
  __host__ __device__ static constexpr int lead(int s) { return s == 0 ? 2 : (s == 1 ? 1 : 0); }
};

__device__ __forceinline__ void el_vel_fwd_march(const ElGeom& g, const ElCta& d, const ElSlot& in, const ElSlot& out,
                                                 const ElMat& mt, double* ring, unsigned long long* bars) {
  typedef ElVelFwdT T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = g.ld, nrows = d.r1 - d.r0;
  el_bars_init<T>(bars);
  if (warp == EL_MW) {
    if (lane == 0) {
      const double* sp[T::NS] = {out.sxy, out.sxx, out.syy, in.vx, in.vy, mt.rho, mt.rhob};
      el_produce<T>(sp, ring, bars, d.r0, nrows, ld, d.c0, d.c1);
    }
    return;
  }
  const int so = 2 + warp * 32 + lane;
  const int q = d.c0 + warp * 32 + lane;
  const bool act = q < d.c1;
  const double dt = g.dt, hx = g.h24x, hy = g.h24y, rx = g.r24x, ry = g.r24y;
  // windows: sxy rows li-1, li, li+1 (+ li+2 new) ; sxx rows li-2, li-1, li (+ li+1 new)
  double qm1 = 0.0, qc, qp1, xm2 = 0.0, xm1 = 0.0, xc;
  if (act) {
    qm1 = out.sxy[(i64)(d.r0 - 1) * ld + q];
    xm2 = out.sxx[(i64)(d.r0 - 2) * ld + q];
    xm1 = out.sxx[(i64)(d.r0 - 1) * ld + q];
  }
  mbar_wait(bars, 0);
  qc = EL_RING(T, 0, 0)[so];
  qp1 = EL_RING(T, 0, 1)[so];
  xc = EL_RING(T, 1, 0)[so];
  ElCursor cu;
  cu.init();
  for (int it = 0; it < nrows; it++) {
    const int li = d.r0 + it, b = it % EL_NB;
    mbar_wait(bars + 1 + b, (unsigned)(it / EL_NB) & 1u);
    const double qp2 = EL_RING(T, 0, cu.n[2])[so];
    const double xp1 = EL_RING(T, 1, cu.n[1])[so];
    const double* cq = EL_RING(T, 0, cu.c[2]) + so;  // centre row of sxy
    const double q_m2 = cq[-2], q_m1 = cq[-1], q_p1 = cq[1];
    const double* cy_ = EL_RING(T, 2, cu.c[0]) + so;  // syy
    const double y_m1 = cy_[-1], y_c = cy_[0], y_p1 = cy_[1], y_p2 = cy_[2];
    const double vx = EL_RING(T, 3, cu.c[0])[so], vy = EL_RING(T, 4, cu.c[0])[so];
    const double rh = EL_RING(T, 5, cu.c[0])[so], rb = EL_RING(T, 6, cu.c[0])[so];
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + b);
    if (act) {
      const i64 c = (i64)li * ld + q;
      auto calc = [&](auto EX) -> bool {  // see el_sigma_fwd_march
        constexpr bool E = decltype(EX)::value;
        bool tiny = false;
        auto DX = [&](double x) { return E ? x / hx : div_core(x, hx, rx, tiny); };
        auto DY = [&](double x) { return E ? x / hy : div_core(x, hy, ry, tiny); };
        const double d5 = DX(27 * xc - 27 * xm1 - xp1 + xm2);
        const double d6 = DY(27 * qc - 27 * q_m1 - q_p1 + q_m2);
        const double d7 = DX(27 * qp1 - 27 * qc - qp2 + qm1);
        const double d8 = DY(27 * y_p1 - 27 * y_c - y_p2 + y_m1);
        out.vx[c] = vx + el_div_var((d5 + d6) * dt, rh);
        out.vy[c] = vy + el_div_var((d7 + d8) * dt, rb);
        return tiny;
      };
      if (calc(std::false_type{})) calc(std::true_type{});
    }
    qm1 = qc; qc = qp1; qp1 = qp2;
    xm2 = xm1; xm1 = xc; xc = xp1;
    cu.next();
  }
}

__global__ void __launch_bounds__(EL_NT, EL_MINB_FWD)
el_vel_fwd(ElGeom g, const ElCta* __restrict__ ctas, ElSlot in, ElSlot out, ElMat mt, ElCoef cf, ElPoints src,
           const double* __restrict__ srcv_row, ElPoints rcv, double* __restrict__ rcvv, int rcv_stride, int slot,
           ElFuse f) {
  extern __shared__ __align__(128) unsigned char el_smem[];
  pdl_launch_dependents();
  const int bid = el_bid(f);
  const ElCta d = el_cta(ctas, bid);
#ifdef EL_DEBUG_SKIP_GENERIC  // timing experiments only (wrong results)
  if (d.kind != 0) return;
#endif
#ifdef EL_DEBUG_SKIP_MARCH
  if (d.kind == 0) return;
#endif
  bool t_lo, t_hi;
  el_cta_edges(g, f, d, &t_lo, &t_hi);
  el_fuse_wait(g, f, d, t_lo, t_hi);
  if (d.kind == 0) {
    double* ring = reinterpret_cast<double*>(el_smem + 64);
    el_vel_fwd_march(g, d, in, out, mt, ring, reinterpret_cast<unsigned long long*>(el_smem));
  } else if (threadIdx.x < EL_BX * EL_BY) {
    const int q = d.c0 + (threadIdx.x & ((1 << d.ltw) - 1));
    if (q < d.c1)
      for (int li = d.r0 + (threadIdx.x >> d.ltw); li < d.r1; li += (EL_BX * EL_BY) >> d.ltw) el_vel_fwd_cell(g, li, q, in, out, mt, cf);
  }
  // ---- epilogue: velocity sources of this step, then receivers of this slot ----
  int ia = 0, ib = 0, ra = 0, rb = 0;
  if (src.blk != nullptr && srcv_row != nullptr) { ia = src.blk[bid]; ib = src.blk[bid + 1]; }
  if (rcv.blk != nullptr && rcvv != nullptr) { ra = rcv.blk[bid]; rb = rcv.blk[bid + 1]; }
  if (ib != ia || rb != ra) {  // CTA-uniform
    __syncthreads();
    for (int k = ia + threadIdx.x; k < ib; k += EL_NT) {
      const int fd = src.field[k];
      if (fd <= 1) {
        double* fld = el_field(out, fd);
        double v = fld[src.cell[k]];
        for (int m = src.start[k]; m < src.start[k + 1]; m++) v += srcv_row[src.perm[m]];
        fld[src.cell[k]] = v;
      }
    }
    if (rb > ra) {
      __syncthreads();
      for (int k = ra + threadIdx.x; k < rb; k += EL_NT) {
        const int fd = rcv.field[k];
        double v = el_field(out, fd)[rcv.cell[k]];
        if (fd >= 2 && srcv_row != nullptr)  // post-injection value of a stress component
          for (int m = rcv.xstart[k]; m < rcv.xstart[k + 1]; m++) v += srcv_row[rcv.xperm[m]];
        for (int m = rcv.start[k]; m < rcv.start[k + 1]; m++) rcvv[(i64)rcv.perm[m] * rcv_stride + slot] = v;
      }
    }
  }
  el_fuse_push(g, f, d, t_lo, t_hi);
}

// ------------------------------------------------------------------------------------------------------------
// adjoint helpers: dbar_k(Q) = ebar + a * (mbar_in(Q) + ebar) for Q inside the region of its sub-step, else 0
// ------------------------------------------------------------------------------------------------------------
struct ElAdjCtx {
  const ElGeom& g;
  const ElSlot& b;      // adjoint fields + adjoint memories (input side)
  const ElMat& mt;
  const ElCoef& cf;
};

// fw3: ebar = dt * vxbar / rho ; x-memory #2 (mem5, x int), y-memory #2 (mem6, y int)
__device__ __forceinline__ double el_db5(const ElAdjCtx& A, int li, int q) {
  const ElGeom& g = A.g;
  const int gp = g.goff + li;
  if (li < 0 || li >= g.Hl || !el_in(g, 2, gp, q)) return 0.0;
  const i64 c = (i64)li * g.ld + q;
  const double eb = g.dt * A.b.vx[c] * A.mt.rinv[c];
  const int kx = gp - g.cx;
  if (!el_xpml(g, kx)) return eb;
  return eb + A.cf.ax[kx] * (A.b.xm[2 * g.xm_sz + el_xidx(g, kx, q)] + eb);
}
__device__ __forceinline__ double el_db6(const ElAdjCtx& A, int li, int q) {
  const ElGeom& g = A.g;
  const int gp = g.goff + li;
  if (q < 0 || q >= g.W || !el_in(g, 2, gp, q)) return 0.0;
  const i64 c = (i64)li * g.ld + q;
  const double eb = g.dt * A.b.vx[c] * A.mt.rinv[c];
  const int ky = q - g.cy;
  if (!el_ypml(g, ky)) return eb;
  return eb + A.cf.ay[ky] * (A.b.ym[2 * g.ym_sz + el_yidx(g, li, ky)] + eb);
}
// fw4: ebar = dt * vybar / rho_bar ; x-memory #3 (mem7, x half), y-memory #3 (mem8, y half)
__device__ __forceinline__ double el_db7(const ElAdjCtx& A, int li, int q) {
  const ElGeom& g = A.g;
  const int gp = g.goff + li;
  if (li < 0 || li >= g.Hl || !el_in(g, 3, gp, q)) return 0.0;
  const i64 c = (i64)li * g.ld + q;
  const double eb = g.dt * A.b.vy[c] * A.mt.rbinv[c];
  const int kx = gp - g.cx;
  if (!el_xpml(g, kx)) return eb;
  return eb + A.cf.ax[g.NX + kx] * (A.b.xm[3 * g.xm_sz + el_xidx(g, kx, q)] + eb);
}
__device__ __forceinline__ double el_db8(const ElAdjCtx& A, int li, int q) {
  const ElGeom& g = A.g;
  const int gp = g.goff + li;
  if (q < 0 || q >= g.W || !el_in(g, 3, gp, q)) return 0.0;
  const i64 c = (i64)li * g.ld + q;
  const double eb = g.dt * A.b.vy[c] * A.mt.rbinv[c];
  const int ky = q - g.cy;
  if (!el_ypml(g, ky)) return eb;
  return eb + A.cf.ay[g.NY + ky] * (A.b.ym[3 * g.ym_sz + el_yidx(g, li, ky)] + eb);
}
// fw1: ebar1 = lm*gx + l_*gy, ebar2 = lm*gy + l_*gx ; x-memory #0 (mem1, x half), y-memory #0 (mem2, y int)
__device__ __forceinline__ double el_db1(const ElAdjCtx& A, int li, int q) {
  const ElGeom& g = A.g;
  const int gp = g.goff + li;
  if (li < 0 || li >= g.Hl || !el_in(g, 0, gp, q)) return 0.0;
  const i64 c = (i64)li * g.ld + q;
  const double gx = g.dt * A.b.sxx[c], gy = g.dt * A.b.syy[c];
  const double eb = A.mt.lmb[c] * gx + A.mt.lamb[c] * gy;
  const int kx = gp - g.cx;
  if (!el_xpml(g, kx)) return eb;
  return eb + A.cf.ax[g.NX + kx] * (A.b.xm[el_xidx(g, kx, q)] + eb);
}
__device__ __forceinline__ double el_db2(const ElAdjCtx& A, int li, int q) {
  const ElGeom& g = A.g;
  const int gp = g.goff + li;
  if (q < 0 || q >= g.W || !el_in(g, 0, gp, q)) return 0.0;
  const i64 c = (i64)li * g.ld + q;
  const double gx = g.dt * A.b.sxx[c], gy = g.dt * A.b.syy[c];
  const double eb = A.mt.lmb[c] * gy + A.mt.lamb[c] * gx;
  const int ky = q - g.cy;
  if (!el_ypml(g, ky)) return eb;
  return eb + A.cf.ay[ky] * (A.b.ym[el_yidx(g, li, ky)] + eb);
}
// fw2: ebar = mu_bar * dt * sxybar ; x-memory #1 (mem3, x int), y-memory #1 (mem4, y half)
__device__ __forceinline__ double el_db3(const ElAdjCtx& A, int li, int q) {
  const ElGeom& g = A.g;
  const int gp = g.goff + li;
  if (li < 0 || li >= g.Hl || !el_in(g, 1, gp, q)) return 0.0;
  const i64 c = (i64)li * g.ld + q;
  const double eb = A.mt.mub2[c] * (g.dt * A.b.sxy[c]);
  const int kx = gp - g.cx;
  if (!el_xpml(g, kx)) return eb;
  return eb + A.cf.ax[kx] * (A.b.xm[g.xm_sz + el_xidx(g, kx, q)] + eb);
}
__device__ __forceinline__ double el_db4(const ElAdjCtx& A, int li, int q) {
  const ElGeom& g = A.g;
  const int gp = g.goff + li;
  if (q < 0 || q >= g.W || !el_in(g, 1, gp, q)) return 0.0;
  const i64 c = (i64)li * g.ld + q;
  const double eb = A.mt.mub2[c] * (g.dt * A.b.sxy[c]);
  const int ky = q - g.cy;
  if (!el_ypml(g, ky)) return eb;
  return eb + A.cf.ay[g.NY + ky] * (A.b.ym[g.ym_sz + el_yidx(g, li, ky)] + eb);
}

// ------------------------------------------------------------------------------------------------------------
// adjoint velocity pass: fw4^T + fw3^T.  Reads vbar (stencil) and mbar5..8 (input side), updates sigma_bar in
// place (own cell), writes mbar5..8 (output side), accumulates the rho gradients.
//   fwd : forward slot s (pre-injection stresses + new memories) -- only used when MATGRAD
//   rcv/res : stress-type receiver residuals of slot s, applied to sigma_bar on load
// ------------------------------------------------------------------------------------------------------------
template <bool MATGRAD>
__device__ __forceinline__ void el_vel_adj_cell(const ElGeom& g, int li, int q, const ElSlot& b, const ElSlot& bout,
                                                const ElSlot& fwd, const ElMat& mt, const ElCoef& cf,
                                                double* __restrict__ Gr3, double* __restrict__ Gr4, const ElPoints& rcv,
                                                int ra, int rb, const double* __restrict__ res, int res_stride, int slot) {
  const ElAdjCtx A{g, b, mt, cf};
  const int ld = g.ld, NX = g.NX, NY = g.NY;
  const double ix = 1.0 / (24 * g.dx), iy = 1.0 / (24 * g.dy);
  const int gp = g.goff + li;
  const i64 c = (i64)li * ld + q;
  double sxx = b.sxx[c], syy = b.syy[c], sxy = b.sxy[c];
  for (int k = ra; k < rb; k++) {  // stress-type receiver residuals of this slot (GetReceive.cpp:48-97)
    if (rcv.cell[k] == (int)c && rcv.field[k] >= 2) {
      double a = 0.0;
      for (int m = rcv.start[k]; m < rcv.start[k + 1]; m++) a += res[(i64)rcv.perm[m] * res_stride + slot];
      if (rcv.field[k] == 2) sxx += a; else if (rcv.field[k] == 3) syy += a; else sxy += a;
    }
  }
  const double d5c = el_db5(A, li, q), d7c = el_db7(A, li, q);
  // (D-x)^T dbar5 -> sxx ; (D-y)^T dbar6 -> sxy ; (D+x)^T dbar7 -> sxy ; (D+y)^T dbar8 -> syy
  sxx += (27 * d5c - 27 * el_db5(A, li + 1, q) - el_db5(A, li - 1, q) + el_db5(A, li + 2, q)) * ix;
  const double d6c = el_db6(A, li, q), d8c = el_db8(A, li, q);
  sxy += (27 * d6c - 27 * el_db6(A, li, q + 1) - el_db6(A, li, q - 1) + el_db6(A, li, q + 2)) * iy;
  sxy += (27 * el_db7(A, li - 1, q) - 27 * d7c - el_db7(A, li - 2, q) + el_db7(A, li + 1, q)) * ix;
  syy += (27 * el_db8(A, li, q - 1) - 27 * d8c - el_db8(A, li, q - 2) + el_db8(A, li, q + 1)) * iy;
  const int kx = gp - g.cx, ky = q - g.cy;
  const bool xp = el_xpml(g, kx), yp = el_ypml(g, ky);
  if (el_in(g, 2, gp, q)) {  // own-cell part of fw3^T
    const double gg = g.dt * b.vx[c];
    const double eb = gg * mt.rinv[c];
    double e56 = 0.0;
    if (MATGRAD) {
      const double* fs = fwd.sxx; const double* fq = fwd.sxy;
      e56 = div_exact(27 * fs[c] - 27 * fs[c - ld] - fs[c + ld] + fs[c - 2 * ld], g.h24x, g.r24x) +
            div_exact(27 * fq[c] - 27 * fq[c - 1] - fq[c + 1] + fq[c - 2], g.h24y, g.r24y);
    }
    if (xp) {
      const i64 m = 2 * g.xm_sz + el_xidx(g, kx, q);
      bout.xm[m] = cf.bx[kx] * (b.xm[m] + eb);
      if (MATGRAD) e56 += fwd.xm[m];
    }
    if (yp) {
      const i64 m = 2 * g.ym_sz + el_yidx(g, li, ky);
      bout.ym[m] = cf.by[ky] * (b.ym[m] + eb);
      if (MATGRAD) e56 += fwd.ym[m];
    }
    if (MATGRAD) Gr3[c] += -gg * e56 * (mt.rinv[c] * mt.rinv[c]);
  }
  if (el_in(g, 3, gp, q)) {  // own-cell part of fw4^T
    const double gg = g.dt * b.vy[c];
    const double eb = gg * mt.rbinv[c];
    double e78 = 0.0;
    if (MATGRAD) {
      const double* fq = fwd.sxy; const double* fy = fwd.syy;
      e78 = div_exact(27 * fq[c + ld] - 27 * fq[c] - fq[c + 2 * ld] + fq[c - ld], g.h24x, g.r24x) +
            div_exact(27 * fy[c + 1] - 27 * fy[c] - fy[c + 2] + fy[c - 1], g.h24y, g.r24y);
    }
    if (xp) {
      const i64 m = 3 * g.xm_sz + el_xidx(g, kx, q);
      bout.xm[m] = cf.bx[NX + kx] * (b.xm[m] + eb);
      if (MATGRAD) e78 += fwd.xm[m];
    }
    if (yp) {
      const i64 m = 3 * g.ym_sz + el_yidx(g, li, ky);
      bout.ym[m] = cf.by[NY + ky] * (b.ym[m] + eb);
      if (MATGRAD) e78 += fwd.ym[m];
    }
    if (MATGRAD) Gr4[c] += -gg * e78 * (mt.rbinv[c] * mt.rbinv[c]);
  }
  bout.sxx[c] = sxx; bout.syy[c] = syy; bout.sxy[c] = sxy;
}

// planes: vbx (2), rinv (2), vby (1), rbinv (1), sbxx, sbyy, sbxy (0) [+ fwd sxy (2), fwd sxx (1), fwd syy, Gr3, Gr4 (0)]
template <bool MATGRAD>
struct ElVelAdjT {
  static constexpr int NS = MATGRAD ? 12 : 7;
  __host__ __device__ static constexpr int lead(int s) {
    return (s == 0 || s == 1 || s == 7) ? 2 : ((s == 2 || s == 3 || s == 8) ? 1 : 0);
  }
};

template <bool MATGRAD>
__device__ __forceinline__ void el_vel_adj_march(const ElGeom& g, const ElCta& d, const ElSlot& b, const ElSlot& bout,
                                                 const ElSlot& fwd, const ElMat& mt, double* __restrict__ Gr3,
                                                 double* __restrict__ Gr4, const ElPoints& rcv, int ra, int rb,
                                                 const double* __restrict__ res, int res_stride, int slot, double* ring,
                                                 unsigned long long* bars) {
  typedef ElVelAdjT<MATGRAD> T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = g.ld, nrows = d.r1 - d.r0;
  el_bars_init<T>(bars);
  if (warp == EL_MW) {
    if (lane == 0) {
      const double* sp[12] = {b.vx, mt.rinv, b.vy, mt.rbinv, b.sxx, b.syy, b.sxy, fwd.sxy, fwd.sxx, fwd.syy, Gr3, Gr4};
      el_produce<T>(sp, ring, bars, d.r0, nrows, ld, d.c0, d.c1);
    }
    return;
  }
  const int so = 2 + warp * 32 + lane;
  const int q = d.c0 + warp * 32 + lane;
  const bool act = q < d.c1;
  const double dt = g.dt, dtix = dt * (1.0 / (24 * g.dx)), dtiy = dt * (1.0 / (24 * g.dy));
  const double rx = g.r24x, ry = g.r24y;  // gradient terms: quotients by 24*dx, 24*dy as products (1e-16 relative)
  // windows: A = vbx*rinv rows li-1, li, li+1 (+ li+2) ; B = vby*rbinv rows li-2, li-1, li (+ li+1); dt is folded
  // into the stencil weights dtix, dtiy
  double Am1 = 0.0, Ac, Ap1, Bm2 = 0.0, Bm1 = 0.0, Bc;
  double fqm1 = 0.0, fqc = 0.0, fqp1 = 0.0, fsm2 = 0.0, fsm1 = 0.0, fsc = 0.0;  // forward sxy / sxx windows (MATGRAD)
  if (act) {
    const i64 o1 = (i64)(d.r0 - 1) * ld + q, o2 = (i64)(d.r0 - 2) * ld + q;
    Am1 = b.vx[o1] * mt.rinv[o1];
    Bm2 = b.vy[o2] * mt.rbinv[o2];
    Bm1 = b.vy[o1] * mt.rbinv[o1];
    if (MATGRAD) { fqm1 = fwd.sxy[o1]; fsm2 = fwd.sxx[o2]; fsm1 = fwd.sxx[o1]; }
  }
  mbar_wait(bars, 0);
  Ac = EL_RING(T, 0, 0)[so] * EL_RING(T, 1, 0)[so];
  Ap1 = EL_RING(T, 0, 1)[so] * EL_RING(T, 1, 1)[so];
  Bc = EL_RING(T, 2, 0)[so] * EL_RING(T, 3, 0)[so];
  if (MATGRAD) { fqc = EL_RING(T, 7, 0)[so]; fqp1 = EL_RING(T, 7, 1)[so]; fsc = EL_RING(T, 8, 0)[so]; }
  ElCursor cu;
  cu.init();
  for (int it = 0; it < nrows; it++) {
    const int li = d.r0 + it, bb = it % EL_NB;
    mbar_wait(bars + 1 + bb, (unsigned)(it / EL_NB) & 1u);
    const double Ap2 = EL_RING(T, 0, cu.n[2])[so] * EL_RING(T, 1, cu.n[2])[so];
    const double Bp1 = EL_RING(T, 2, cu.n[1])[so] * EL_RING(T, 3, cu.n[1])[so];
    const double* va = EL_RING(T, 0, cu.c[2]) + so; const double* ria = EL_RING(T, 1, cu.c[2]) + so;  // centre row of A
    const double A_m1 = va[-1] * ria[-1], A_p1 = va[1] * ria[1], A_p2 = va[2] * ria[2];
    const double* vb = EL_RING(T, 2, cu.c[1]) + so; const double* rba = EL_RING(T, 3, cu.c[1]) + so;  // centre row of B
    const double B_m2 = vb[-2] * rba[-2], B_m1 = vb[-1] * rba[-1], B_p1 = vb[1] * rba[1];
    double sxx = EL_RING(T, 4, cu.c[0])[so], syy = EL_RING(T, 5, cu.c[0])[so], sxy = EL_RING(T, 6, cu.c[0])[so];
    double vxc = 0.0, ric = 0.0, vyc = 0.0, rbc = 0.0;  // raw centre values (MATGRAD)
    double fqp2 = 0.0, fsp1 = 0.0, fq_m2 = 0.0, fq_m1 = 0.0, fq_p1 = 0.0, fy_m1 = 0.0, fy_c = 0.0, fy_p1 = 0.0, fy_p2 = 0.0;
    double G3 = 0.0, G4 = 0.0;
    if (MATGRAD) {
      vxc = va[0]; ric = ria[0]; vyc = vb[0]; rbc = rba[0];
      fqp2 = EL_RING(T, 7, cu.n[2])[so];
      fsp1 = EL_RING(T, 8, cu.n[1])[so];
      const double* cq = EL_RING(T, 7, cu.c[2]) + so;  // centre row of forward sxy
      fq_m2 = cq[-2]; fq_m1 = cq[-1]; fq_p1 = cq[1];
      const double* cy_ = EL_RING(T, 9, cu.c[0]) + so;  // forward syy
      fy_m1 = cy_[-1]; fy_c = cy_[0]; fy_p1 = cy_[1]; fy_p2 = cy_[2];
      G3 = EL_RING(T, 10, cu.c[0])[so]; G4 = EL_RING(T, 11, cu.c[0])[so];
    }
#if !EL_LATE_RELEASE
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + bb);
#endif
    if (act) {
      const i64 c = (i64)li * ld + q;
      // (stress-type receiver residuals of this slot are added by the kernel after the march: addition commutes)
      // (D-x)^T A -> sxx ; (D-y)^T A -> sxy ; (D+x)^T B -> sxy ; (D+y)^T B -> syy
      sxx = el_fma(el_st4(Ac, Ap1, Ap2, Am1), dtix, sxx);
      sxy = el_fma(el_st4(Ac, A_p1, A_p2, A_m1), dtiy, sxy);
      sxy = el_fma(el_st4(Bm1, Bc, Bp1, Bm2), dtix, sxy);
      syy = el_fma(el_st4(B_m1, Bc, B_p1, B_m2), dtiy, syy);
      bout.sxx[c] = sxx; bout.syy[c] = syy; bout.sxy[c] = sxy;
      if (MATGRAD) {
        const double e56 = el_fma(el_st4(fsc, fsm1, fsm2, fsp1), rx, el_st4(fqc, fq_m1, fq_m2, fq_p1) * ry);
        const double e78 = el_fma(el_st4(fqp1, fqc, fqm1, fqp2), rx, el_st4(fy_p1, fy_c, fy_m1, fy_p2) * ry);
        Gr3[c] = el_fma(-(dt * vxc) * (ric * ric), e56, G3);
        Gr4[c] = el_fma(-(dt * vyc) * (rbc * rbc), e78, G4);
      }
    }
#if EL_LATE_RELEASE
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + bb);
#endif
    Am1 = Ac; Ac = Ap1; Ap1 = Ap2;
    Bm2 = Bm1; Bm1 = Bc; Bc = Bp1;
    if (MATGRAD) { fqm1 = fqc; fqc = fqp1; fqp1 = fqp2; fsm2 = fsm1; fsm1 = fsc; fsc = fsp1; }
    cu.next();
  }
}

template <bool MATGRAD>
// (9 warps per CTA: two CTAs per SM need <= 96 registers, three <= 72 -- the per-scheduler register files hold 5 / 7 warps)
__global__ void __launch_bounds__(EL_NT, MATGRAD ? 2 : 3)
el_vel_adj(ElGeom g, const ElCta* __restrict__ ctas, ElSlot b, ElSlot bout, ElSlot fwd, ElMat mt, ElCoef cf,
           double* __restrict__ Gr3, double* __restrict__ Gr4, ElPoints rcv, const double* __restrict__ res,
           int res_stride, int slot, ElFuse f) {
  extern __shared__ __align__(128) unsigned char el_smem[];
  pdl_launch_dependents();
  const int bid = el_bid(f);
  const ElCta d = el_cta(ctas, bid);
#ifdef EL_DEBUG_SKIP_GENERIC  // timing experiments only (wrong results)
  if (d.kind != 0) return;
#endif
#ifdef EL_DEBUG_SKIP_MARCH
  if (d.kind == 0) return;
#endif
  bool t_lo, t_hi;
  el_cta_edges(g, f, d, &t_lo, &t_hi);
  el_fuse_wait(g, f, d, t_lo, t_hi);
  int ra = 0, rb = 0;
  if (rcv.blk != nullptr && res != nullptr) { ra = rcv.blk[bid]; rb = rcv.blk[bid + 1]; }
  if (d.kind == 0) {
    double* ring = reinterpret_cast<double*>(el_smem + 64);
    el_vel_adj_march<MATGRAD>(g, d, b, bout, fwd, mt, Gr3, Gr4, rcv, ra, rb, res, res_stride, slot, ring,
                              reinterpret_cast<unsigned long long*>(el_smem));
    if (rb > ra) {  // CTA-uniform: stress-type receiver residuals of this slot (GetReceive.cpp:48-97) on my cells
      __syncthreads();
      for (int k = ra + threadIdx.x; k < rb; k += blockDim.x) {
        const int fd = rcv.field[k];
        if (fd >= 2) {
          double a = 0.0;
          for (int m = rcv.start[k]; m < rcv.start[k + 1]; m++) a += res[(i64)rcv.perm[m] * res_stride + slot];
          el_field(bout, fd)[rcv.cell[k]] += a;
        }
      }
    }
  } else if (threadIdx.x < EL_BX * EL_BY) {
    const int q = d.c0 + (threadIdx.x & ((1 << d.ltw) - 1));
    if (q < d.c1)
      for (int li = d.r0 + (threadIdx.x >> d.ltw); li < d.r1; li += (EL_BX * EL_BY) >> d.ltw)
        el_vel_adj_cell<MATGRAD>(g, li, q, b, bout, fwd, mt, cf, Gr3, Gr4, rcv, ra, rb, res, res_stride, slot);
  }
  el_fuse_push(g, f, d, t_lo, t_hi);
}

// Epilogue of the adjoint sigma pass (also launched on its own to start the reverse sweep at slot NSTEP):
// velocity-type receiver residuals of slot `slot_prev` are added into vbar, then grad_srcv is sampled --
// velocity types from vbar, stress types from sigma_bar plus the still pending stress residuals of that slot.
__device__ __forceinline__ void el_adj_epilogue(const ElGeom& g, const ElSlot& bout, const ElPoints& rcv,
                                                const double* __restrict__ res, int res_stride, int slot_prev,
                                                const ElPoints& src, double* __restrict__ gsrcv_row, int bid) {
  int ra = 0, rb = 0, sa = 0, sb = 0;
  if (rcv.blk != nullptr && res != nullptr) { ra = rcv.blk[bid]; rb = rcv.blk[bid + 1]; }
  if (src.blk != nullptr && gsrcv_row != nullptr) { sa = src.blk[bid]; sb = src.blk[bid + 1]; }
  if (rb == ra && sb == sa) return;
  __syncthreads();
  for (int k = ra + threadIdx.x; k < rb; k += blockDim.x) {
    const int f = rcv.field[k];
    if (f <= 1) {
      double* fld = el_field(bout, f);
      double v = fld[rcv.cell[k]];
      for (int m = rcv.start[k]; m < rcv.start[k + 1]; m++) v += res[(i64)rcv.perm[m] * res_stride + slot_prev];
      fld[rcv.cell[k]] = v;
    }
  }
  if (sb > sa) {
    __syncthreads();
    for (int k = sa + threadIdx.x; k < sb; k += blockDim.x) {
      const int f = src.field[k];
      double v = el_field(bout, f)[src.cell[k]];
      if (f >= 2 && res != nullptr)  // stress residuals of that slot are still pending on sigma_bar
        for (int m = src.xstart[k]; m < src.xstart[k + 1]; m++) v += res[(i64)src.xperm[m] * res_stride + slot_prev];
      for (int m = src.start[k]; m < src.start[k + 1]; m++) gsrcv_row[src.perm[m]] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
el_adj_start(ElGeom g, ElSlot bout, ElPoints rcv, const double* __restrict__ res, int res_stride, int slot_prev,
             ElPoints src, double* __restrict__ gsrcv_row) {
  el_adj_epilogue(g, bout, rcv, res, res_stride, slot_prev, src, gsrcv_row, blockIdx.x);
}

// ------------------------------------------------------------------------------------------------------------
// adjoint sigma pass: fw2^T + fw1^T.  Reads sigma_bar (stencil, already updated by el_vel_adj) and mbar1..4
// (input side), updates vbar in place (own cell), writes mbar1..4 (output side), accumulates lambda/mu gradients.
//   fwdv : forward slot s-1 (velocities, post-injection) ; fwdm : forward slot s (new memories 1..4)
// Epilogue: velocity-type receiver residuals of slot s-1 are injected into vbar; grad_srcv row (s-2) is sampled
// (AddSource.cpp:131-154): velocity types from vbar, stress types from sigma_bar + the pending stress residuals.
// ------------------------------------------------------------------------------------------------------------
template <bool MATGRAD>
__device__ __forceinline__ void el_sigma_adj_cell(const ElGeom& g, int li, int q, const ElSlot& b, const ElSlot& bout,
                                                  const ElSlot& fwdv, const ElSlot& fwdm, const ElMat& mt,
                                                  const ElCoef& cf, double* __restrict__ Gl, double* __restrict__ Gm1,
                                                  double* __restrict__ Gm2) {
  const ElAdjCtx A{g, b, mt, cf};
  const int ld = g.ld, NX = g.NX, NY = g.NY;
  const double ix = 1.0 / (24 * g.dx), iy = 1.0 / (24 * g.dy);
  const int gp = g.goff + li;
  const i64 c = (i64)li * ld + q;
  double vx = b.vx[c], vy = b.vy[c];
  const double d1c = el_db1(A, li, q), d3c = el_db3(A, li, q);
  const double d2c = el_db2(A, li, q), d4c = el_db4(A, li, q);
  // (D+x)^T dbar1 -> vx ; (D-y)^T dbar2 -> vy ; (D-x)^T dbar3 -> vy ; (D+y)^T dbar4 -> vx
  vx += (27 * el_db1(A, li - 1, q) - 27 * d1c - el_db1(A, li - 2, q) + el_db1(A, li + 1, q)) * ix;
  vy += (27 * d2c - 27 * el_db2(A, li, q + 1) - el_db2(A, li, q - 1) + el_db2(A, li, q + 2)) * iy;
  vy += (27 * d3c - 27 * el_db3(A, li + 1, q) - el_db3(A, li - 1, q) + el_db3(A, li + 2, q)) * ix;
  vx += (27 * el_db4(A, li, q - 1) - 27 * d4c - el_db4(A, li, q - 2) + el_db4(A, li, q + 1)) * iy;
  const int kx = gp - g.cx, ky = q - g.cy;
  const bool xp = el_xpml(g, kx), yp = el_ypml(g, ky);
  if (el_in(g, 0, gp, q)) {  // own-cell part of fw1^T
    const double gx = g.dt * b.sxx[c], gy = g.dt * b.syy[c];
    const double lm = mt.lmb[c], l_ = mt.lamb[c];
    const double eb1 = lm * gx + l_ * gy, eb2 = lm * gy + l_ * gx;
    double e1 = 0.0, e2 = 0.0;
    if (MATGRAD) {
      const double* fx = fwdv.vx; const double* fy = fwdv.vy;
      e1 = div_exact(27 * fx[c + ld] - 27 * fx[c] - fx[c + 2 * ld] + fx[c - ld], g.h24x, g.r24x);
      e2 = div_exact(27 * fy[c] - 27 * fy[c - 1] - fy[c + 1] + fy[c - 2], g.h24y, g.r24y);
    }
    if (xp) {
      const i64 m = el_xidx(g, kx, q);
      bout.xm[m] = cf.bx[NX + kx] * (b.xm[m] + eb1);
      if (MATGRAD) e1 += fwdm.xm[m];
    }
    if (yp) {
      const i64 m = el_yidx(g, li, ky);
      bout.ym[m] = cf.by[ky] * (b.ym[m] + eb2);
      if (MATGRAD) e2 += fwdm.ym[m];
    }
    if (MATGRAD) {
      Gl[c] += (gx + gy) * (e1 + e2);
      Gm1[c] += 2 * (gx * e1 + gy * e2);
    }
  }
  if (el_in(g, 1, gp, q)) {  // own-cell part of fw2^T
    const double gg = g.dt * b.sxy[c];
    const double eb = mt.mub2[c] * gg;
    double e34 = 0.0;
    if (MATGRAD) {
      const double* fx = fwdv.vx; const double* fy = fwdv.vy;
      e34 = div_exact(27 * fy[c] - 27 * fy[c - ld] - fy[c + ld] + fy[c - 2 * ld], g.h24x, g.r24x) +
            div_exact(27 * fx[c + 1] - 27 * fx[c] - fx[c + 2] + fx[c - 1], g.h24y, g.r24y);
    }
    if (xp) {
      const i64 m = g.xm_sz + el_xidx(g, kx, q);
      bout.xm[m] = cf.bx[kx] * (b.xm[m] + eb);
      if (MATGRAD) e34 += fwdm.xm[m];
    }
    if (yp) {
      const i64 m = g.ym_sz + el_yidx(g, li, ky);
      bout.ym[m] = cf.by[NY + ky] * (b.ym[m] + eb);
      if (MATGRAD) e34 += fwdm.ym[m];
    }
    if (MATGRAD) Gm2[c] += gg * e34;
  }
  bout.vx[c] = vx; bout.vy[c] = vy;
}

// planes: sbxy (2), mub2 (2), sbxx, sbyy, lmb, lamb (1), vbx, vby (0) [+ fwd vx (2), fwd vy (1), Gl, Gm1, Gm2 (0)]
template <bool MATGRAD>
struct ElSigAdjT {
  static constexpr int NS = MATGRAD ? 13 : 8;
  __host__ __device__ static constexpr int lead(int s) {
    return (s == 0 || s == 1 || s == 8) ? 2 : ((s >= 2 && s <= 5) || s == 9 ? 1 : 0);
  }
};

template <bool MATGRAD>
__device__ __forceinline__ void el_sigma_adj_march(const ElGeom& g, const ElCta& d, const ElSlot& b, const ElSlot& bout,
                                                   const ElSlot& fwdv, const ElMat& mt, double* __restrict__ Gl,
                                                   double* __restrict__ Gm1, double* __restrict__ Gm2, double* ring,
                                                   unsigned long long* bars) {
  typedef ElSigAdjT<MATGRAD> T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = g.ld, nrows = d.r1 - d.r0;
  el_bars_init<T>(bars);
  if (warp == EL_MW) {
    if (lane == 0) {
      const double* sp[13] = {b.sxy, mt.mub2, b.sxx, b.syy, mt.lmb, mt.lamb, b.vx, b.vy, fwdv.vx, fwdv.vy, Gl, Gm1, Gm2};
      el_produce<T>(sp, ring, bars, d.r0, nrows, ld, d.c0, d.c1);
    }
    return;
  }
  const int so = 2 + warp * 32 + lane;
  const int q = d.c0 + warp * 32 + lane;
  const bool act = q < d.c1;
  const double dt = g.dt, dtix = dt * (1.0 / (24 * g.dx)), dtiy = dt * (1.0 / (24 * g.dy));
  const double rx = g.r24x, ry = g.r24y;  // gradient terms: quotients by 24*dx, 24*dy as products (1e-16 relative)
  // e3 = mub2 * sbxy ; e1 = lmb*sbxx + lamb*sbyy ; e2 = lmb*sbyy + lamb*sbxx ; dt is folded into dtix, dtiy
  auto e3f = [](double s, double m) { return m * s; };
  auto e1f = [](double sx, double sy, double lm, double l_) { return el_fma(lm, sx, l_ * sy); };
  // windows: E3 rows li-1, li, li+1 (+ li+2) ; E1 rows li-2, li-1, li (+ li+1)
  double E3m1 = 0.0, E3c, E3p1, E1m2 = 0.0, E1m1 = 0.0, E1c;
  double fxm1 = 0.0, fxc = 0.0, fxp1 = 0.0, fym2 = 0.0, fym1 = 0.0, fyc = 0.0;  // forward vx / vy windows (MATGRAD)
  if (act) {
    const i64 o1 = (i64)(d.r0 - 1) * ld + q, o2 = (i64)(d.r0 - 2) * ld + q;
    E3m1 = e3f(b.sxy[o1], mt.mub2[o1]);
    E1m2 = e1f(b.sxx[o2], b.syy[o2], mt.lmb[o2], mt.lamb[o2]);
    E1m1 = e1f(b.sxx[o1], b.syy[o1], mt.lmb[o1], mt.lamb[o1]);
    if (MATGRAD) { fxm1 = fwdv.vx[o1]; fym2 = fwdv.vy[o2]; fym1 = fwdv.vy[o1]; }
  }
  mbar_wait(bars, 0);
  E3c = e3f(EL_RING(T, 0, 0)[so], EL_RING(T, 1, 0)[so]);
  E3p1 = e3f(EL_RING(T, 0, 1)[so], EL_RING(T, 1, 1)[so]);
  E1c = e1f(EL_RING(T, 2, 0)[so], EL_RING(T, 3, 0)[so], EL_RING(T, 4, 0)[so], EL_RING(T, 5, 0)[so]);
  if (MATGRAD) { fxc = EL_RING(T, 8, 0)[so]; fxp1 = EL_RING(T, 8, 1)[so]; fyc = EL_RING(T, 9, 0)[so]; }
  ElCursor cu;
  cu.init();
  for (int it = 0; it < nrows; it++) {
    const int li = d.r0 + it, bb = it % EL_NB;
    mbar_wait(bars + 1 + bb, (unsigned)(it / EL_NB) & 1u);
    const double E3p2 = e3f(EL_RING(T, 0, cu.n[2])[so], EL_RING(T, 1, cu.n[2])[so]);
    const double E1p1 = e1f(EL_RING(T, 2, cu.n[1])[so], EL_RING(T, 3, cu.n[1])[so], EL_RING(T, 4, cu.n[1])[so],
                            EL_RING(T, 5, cu.n[1])[so]);
    const double* sq = EL_RING(T, 0, cu.c[2]) + so; const double* mu = EL_RING(T, 1, cu.c[2]) + so;  // centre row of e3
    const double E3_m2 = e3f(sq[-2], mu[-2]), E3_m1 = e3f(sq[-1], mu[-1]), E3_p1 = e3f(sq[1], mu[1]);
    const double* sx = EL_RING(T, 2, cu.c[1]) + so; const double* sy = EL_RING(T, 3, cu.c[1]) + so;   // centre row of e2
    const double* lm = EL_RING(T, 4, cu.c[1]) + so; const double* l_ = EL_RING(T, 5, cu.c[1]) + so;
    // e2 = e1 with gx <-> gy
    const double E2_m1 = e1f(sy[-1], sx[-1], lm[-1], l_[-1]), E2c = e1f(sy[0], sx[0], lm[0], l_[0]);
    const double E2_p1 = e1f(sy[1], sx[1], lm[1], l_[1]), E2_p2 = e1f(sy[2], sx[2], lm[2], l_[2]);
    double vx = EL_RING(T, 6, cu.c[0])[so], vy = EL_RING(T, 7, cu.c[0])[so];
    double sxc = 0.0, syc = 0.0, sqc = 0.0;  // raw centre sigma_bar (MATGRAD)
    double fxp2 = 0.0, fyp1 = 0.0, fx_m1 = 0.0, fx_p1 = 0.0, fx_p2 = 0.0, fy_m2 = 0.0, fy_m1 = 0.0, fy_p1 = 0.0;
    double GL = 0.0, GM1 = 0.0, GM2 = 0.0;
    if (MATGRAD) {
      sxc = sx[0]; syc = sy[0]; sqc = sq[0];
      fxp2 = EL_RING(T, 8, cu.n[2])[so];
      fyp1 = EL_RING(T, 9, cu.n[1])[so];
      const double* cx_ = EL_RING(T, 8, cu.c[2]) + so;  // centre row of forward vx
      fx_m1 = cx_[-1]; fx_p1 = cx_[1]; fx_p2 = cx_[2];
      const double* cy_ = EL_RING(T, 9, cu.c[1]) + so;  // centre row of forward vy
      fy_m2 = cy_[-2]; fy_m1 = cy_[-1]; fy_p1 = cy_[1];
      GL = EL_RING(T, 10, cu.c[0])[so]; GM1 = EL_RING(T, 11, cu.c[0])[so]; GM2 = EL_RING(T, 12, cu.c[0])[so];
    }
#if !EL_LATE_RELEASE
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + bb);
#endif
    if (act) {
      const i64 c = (i64)li * ld + q;
      // (D+x)^T e1 -> vx ; (D-y)^T e2 -> vy ; (D-x)^T e3 -> vy ; (D+y)^T e3 -> vx
      vx = el_fma(el_st4(E1m1, E1c, E1p1, E1m2), dtix, vx);
      vy = el_fma(el_st4(E2c, E2_p1, E2_p2, E2_m1), dtiy, vy);
      vy = el_fma(el_st4(E3c, E3p1, E3p2, E3m1), dtix, vy);
      vx = el_fma(el_st4(E3_m1, E3c, E3_p1, E3_m2), dtiy, vx);
      bout.vx[c] = vx; bout.vy[c] = vy;
      if (MATGRAD) {
        const double gx = dt * sxc, gy = dt * syc, gg = dt * sqc;
        const double e1 = el_st4(fxp1, fxc, fxm1, fxp2) * rx;
        const double e2 = el_st4(fyc, fy_m1, fy_m2, fy_p1) * ry;
        const double e34 = el_fma(el_st4(fyc, fym1, fym2, fyp1), rx, el_st4(fx_p1, fxc, fx_m1, fx_p2) * ry);
        Gl[c] = el_fma(gx + gy, e1 + e2, GL);
        Gm1[c] = el_fma(2.0, el_fma(gx, e1, gy * e2), GM1);
        Gm2[c] = el_fma(gg, e34, GM2);
      }
    }
#if EL_LATE_RELEASE
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + bb);
#endif
    E3m1 = E3c; E3c = E3p1; E3p1 = E3p2;
    E1m2 = E1m1; E1m1 = E1c; E1c = E1p1;
    if (MATGRAD) { fxm1 = fxc; fxc = fxp1; fxp1 = fxp2; fym2 = fym1; fym1 = fyc; fyc = fyp1; }
    cu.next();
  }
}

template <bool MATGRAD>
__global__ void __launch_bounds__(EL_NT, MATGRAD ? 2 : 3)
el_sigma_adj(ElGeom g, const ElCta* __restrict__ ctas, ElSlot b, ElSlot bout, ElSlot fwdv, ElSlot fwdm, ElMat mt,
             ElCoef cf, double* __restrict__ Gl, double* __restrict__ Gm1, double* __restrict__ Gm2, ElPoints rcv,
             const double* __restrict__ res, int res_stride, int slot_prev, ElPoints src,
             double* __restrict__ gsrcv_row, ElFuse f) {
  extern __shared__ __align__(128) unsigned char el_smem[];
  pdl_launch_dependents();
  const int bid = el_bid(f);
  const ElCta d = el_cta(ctas, bid);
#ifdef EL_DEBUG_SKIP_GENERIC  // timing experiments only (wrong results)
  if (d.kind != 0) return;
#endif
#ifdef EL_DEBUG_SKIP_MARCH
  if (d.kind == 0) return;
#endif
  bool t_lo, t_hi;
  el_cta_edges(g, f, d, &t_lo, &t_hi);
  el_fuse_wait(g, f, d, t_lo, t_hi);
  if (d.kind == 0) {
    double* ring = reinterpret_cast<double*>(el_smem + 64);
    el_sigma_adj_march<MATGRAD>(g, d, b, bout, fwdv, mt, Gl, Gm1, Gm2, ring, reinterpret_cast<unsigned long long*>(el_smem));
  } else if (threadIdx.x < EL_BX * EL_BY) {
    const int q = d.c0 + (threadIdx.x & ((1 << d.ltw) - 1));
    if (q < d.c1)
      for (int li = d.r0 + (threadIdx.x >> d.ltw); li < d.r1; li += (EL_BX * EL_BY) >> d.ltw)
        el_sigma_adj_cell<MATGRAD>(g, li, q, b, bout, fwdv, fwdm, mt, cf, Gl, Gm1, Gm2);
  }
  el_adj_epilogue(g, bout, rcv, res, res_stride, slot_prev, src, gsrcv_row, bid);
  el_fuse_push(g, f, d, t_lo, t_hi);
}
