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
#include "common.cuh"

#define EL_BX 64                    // generic tile: columns
#define EL_BY 4                     // generic tile: thread rows
#define EL_ROWS 16                  // generic tile: rows
#define EL_MW 8                     // marching CTA: consumer warps (64 columns each)
#define EL_NT (EL_MW * 32 + 32)     // threads per CTA: consumers + one producer warp (generic tiles use the first 256)
#define EL_TCOLS (EL_MW * 64)       // marching CTA: columns
#define EL_RC (EL_TCOLS + 4)        // ring row: tile columns + a 2-column (16-byte) halo on each side
#ifndef EL_PF
#define EL_PF 2                     // row bundles in flight beyond the one being consumed
#endif
#define EL_NB (EL_PF + 1)           // bundle barriers
#define EL_HALO 2                   // slab decomposition: halo rows per interior side

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

__device__ __forceinline__ void el_fuse_wait(const ElFuse& f, bool t_lo, bool t_hi) {
  if (!(t_lo || t_hi)) return;  // CTA-uniform
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
    if (it >= EL_NB) mbar_wait(bars + 1 + EL_NB + b, (unsigned)(it / EL_NB - 1) & 1u);
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

__device__ __forceinline__ double2 mk2(double a, double b) { return make_double2(a, b); }
// x / r with the zero numerators of the quiet zone kept off the divide's slow path
__device__ __forceinline__ double el_div_var(double x, double r) { return x == 0.0 ? x : x / r; }

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
  const int so = 2 + warp * 64 + 2 * lane;  // this lane's column pair inside a ring row
  const int q = d.c0 + warp * 64 + 2 * lane;
  const bool act = q < d.c1;  // tile widths are even
  const bool wact = d.c0 + warp * 64 < d.c1;
  const double dt = g.dt, hx = g.h24x, hy = g.h24y, rx = g.r24x, ry = g.r24y;
  const double2 z2 = mk2(0.0, 0.0);
  // x-direction windows (own column pair): vx rows li-1, li, li+1 (+ li+2 new), vy rows li-2, li-1, li (+ li+1 new)
  double2 vxm1 = z2, vxc = z2, vxp1 = z2, vym2 = z2, vym1 = z2, vyc = z2;
  if (act) {
    vxm1 = ld2(in.vx + (i64)(d.r0 - 1) * ld + q);
    vym2 = ld2(in.vy + (i64)(d.r0 - 2) * ld + q);
    vym1 = ld2(in.vy + (i64)(d.r0 - 1) * ld + q);
  }
  mbar_wait(bars, 0);
  if (wact) {
    vxc = ld2(EL_RING(T, 0, 0) + so);
    vxp1 = ld2(EL_RING(T, 0, 1) + so);
    vyc = ld2(EL_RING(T, 1, 0) + so);
  }
  ElCursor cu;
  cu.init();
  for (int it = 0; it < nrows; it++) {
    const int li = d.r0 + it, b = it % EL_NB;
    mbar_wait(bars + 1 + b, (unsigned)(it / EL_NB) & 1u);
    double2 vxp2 = z2, vyp1 = z2, vxr = z2, vyl = z2, sxx = z2, syy = z2, sxy = z2, l_ = z2, lm = z2, m_ = z2;
    double vxl = 0.0, vyr = 0.0;
    if (wact) {
      vxp2 = ld2(EL_RING(T, 0, cu.n[2]) + so);
      vyp1 = ld2(EL_RING(T, 1, cu.n[1]) + so);
      const double* cx_ = EL_RING(T, 0, cu.c[2]);  // centre row of vx: columns q-1, q+2, q+3
      vxl = cx_[so - 1]; vxr = ld2(cx_ + so + 2);
      const double* cy_ = EL_RING(T, 1, cu.c[1]);  // centre row of vy: columns q-2, q-1, q+2
      vyl = ld2(cy_ + so - 2); vyr = cy_[so + 2];
      sxx = ld2(EL_RING(T, 2, cu.c[0]) + so); syy = ld2(EL_RING(T, 3, cu.c[0]) + so); sxy = ld2(EL_RING(T, 4, cu.c[0]) + so);
      l_ = ld2(EL_RING(T, 5, cu.c[0]) + so); lm = ld2(EL_RING(T, 6, cu.c[0]) + so); m_ = ld2(EL_RING(T, 7, cu.c[0]) + so);
    }
    __syncwarp();  // every lane has read its ring rows
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + b);
    if (act) {
      const i64 c = (i64)li * ld + q;
      if (sb > sa) {  // pending stress injection of the previous step (AddSource.cpp:69-84)
        sxx.x = el_apply_points(sxx.x, src, sa, sb, (int)c, 2, srcv_prev);
        syy.x = el_apply_points(syy.x, src, sa, sb, (int)c, 3, srcv_prev);
        sxy.x = el_apply_points(sxy.x, src, sa, sb, (int)c, 4, srcv_prev);
        sxx.y = el_apply_points(sxx.y, src, sa, sb, (int)c + 1, 2, srcv_prev);
        syy.y = el_apply_points(syy.y, src, sa, sb, (int)c + 1, 3, srcv_prev);
        sxy.y = el_apply_points(sxy.y, src, sa, sb, (int)c + 1, 4, srcv_prev);
      }
      double2 o1, o2, o3;
      {  // column q
        const double d1 = div_exact(27 * vxp1.x - 27 * vxc.x - vxp2.x + vxm1.x, hx, rx);
        const double d2 = div_exact(27 * vyc.x - 27 * vyl.y - vyc.y + vyl.x, hy, ry);
        const double d3 = div_exact(27 * vyc.x - 27 * vym1.x - vyp1.x + vym2.x, hx, rx);
        const double d4 = div_exact(27 * vxc.y - 27 * vxc.x - vxr.x + vxl, hy, ry);
        o1.x = sxx.x + (lm.x * d1 + l_.x * d2) * dt;
        o2.x = syy.x + (lm.x * d2 + l_.x * d1) * dt;
        o3.x = sxy.x + m_.x * (d3 + d4) * dt;
      }
      {  // column q+1
        const double d1 = div_exact(27 * vxp1.y - 27 * vxc.y - vxp2.y + vxm1.y, hx, rx);
        const double d2 = div_exact(27 * vyc.y - 27 * vyc.x - vyr + vyl.y, hy, ry);
        const double d3 = div_exact(27 * vyc.y - 27 * vym1.y - vyp1.y + vym2.y, hx, rx);
        const double d4 = div_exact(27 * vxr.x - 27 * vxc.y - vxr.y + vxc.x, hy, ry);
        o1.y = sxx.y + (lm.y * d1 + l_.y * d2) * dt;
        o2.y = syy.y + (lm.y * d2 + l_.y * d1) * dt;
        o3.y = sxy.y + m_.y * (d3 + d4) * dt;
      }
      st2(out.sxx + c, o1); st2(out.syy + c, o2); st2(out.sxy + c, o3);
    }
    vxm1 = vxc; vxc = vxp1; vxp1 = vxp2;
    vym2 = vym1; vym1 = vyc; vyc = vyp1;
    cu.next();
  }
}

__global__ void __launch_bounds__(EL_NT, 2)
el_sigma_fwd(ElGeom g, const ElCta* __restrict__ ctas, ElSlot in, ElSlot out, ElMat mt, ElCoef cf, ElPoints src,
             const double* __restrict__ srcv_prev, ElFuse f) {
  extern __shared__ __align__(128) unsigned char el_smem[];
  const int bid = el_bid(f);
  const ElCta d = el_cta(ctas, bid);
  bool t_lo, t_hi;
  el_cta_edges(g, f, d, &t_lo, &t_hi);
  el_fuse_wait(f, t_lo, t_hi);
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
  const int so = 2 + warp * 64 + 2 * lane;
  const int q = d.c0 + warp * 64 + 2 * lane;
  const bool act = q < d.c1;
  const bool wact = d.c0 + warp * 64 < d.c1;
  const double dt = g.dt, hx = g.h24x, hy = g.h24y, rx = g.r24x, ry = g.r24y;
  const double2 z2 = mk2(0.0, 0.0);
  // windows: sxy rows li-1, li, li+1 (+ li+2 new) ; sxx rows li-2, li-1, li (+ li+1 new)
  double2 qm1 = z2, qc = z2, qp1 = z2, xm2 = z2, xm1 = z2, xc = z2;
  if (act) {
    qm1 = ld2(out.sxy + (i64)(d.r0 - 1) * ld + q);
    xm2 = ld2(out.sxx + (i64)(d.r0 - 2) * ld + q);
    xm1 = ld2(out.sxx + (i64)(d.r0 - 1) * ld + q);
  }
  mbar_wait(bars, 0);
  if (wact) {
    qc = ld2(EL_RING(T, 0, 0) + so);
    qp1 = ld2(EL_RING(T, 0, 1) + so);
    xc = ld2(EL_RING(T, 1, 0) + so);
  }
  ElCursor cu;
  cu.init();
  for (int it = 0; it < nrows; it++) {
    const int li = d.r0 + it, b = it % EL_NB;
    mbar_wait(bars + 1 + b, (unsigned)(it / EL_NB) & 1u);
    double2 qp2 = z2, xp1 = z2, ql = z2, yy = z2, yr = z2, vx = z2, vy = z2, rh = z2, rb = z2;
    double qr = 0.0, yl = 0.0;
    if (wact) {
      qp2 = ld2(EL_RING(T, 0, cu.n[2]) + so);
      xp1 = ld2(EL_RING(T, 1, cu.n[1]) + so);
      const double* cq = EL_RING(T, 0, cu.c[2]);  // centre row of sxy: columns q-2, q-1, q+2
      ql = ld2(cq + so - 2); qr = cq[so + 2];
      const double* cy_ = EL_RING(T, 2, cu.c[0]);  // syy: columns q-1, q, q+1, q+2, q+3
      yl = cy_[so - 1]; yy = ld2(cy_ + so); yr = ld2(cy_ + so + 2);
      vx = ld2(EL_RING(T, 3, cu.c[0]) + so); vy = ld2(EL_RING(T, 4, cu.c[0]) + so);
      rh = ld2(EL_RING(T, 5, cu.c[0]) + so); rb = ld2(EL_RING(T, 6, cu.c[0]) + so);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + b);
    if (act) {
      const i64 c = (i64)li * ld + q;
      double2 o1, o2;
      {  // column q
        const double d5 = div_exact(27 * xc.x - 27 * xm1.x - xp1.x + xm2.x, hx, rx);
        const double d6 = div_exact(27 * qc.x - 27 * ql.y - qc.y + ql.x, hy, ry);
        const double d7 = div_exact(27 * qp1.x - 27 * qc.x - qp2.x + qm1.x, hx, rx);
        const double d8 = div_exact(27 * yy.y - 27 * yy.x - yr.x + yl, hy, ry);
        o1.x = vx.x + el_div_var((d5 + d6) * dt, rh.x);
        o2.x = vy.x + el_div_var((d7 + d8) * dt, rb.x);
      }
      {  // column q+1
        const double d5 = div_exact(27 * xc.y - 27 * xm1.y - xp1.y + xm2.y, hx, rx);
        const double d6 = div_exact(27 * qc.y - 27 * qc.x - qr + ql.y, hy, ry);
        const double d7 = div_exact(27 * qp1.y - 27 * qc.y - qp2.y + qm1.y, hx, rx);
        const double d8 = div_exact(27 * yr.x - 27 * yy.y - yr.y + yy.x, hy, ry);
        o1.y = vx.y + el_div_var((d5 + d6) * dt, rh.y);
        o2.y = vy.y + el_div_var((d7 + d8) * dt, rb.y);
      }
      st2(out.vx + c, o1); st2(out.vy + c, o2);
    }
    qm1 = qc; qc = qp1; qp1 = qp2;
    xm2 = xm1; xm1 = xc; xc = xp1;
    cu.next();
  }
}

__global__ void __launch_bounds__(EL_NT, 2)
el_vel_fwd(ElGeom g, const ElCta* __restrict__ ctas, ElSlot in, ElSlot out, ElMat mt, ElCoef cf, ElPoints src,
           const double* __restrict__ srcv_row, ElPoints rcv, double* __restrict__ rcvv, int rcv_stride, int slot,
           ElFuse f) {
  extern __shared__ __align__(128) unsigned char el_smem[];
  const int bid = el_bid(f);
  const ElCta d = el_cta(ctas, bid);
  bool t_lo, t_hi;
  el_cta_edges(g, f, d, &t_lo, &t_hi);
  el_fuse_wait(f, t_lo, t_hi);
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
  const int so = 2 + warp * 64 + 2 * lane;
  const int q = d.c0 + warp * 64 + 2 * lane;
  const bool act = q < d.c1;
  const bool wact = d.c0 + warp * 64 < d.c1;
  const double dt = g.dt, ix = 1.0 / (24 * g.dx), iy = 1.0 / (24 * g.dy);
  const double hx = g.h24x, hy = g.h24y, rx = g.r24x, ry = g.r24y;
  const double2 z2 = mk2(0.0, 0.0);
  auto prod = [dt](double2 v, double2 r) { return mk2(dt * v.x * r.x, dt * v.y * r.y); };
  // windows: A = dt*vbx*rinv rows li-1, li, li+1 (+ li+2) ; B = dt*vby*rbinv rows li-2, li-1, li (+ li+1)
  double2 Am1 = z2, Ac = z2, Ap1 = z2, Bm2 = z2, Bm1 = z2, Bc = z2;
  double2 fqm1 = z2, fqc = z2, fqp1 = z2, fsm2 = z2, fsm1 = z2, fsc = z2;  // forward sxy / sxx windows (MATGRAD)
  if (act) {
    const i64 o1 = (i64)(d.r0 - 1) * ld + q, o2 = (i64)(d.r0 - 2) * ld + q;
    Am1 = prod(ld2(b.vx + o1), ld2(mt.rinv + o1));
    Bm2 = prod(ld2(b.vy + o2), ld2(mt.rbinv + o2));
    Bm1 = prod(ld2(b.vy + o1), ld2(mt.rbinv + o1));
    if (MATGRAD) { fqm1 = ld2(fwd.sxy + o1); fsm2 = ld2(fwd.sxx + o2); fsm1 = ld2(fwd.sxx + o1); }
  }
  mbar_wait(bars, 0);
  if (wact) {
    Ac = prod(ld2(EL_RING(T, 0, 0) + so), ld2(EL_RING(T, 1, 0) + so));
    Ap1 = prod(ld2(EL_RING(T, 0, 1) + so), ld2(EL_RING(T, 1, 1) + so));
    Bc = prod(ld2(EL_RING(T, 2, 0) + so), ld2(EL_RING(T, 3, 0) + so));
    if (MATGRAD) { fqc = ld2(EL_RING(T, 7, 0) + so); fqp1 = ld2(EL_RING(T, 7, 1) + so); fsc = ld2(EL_RING(T, 8, 0) + so); }
  }
  ElCursor cu;
  cu.init();
  for (int it = 0; it < nrows; it++) {
    const int li = d.r0 + it, bb = it % EL_NB;
    mbar_wait(bars + 1 + bb, (unsigned)(it / EL_NB) & 1u);
    double2 Ap2 = z2, Bp1 = z2, Ar = z2, Bl = z2, sxx = z2, syy = z2, sxy = z2;
    double Al = 0.0, Br = 0.0;
    double2 vxc = z2, ric = z2, vyc = z2, rbc = z2;                       // raw centre values (MATGRAD)
    double2 fqp2 = z2, fsp1 = z2, fql = z2, fy = z2, fyr = z2, G3 = z2, G4 = z2;
    double fqr = 0.0, fyl = 0.0;
    if (wact) {
      Ap2 = prod(ld2(EL_RING(T, 0, cu.n[2]) + so), ld2(EL_RING(T, 1, cu.n[2]) + so));
      Bp1 = prod(ld2(EL_RING(T, 2, cu.n[1]) + so), ld2(EL_RING(T, 3, cu.n[1]) + so));
      const double* va = EL_RING(T, 0, cu.c[2]); const double* ria = EL_RING(T, 1, cu.c[2]);  // centre row of A: q-1, q+2, q+3
      Al = dt * va[so - 1] * ria[so - 1];
      Ar = prod(ld2(va + so + 2), ld2(ria + so + 2));
      const double* vb = EL_RING(T, 2, cu.c[1]); const double* rba = EL_RING(T, 3, cu.c[1]);  // centre row of B: q-2, q-1, q+2
      Bl = prod(ld2(vb + so - 2), ld2(rba + so - 2));
      Br = dt * vb[so + 2] * rba[so + 2];
      sxx = ld2(EL_RING(T, 4, cu.c[0]) + so); syy = ld2(EL_RING(T, 5, cu.c[0]) + so); sxy = ld2(EL_RING(T, 6, cu.c[0]) + so);
      if (MATGRAD) {
        vxc = ld2(va + so); ric = ld2(ria + so); vyc = ld2(vb + so); rbc = ld2(rba + so);
        fqp2 = ld2(EL_RING(T, 7, cu.n[2]) + so);
        fsp1 = ld2(EL_RING(T, 8, cu.n[1]) + so);
        const double* cq = EL_RING(T, 7, cu.c[2]);  // centre row of forward sxy: q-2, q-1, q+2
        fql = ld2(cq + so - 2); fqr = cq[so + 2];
        const double* cy_ = EL_RING(T, 9, cu.c[0]);  // forward syy: q-1 .. q+3
        fyl = cy_[so - 1]; fy = ld2(cy_ + so); fyr = ld2(cy_ + so + 2);
        G3 = ld2(EL_RING(T, 10, cu.c[0]) + so); G4 = ld2(EL_RING(T, 11, cu.c[0]) + so);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + bb);
    if (act) {
      const i64 c = (i64)li * ld + q;
      for (int k = ra; k < rb; k++) {  // stress-type receiver residuals of this slot (GetReceive.cpp:48-97)
        const int fd = rcv.field[k];
        if (fd >= 2 && (rcv.cell[k] == (int)c || rcv.cell[k] == (int)c + 1)) {
          double a = 0.0;
          for (int m = rcv.start[k]; m < rcv.start[k + 1]; m++) a += res[(i64)rcv.perm[m] * res_stride + slot];
          const bool hi = rcv.cell[k] == (int)c + 1;
          double2& t = fd == 2 ? sxx : (fd == 3 ? syy : sxy);
          if (hi) t.y += a; else t.x += a;
        }
      }
      // (D-x)^T A -> sxx ; (D-y)^T A -> sxy ; (D+x)^T B -> sxy ; (D+y)^T B -> syy
      sxx.x += (27 * Ac.x - 27 * Ap1.x - Am1.x + Ap2.x) * ix;
      sxy.x += (27 * Ac.x - 27 * Ac.y - Al + Ar.x) * iy;
      sxy.x += (27 * Bm1.x - 27 * Bc.x - Bm2.x + Bp1.x) * ix;
      syy.x += (27 * Bl.y - 27 * Bc.x - Bl.x + Bc.y) * iy;
      sxx.y += (27 * Ac.y - 27 * Ap1.y - Am1.y + Ap2.y) * ix;
      sxy.y += (27 * Ac.y - 27 * Ar.x - Ac.x + Ar.y) * iy;
      sxy.y += (27 * Bm1.y - 27 * Bc.y - Bm2.y + Bp1.y) * ix;
      syy.y += (27 * Bc.x - 27 * Bc.y - Bl.y + Br) * iy;
      st2(bout.sxx + c, sxx); st2(bout.syy + c, syy); st2(bout.sxy + c, sxy);
      if (MATGRAD) {
        {
          const double e56 = div_exact(27 * fsc.x - 27 * fsm1.x - fsp1.x + fsm2.x, hx, rx) +
                             div_exact(27 * fqc.x - 27 * fql.y - fqc.y + fql.x, hy, ry);
          const double e78 = div_exact(27 * fqp1.x - 27 * fqc.x - fqp2.x + fqm1.x, hx, rx) +
                             div_exact(27 * fy.y - 27 * fy.x - fyr.x + fyl, hy, ry);
          G3.x += -(dt * vxc.x) * e56 * (ric.x * ric.x);
          G4.x += -(dt * vyc.x) * e78 * (rbc.x * rbc.x);
        }
        {
          const double e56 = div_exact(27 * fsc.y - 27 * fsm1.y - fsp1.y + fsm2.y, hx, rx) +
                             div_exact(27 * fqc.y - 27 * fqc.x - fqr + fql.y, hy, ry);
          const double e78 = div_exact(27 * fqp1.y - 27 * fqc.y - fqp2.y + fqm1.y, hx, rx) +
                             div_exact(27 * fyr.x - 27 * fy.y - fyr.y + fy.x, hy, ry);
          G3.y += -(dt * vxc.y) * e56 * (ric.y * ric.y);
          G4.y += -(dt * vyc.y) * e78 * (rbc.y * rbc.y);
        }
        st2(Gr3 + c, G3); st2(Gr4 + c, G4);
      }
    }
    Am1 = Ac; Ac = Ap1; Ap1 = Ap2;
    Bm2 = Bm1; Bm1 = Bc; Bc = Bp1;
    if (MATGRAD) { fqm1 = fqc; fqc = fqp1; fqp1 = fqp2; fsm2 = fsm1; fsm1 = fsc; fsc = fsp1; }
    cu.next();
  }
}

template <bool MATGRAD>
__global__ void __launch_bounds__(EL_NT, MATGRAD ? 1 : 2)
el_vel_adj(ElGeom g, const ElCta* __restrict__ ctas, ElSlot b, ElSlot bout, ElSlot fwd, ElMat mt, ElCoef cf,
           double* __restrict__ Gr3, double* __restrict__ Gr4, ElPoints rcv, const double* __restrict__ res,
           int res_stride, int slot, ElFuse f) {
  extern __shared__ __align__(128) unsigned char el_smem[];
  const int bid = el_bid(f);
  const ElCta d = el_cta(ctas, bid);
  bool t_lo, t_hi;
  el_cta_edges(g, f, d, &t_lo, &t_hi);
  el_fuse_wait(f, t_lo, t_hi);
  int ra = 0, rb = 0;
  if (rcv.blk != nullptr && res != nullptr) { ra = rcv.blk[bid]; rb = rcv.blk[bid + 1]; }
  if (d.kind == 0) {
    double* ring = reinterpret_cast<double*>(el_smem + 64);
    el_vel_adj_march<MATGRAD>(g, d, b, bout, fwd, mt, Gr3, Gr4, rcv, ra, rb, res, res_stride, slot, ring,
                              reinterpret_cast<unsigned long long*>(el_smem));
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
  const int so = 2 + warp * 64 + 2 * lane;
  const int q = d.c0 + warp * 64 + 2 * lane;
  const bool act = q < d.c1;
  const bool wact = d.c0 + warp * 64 < d.c1;
  const double dt = g.dt, ix = 1.0 / (24 * g.dx), iy = 1.0 / (24 * g.dy);
  const double hx = g.h24x, hy = g.h24y, rx = g.r24x, ry = g.r24y;
  const double2 z2 = mk2(0.0, 0.0);
  // e3 = mub2 * (dt * sbxy) ; e1 = lmb*gx + lamb*gy ; e2 = lmb*gy + lamb*gx with gx = dt*sbxx, gy = dt*sbyy
  auto e3f = [dt](double2 s, double2 m) { return mk2(m.x * (dt * s.x), m.y * (dt * s.y)); };
  auto e1f = [dt](double2 sx, double2 sy, double2 lm, double2 l_) {
    return mk2(lm.x * (dt * sx.x) + l_.x * (dt * sy.x), lm.y * (dt * sx.y) + l_.y * (dt * sy.y));
  };
  // windows: E3 rows li-1, li, li+1 (+ li+2) ; E1 rows li-2, li-1, li (+ li+1)
  double2 E3m1 = z2, E3c = z2, E3p1 = z2, E1m2 = z2, E1m1 = z2, E1c = z2;
  double2 fxm1 = z2, fxc = z2, fxp1 = z2, fym2 = z2, fym1 = z2, fyc = z2;  // forward vx / vy windows (MATGRAD)
  if (act) {
    const i64 o1 = (i64)(d.r0 - 1) * ld + q, o2 = (i64)(d.r0 - 2) * ld + q;
    E3m1 = e3f(ld2(b.sxy + o1), ld2(mt.mub2 + o1));
    E1m2 = e1f(ld2(b.sxx + o2), ld2(b.syy + o2), ld2(mt.lmb + o2), ld2(mt.lamb + o2));
    E1m1 = e1f(ld2(b.sxx + o1), ld2(b.syy + o1), ld2(mt.lmb + o1), ld2(mt.lamb + o1));
    if (MATGRAD) { fxm1 = ld2(fwdv.vx + o1); fym2 = ld2(fwdv.vy + o2); fym1 = ld2(fwdv.vy + o1); }
  }
  mbar_wait(bars, 0);
  if (wact) {
    E3c = e3f(ld2(EL_RING(T, 0, 0) + so), ld2(EL_RING(T, 1, 0) + so));
    E3p1 = e3f(ld2(EL_RING(T, 0, 1) + so), ld2(EL_RING(T, 1, 1) + so));
    E1c = e1f(ld2(EL_RING(T, 2, 0) + so), ld2(EL_RING(T, 3, 0) + so), ld2(EL_RING(T, 4, 0) + so), ld2(EL_RING(T, 5, 0) + so));
    if (MATGRAD) { fxc = ld2(EL_RING(T, 8, 0) + so); fxp1 = ld2(EL_RING(T, 8, 1) + so); fyc = ld2(EL_RING(T, 9, 0) + so); }
  }
  ElCursor cu;
  cu.init();
  for (int it = 0; it < nrows; it++) {
    const int li = d.r0 + it, bb = it % EL_NB;
    mbar_wait(bars + 1 + bb, (unsigned)(it / EL_NB) & 1u);
    double2 E3p2 = z2, E1p1 = z2, E3l = z2, E2c = z2, E2r = z2, vx = z2, vy = z2;
    double E3r = 0.0, E2l = 0.0;
    double2 sxc = z2, syc = z2, sqc = z2;                                   // raw centre sigma_bar (MATGRAD)
    double2 fxp2 = z2, fyp1 = z2, fxr = z2, fyl = z2, GL = z2, GM1 = z2, GM2 = z2;
    double fxl = 0.0, fyr = 0.0;
    if (wact) {
      E3p2 = e3f(ld2(EL_RING(T, 0, cu.n[2]) + so), ld2(EL_RING(T, 1, cu.n[2]) + so));
      E1p1 = e1f(ld2(EL_RING(T, 2, cu.n[1]) + so), ld2(EL_RING(T, 3, cu.n[1]) + so), ld2(EL_RING(T, 4, cu.n[1]) + so),
                 ld2(EL_RING(T, 5, cu.n[1]) + so));
      const double* sq = EL_RING(T, 0, cu.c[2]); const double* mu = EL_RING(T, 1, cu.c[2]);  // centre row of e3: q-2, q-1, q+2
      E3l = e3f(ld2(sq + so - 2), ld2(mu + so - 2));
      E3r = mu[so + 2] * (dt * sq[so + 2]);
      const double* sx = EL_RING(T, 2, cu.c[1]); const double* sy = EL_RING(T, 3, cu.c[1]);   // centre row of e2: q-1 .. q+3
      const double* lm = EL_RING(T, 4, cu.c[1]); const double* l_ = EL_RING(T, 5, cu.c[1]);
      E2l = lm[so - 1] * (dt * sy[so - 1]) + l_[so - 1] * (dt * sx[so - 1]);
      E2c = e1f(ld2(sy + so), ld2(sx + so), ld2(lm + so), ld2(l_ + so));          // e2 = e1 with gx <-> gy
      E2r = e1f(ld2(sy + so + 2), ld2(sx + so + 2), ld2(lm + so + 2), ld2(l_ + so + 2));
      vx = ld2(EL_RING(T, 6, cu.c[0]) + so); vy = ld2(EL_RING(T, 7, cu.c[0]) + so);
      if (MATGRAD) {
        sxc = ld2(sx + so); syc = ld2(sy + so); sqc = ld2(sq + so);
        fxp2 = ld2(EL_RING(T, 8, cu.n[2]) + so);
        fyp1 = ld2(EL_RING(T, 9, cu.n[1]) + so);
        const double* cx_ = EL_RING(T, 8, cu.c[2]);  // centre row of forward vx: q-1, q+2, q+3
        fxl = cx_[so - 1]; fxr = ld2(cx_ + so + 2);
        const double* cy_ = EL_RING(T, 9, cu.c[1]);  // centre row of forward vy: q-2, q-1, q+2
        fyl = ld2(cy_ + so - 2); fyr = cy_[so + 2];
        GL = ld2(EL_RING(T, 10, cu.c[0]) + so); GM1 = ld2(EL_RING(T, 11, cu.c[0]) + so); GM2 = ld2(EL_RING(T, 12, cu.c[0]) + so);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 1 + EL_NB + bb);
    if (act) {
      const i64 c = (i64)li * ld + q;
      // (D+x)^T e1 -> vx ; (D-y)^T e2 -> vy ; (D-x)^T e3 -> vy ; (D+y)^T e3 -> vx
      vx.x += (27 * E1m1.x - 27 * E1c.x - E1m2.x + E1p1.x) * ix;
      vy.x += (27 * E2c.x - 27 * E2c.y - E2l + E2r.x) * iy;
      vy.x += (27 * E3c.x - 27 * E3p1.x - E3m1.x + E3p2.x) * ix;
      vx.x += (27 * E3l.y - 27 * E3c.x - E3l.x + E3c.y) * iy;
      vx.y += (27 * E1m1.y - 27 * E1c.y - E1m2.y + E1p1.y) * ix;
      vy.y += (27 * E2c.y - 27 * E2r.x - E2c.x + E2r.y) * iy;
      vy.y += (27 * E3c.y - 27 * E3p1.y - E3m1.y + E3p2.y) * ix;
      vx.y += (27 * E3c.x - 27 * E3c.y - E3l.y + E3r) * iy;
      st2(bout.vx + c, vx); st2(bout.vy + c, vy);
      if (MATGRAD) {
        {
          const double gx = dt * sxc.x, gy = dt * syc.x, gg = dt * sqc.x;
          const double e1 = div_exact(27 * fxp1.x - 27 * fxc.x - fxp2.x + fxm1.x, hx, rx);
          const double e2 = div_exact(27 * fyc.x - 27 * fyl.y - fyc.y + fyl.x, hy, ry);
          const double e34 = div_exact(27 * fyc.x - 27 * fym1.x - fyp1.x + fym2.x, hx, rx) +
                             div_exact(27 * fxc.y - 27 * fxc.x - fxr.x + fxl, hy, ry);
          GL.x += (gx + gy) * (e1 + e2);
          GM1.x += 2 * (gx * e1 + gy * e2);
          GM2.x += gg * e34;
        }
        {
          const double gx = dt * sxc.y, gy = dt * syc.y, gg = dt * sqc.y;
          const double e1 = div_exact(27 * fxp1.y - 27 * fxc.y - fxp2.y + fxm1.y, hx, rx);
          const double e2 = div_exact(27 * fyc.y - 27 * fyc.x - fyr + fyl.y, hy, ry);
          const double e34 = div_exact(27 * fyc.y - 27 * fym1.y - fyp1.y + fym2.y, hx, rx) +
                             div_exact(27 * fxr.x - 27 * fxc.y - fxr.y + fxc.x, hy, ry);
          GL.y += (gx + gy) * (e1 + e2);
          GM1.y += 2 * (gx * e1 + gy * e2);
          GM2.y += gg * e34;
        }
        st2(Gl + c, GL); st2(Gm1 + c, GM1); st2(Gm2 + c, GM2);
      }
    }
    E3m1 = E3c; E3c = E3p1; E3p1 = E3p2;
    E1m2 = E1m1; E1m1 = E1c; E1c = E1p1;
    if (MATGRAD) { fxm1 = fxc; fxc = fxp1; fxp1 = fxp2; fym2 = fym1; fym1 = fyc; fyc = fyp1; }
    cu.next();
  }
}

template <bool MATGRAD>
__global__ void __launch_bounds__(EL_NT, 1)
el_sigma_adj(ElGeom g, const ElCta* __restrict__ ctas, ElSlot b, ElSlot bout, ElSlot fwdv, ElSlot fwdm, ElMat mt,
             ElCoef cf, double* __restrict__ Gl, double* __restrict__ Gm1, double* __restrict__ Gm2, ElPoints rcv,
             const double* __restrict__ res, int res_stride, int slot_prev, ElPoints src,
             double* __restrict__ gsrcv_row, ElFuse f) {
  extern __shared__ __align__(128) unsigned char el_smem[];
  const int bid = el_bid(f);
  const ElCta d = el_cta(ctas, bid);
  bool t_lo, t_hi;
  el_cta_edges(g, f, d, &t_lo, &t_hi);
  el_fuse_wait(f, t_lo, t_hi);
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
