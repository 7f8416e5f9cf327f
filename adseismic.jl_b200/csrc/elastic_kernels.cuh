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
// State layout.  A wavefield SLOT is 5 pitched planes (vx, vy, sxx, syy, sxy) + the 8 CPML memories stored
// COMPACTLY: x-memories only on the rows where the x-profile is non-zero (all columns), y-memories only on the
// columns where the y-profile is non-zero (all rows) -- ~4 % of a slot at 2000^2.  Stress sources are injected
// LAZILY: a slot holds v AFTER and sigma BEFORE the injection of its step, and el_sigma_fwd adds the pending stress
// sources when it loads sigma (same `+=` arithmetic and order as the reference).  That keeps in the history exactly
// the quantities the adjoint needs (fw3/fw4 see pre-injection stresses, fw1/fw2 post-injection velocities) and
// avoids a separate injection launch racing with stencil reads.
// Forward arithmetic follows the reference's evaluation order (bit-identical with -fmad=false).
#pragma once
#include "common.cuh"

#define EL_BX 64   // threads along columns (fast axis)
#define EL_BY 4    // threads along rows
#define EL_THREADS (EL_BX * EL_BY)
#define EL_ROWS 16 // rows per CTA (EL_ROWS/EL_BY iterations)

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
  i64 plane;      // Hl*ld
  i64 xm_sz, ym_sz;  // doubles per compact x-/y-memory array
  int ntc, ntr;   // CTA tiling: column tiles of EL_BX, row tiles of EL_ROWS over local rows [own0, own1)
  int own0, own1;
};

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
// null / zero on a single GPU.  The row tiles that contain my first / last owned rows are launched first (perm),
// wait until the neighbour's PREVIOUS launch has delivered the halo rows they read, and -- after their epilogue --
// store their columns of my two edge rows of every field this launch produces straight into the neighbour's halo
// rows over NVLink, then publish with a system-scope fence + atomic on the neighbour's flag.  Replaces the 18
// mpi_halo_exchange2 ops per step of src/MPIElastic.jl:411-436, 515-516, 551, 582, 618.
#define EL_HALO 2
struct ElFuse {
  const int* perm;              // launch order -> logical CTA id (edge row tiles first), or null
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

__device__ __forceinline__ void el_fuse_wait(const ElFuse& f, bool t_lo, bool t_hi) {
  if (!(t_lo || t_hi)) return;  // CTA-uniform
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    volatile unsigned long long* fl = f.my_flags;
    unsigned long long spins = 0;
    while ((t_lo && fl[3] < f.expect_lo) || (t_hi && fl[4] < f.expect_hi)) {
      if (++spins > (1ULL << 26)) { fl[2] = 1ULL; break; }  // neighbour lost: report, do not hang
    }
    __threadfence_system();
  }
  __syncthreads();
}

// A row tile is an edge tile when it holds one of my first / last EL_HALO owned rows next to a neighbour: those are
// the rows it must push, and (a superset of) the tiles whose stencils read halo rows.
__device__ __forceinline__ void el_tile_edges(const ElGeom& g, const ElFuse& f, int tr, bool* t_lo, bool* t_hi) {
  const int ra = g.own0 + tr * EL_ROWS, rb = min(g.own1, ra + EL_ROWS);
  *t_lo = f.has_lo && ra < g.own0 + EL_HALO;
  *t_hi = f.has_hi && rb > g.own1 - EL_HALO;
}

// push this tile's columns of my edge rows, then signal (called by all threads of an edge CTA)
__device__ __forceinline__ void el_fuse_push(const ElGeom& g, const ElFuse& f, bool t_lo, bool t_hi, int tr, int q) {
  if (!(t_lo || t_hi)) return;
  __syncthreads();  // all cells (and point injections) of this CTA are written
  if (q < g.ld) {
    const int ra = g.own0 + tr * EL_ROWS, rb = min(g.own1, ra + EL_ROWS);
    for (int k = 0; k < f.nf; k++) {
      // threadIdx.y = 0..3 -> (row 0/1) x (lo/hi)
      const int r = threadIdx.y & 1;
      if ((threadIdx.y >> 1) == 0) {
        const int li = g.own0 + r;
        if (t_lo && li >= ra && li < rb) f.lo[k][(i64)r * g.ld + q] = f.src[k][(i64)li * g.ld + q];
      } else {
        const int li = g.own1 - EL_HALO + r;
        if (t_hi && li >= ra && li < rb) f.hi[k][(i64)r * g.ld + q] = f.src[k][(i64)li * g.ld + q];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
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
// forward sigma pass: fw1 + fw2   (src/Core.jl:96-155, src/MPIElastic.jl:483-555)
//   in : slot s-1 (v post-injection, sigma pre-injection, memories)      out: sigma and mem1..4 of slot s
//   `src`/`srcv_prev`: stress sources of step s-1 that are still pending (null for s == 1)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EL_THREADS)
el_sigma_fwd(ElGeom g, ElSlot in, ElSlot out, ElMat mt, ElCoef cf, ElPoints src, const double* __restrict__ srcv_prev,
             ElFuse f) {
  const int bid = el_bid(f);
  const int tc = bid % g.ntc, tr = bid / g.ntc;
  const int q = tc * EL_BX + threadIdx.x;
  bool t_lo, t_hi;
  el_tile_edges(g, f, tr, &t_lo, &t_hi);
  el_fuse_wait(f, t_lo, t_hi);
  int sa = 0, sb = 0;
  if (src.blk != nullptr && srcv_prev != nullptr) { sa = src.blk[bid]; sb = src.blk[bid + 1]; }
  const double* __restrict__ vx = in.vx;
  const double* __restrict__ vy = in.vy;
  const int ld = g.ld, NX = g.NX, NY = g.NY;
  const double dt = g.dt, dx = g.dx, dy = g.dy;
#pragma unroll 1
  for (int it = 0; it < EL_ROWS / EL_BY; it++) {
    const int li = g.own0 + tr * EL_ROWS + it * EL_BY + threadIdx.y;
    if (li >= g.own1 || q >= ld) continue;
    const int gp = g.goff + li;
    const i64 c = (i64)li * ld + q;
    double sxx = in.sxx[c], syy = in.syy[c], sxy = in.sxy[c];
    if (sb > sa) {  // pending stress injection of the previous step (AddSource.cpp:69-84)
      sxx = el_apply_points(sxx, src, sa, sb, (int)c, 2, srcv_prev);
      syy = el_apply_points(syy, src, sa, sb, (int)c, 3, srcv_prev);
      sxy = el_apply_points(sxy, src, sa, sb, (int)c, 4, srcv_prev);
    }
    const int kx = gp - g.cx, ky = q - g.cy;
    if (q < g.W && el_in(g, 0, gp, q)) {  // fw1
      const double l_ = mt.lamb[c], lm = mt.lmb[c];
      double d1 = (27 * vx[c + ld] - 27 * vx[c] - vx[c + 2 * ld] + vx[c - ld]) / (24 * dx);
      double d2 = (27 * vy[c] - 27 * vy[c - 1] - vy[c + 1] + vy[c - 2]) / (24 * dy);
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
    if (q < g.W && el_in(g, 1, gp, q)) {  // fw2
      const double m_ = mt.mub2[c];
      double d3 = (27 * vy[c] - 27 * vy[c - ld] - vy[c + ld] + vy[c - 2 * ld]) / (24 * dx);
      double d4 = (27 * vx[c + 1] - 27 * vx[c] - vx[c + 2] + vx[c - 1]) / (24 * dy);
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
  el_fuse_push(g, f, t_lo, t_hi, tr, q);
}

// ------------------------------------------------------------------------------------------------------------
// forward velocity pass: fw3 + fw4 (src/Core.jl:158-213), then add_source (velocity types now, stress types lazily)
// and get_receive for this slot.
//   in : v of slot s-1, sigma (already updated) of slot s       out: v and mem5..8 of slot s
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EL_THREADS)
el_vel_fwd(ElGeom g, ElSlot in, ElSlot out, ElMat mt, ElCoef cf, ElPoints src, const double* __restrict__ srcv_row,
           ElPoints rcv, double* __restrict__ rcvv, int rcv_stride, int slot, ElFuse f) {
  const int bid = el_bid(f);
  const int tc = bid % g.ntc, tr = bid / g.ntc;
  const int q = tc * EL_BX + threadIdx.x;
  bool t_lo, t_hi;
  el_tile_edges(g, f, tr, &t_lo, &t_hi);
  el_fuse_wait(f, t_lo, t_hi);
  const double* __restrict__ sxx = out.sxx;
  const double* __restrict__ syy = out.syy;
  const double* __restrict__ sxy = out.sxy;
  const int ld = g.ld, NX = g.NX, NY = g.NY;
  const double dt = g.dt, dx = g.dx, dy = g.dy;
#pragma unroll 1
  for (int it = 0; it < EL_ROWS / EL_BY; it++) {
    const int li = g.own0 + tr * EL_ROWS + it * EL_BY + threadIdx.y;
    if (li >= g.own1 || q >= ld) continue;
    const int gp = g.goff + li;
    const i64 c = (i64)li * ld + q;
    double vx = in.vx[c], vy = in.vy[c];
    const int kx = gp - g.cx, ky = q - g.cy;
    if (q < g.W && el_in(g, 2, gp, q)) {  // fw3
      double d5 = (27 * sxx[c] - 27 * sxx[c - ld] - sxx[c + ld] + sxx[c - 2 * ld]) / (24 * dx);
      double d6 = (27 * sxy[c] - 27 * sxy[c - 1] - sxy[c + 1] + sxy[c - 2]) / (24 * dy);
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
      vx += (d5 + d6) * dt / mt.rho[c];
    }
    if (q < g.W && el_in(g, 3, gp, q)) {  // fw4
      const double r_ = mt.rhob[c];
      double d7 = (27 * sxy[c + ld] - 27 * sxy[c] - sxy[c + 2 * ld] + sxy[c - ld]) / (24 * dx);
      double d8 = (27 * syy[c + 1] - 27 * syy[c] - syy[c + 2] + syy[c - 1]) / (24 * dy);
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
      vy += (d7 + d8) * dt / r_;
    }
    out.vx[c] = vx; out.vy[c] = vy;
  }
  // ---- epilogue: velocity sources of this step, then receivers of this slot ----
  int ia = 0, ib = 0, ra = 0, rb = 0;
  if (src.blk != nullptr && srcv_row != nullptr) { ia = src.blk[bid]; ib = src.blk[bid + 1]; }
  if (rcv.blk != nullptr && rcvv != nullptr) { ra = rcv.blk[bid]; rb = rcv.blk[bid + 1]; }
  if (ib != ia || rb != ra) {  // CTA-uniform
    __syncthreads();
    const int tid = threadIdx.y * EL_BX + threadIdx.x;
    for (int k = ia + tid; k < ib; k += EL_THREADS) {
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
      for (int k = ra + tid; k < rb; k += EL_THREADS) {
        const int fd = rcv.field[k];
        double v = el_field(out, fd)[rcv.cell[k]];
        if (fd >= 2 && srcv_row != nullptr)  // post-injection value of a stress component
          for (int m = rcv.xstart[k]; m < rcv.xstart[k + 1]; m++) v += srcv_row[rcv.xperm[m]];
        for (int m = rcv.start[k]; m < rcv.start[k + 1]; m++) rcvv[(i64)rcv.perm[m] * rcv_stride + slot] = v;
      }
    }
  }
  el_fuse_push(g, f, t_lo, t_hi, tr, q);
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
__global__ void __launch_bounds__(EL_THREADS)
el_vel_adj(ElGeom g, ElSlot b, ElSlot bout, ElSlot fwd, ElMat mt, ElCoef cf, double* __restrict__ Gr3,
           double* __restrict__ Gr4, ElPoints rcv, const double* __restrict__ res, int res_stride, int slot,
           ElFuse f) {
  const int bid = el_bid(f);
  const int tc = bid % g.ntc, tr = bid / g.ntc;
  const int q = tc * EL_BX + threadIdx.x;
  bool t_lo, t_hi;
  el_tile_edges(g, f, tr, &t_lo, &t_hi);
  el_fuse_wait(f, t_lo, t_hi);
  int ra = 0, rb = 0;
  if (rcv.blk != nullptr && res != nullptr) { ra = rcv.blk[bid]; rb = rcv.blk[bid + 1]; }
  const ElAdjCtx A{g, b, mt, cf};
  const int ld = g.ld, NX = g.NX, NY = g.NY;
  const double ix = 1.0 / (24 * g.dx), iy = 1.0 / (24 * g.dy);
#pragma unroll 1
  for (int it = 0; it < EL_ROWS / EL_BY; it++) {
    const int li = g.own0 + tr * EL_ROWS + it * EL_BY + threadIdx.y;
    if (li >= g.own1 || q >= ld) continue;
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
    if (q < g.W) {
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
          e56 = (27 * fs[c] - 27 * fs[c - ld] - fs[c + ld] + fs[c - 2 * ld]) / (24 * g.dx) +
                (27 * fq[c] - 27 * fq[c - 1] - fq[c + 1] + fq[c - 2]) / (24 * g.dy);
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
          e78 = (27 * fq[c + ld] - 27 * fq[c] - fq[c + 2 * ld] + fq[c - ld]) / (24 * g.dx) +
                (27 * fy[c + 1] - 27 * fy[c] - fy[c + 2] + fy[c - 1]) / (24 * g.dy);
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
    }
    bout.sxx[c] = sxx; bout.syy[c] = syy; bout.sxy[c] = sxy;
  }
  el_fuse_push(g, f, t_lo, t_hi, tr, q);
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
  const int tid = threadIdx.y * EL_BX + threadIdx.x;
  for (int k = ra + tid; k < rb; k += EL_THREADS) {
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
    for (int k = sa + tid; k < sb; k += EL_THREADS) {
      const int f = src.field[k];
      double v = el_field(bout, f)[src.cell[k]];
      if (f >= 2 && res != nullptr)  // stress residuals of that slot are still pending on sigma_bar
        for (int m = src.xstart[k]; m < src.xstart[k + 1]; m++) v += res[(i64)src.xperm[m] * res_stride + slot_prev];
      for (int m = src.start[k]; m < src.start[k + 1]; m++) gsrcv_row[src.perm[m]] = v;
    }
  }
}

__global__ void __launch_bounds__(EL_THREADS)
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
__global__ void __launch_bounds__(EL_THREADS)
el_sigma_adj(ElGeom g, ElSlot b, ElSlot bout, ElSlot fwdv, ElSlot fwdm, ElMat mt, ElCoef cf, double* __restrict__ Gl,
             double* __restrict__ Gm1, double* __restrict__ Gm2, ElPoints rcv, const double* __restrict__ res,
             int res_stride, int slot_prev, ElPoints src, double* __restrict__ gsrcv_row, ElFuse f) {
  const int bid = el_bid(f);
  const int tc = bid % g.ntc, tr = bid / g.ntc;
  const int q = tc * EL_BX + threadIdx.x;
  bool t_lo, t_hi;
  el_tile_edges(g, f, tr, &t_lo, &t_hi);
  el_fuse_wait(f, t_lo, t_hi);
  const ElAdjCtx A{g, b, mt, cf};
  const int ld = g.ld, NX = g.NX, NY = g.NY;
  const double ix = 1.0 / (24 * g.dx), iy = 1.0 / (24 * g.dy);
#pragma unroll 1
  for (int it = 0; it < EL_ROWS / EL_BY; it++) {
    const int li = g.own0 + tr * EL_ROWS + it * EL_BY + threadIdx.y;
    if (li >= g.own1 || q >= ld) continue;
    const int gp = g.goff + li;
    const i64 c = (i64)li * ld + q;
    double vx = b.vx[c], vy = b.vy[c];
    if (q < g.W) {
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
          e1 = (27 * fx[c + ld] - 27 * fx[c] - fx[c + 2 * ld] + fx[c - ld]) / (24 * g.dx);
          e2 = (27 * fy[c] - 27 * fy[c - 1] - fy[c + 1] + fy[c - 2]) / (24 * g.dy);
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
          e34 = (27 * fy[c] - 27 * fy[c - ld] - fy[c + ld] + fy[c - 2 * ld]) / (24 * g.dx) +
                (27 * fx[c + 1] - 27 * fx[c] - fx[c + 2] + fx[c - 1]) / (24 * g.dy);
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
    }
    bout.vx[c] = vx; bout.vy[c] = vy;
  }
  el_adj_epilogue(g, bout, rcv, res, res_stride, slot_prev, src, gsrcv_row, bid);
  el_fuse_push(g, f, t_lo, t_hi, tr, q);
}
