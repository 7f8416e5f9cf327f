// acoustic_kernels.cuh -- fused acoustic time-step kernels (forward and exact adjoint) for sm_100a.
//
// One launch = one time step over the rows this GPU owns:
//   forward : u[s]   = Step(u[s-1], u[s-2], phi, psi; c^2, sigma_x(i), tau_y(j))  + source injection into u[s]
//                      + receiver sampling of u[s]                       (reference: AcousticOneStepCpu.h:1-48,
//                      ScatterAddOps.h:1-7 via Core.jl:600-601, gather via Core.jl:727-728)
//   adjoint : ubar[s-1] = Step^T in GATHER form (scatter form AcousticOneStepCpu.h:51-125)
//                      + cross-correlation gbar_c2 += cbar_s in the same pass
//                      + receiver-residual injection into ubar[s-1] + grad_srcv sampling
//
// Work decomposition (HBM-bound fp64 stencil, no tensor cores).  The grid of one launch holds two kinds of CTAs:
//   * MARCHING CTAs cover the PML-free box.  A CTA owns a 512-column tile and marches down `rb` rows.  One producer
//     warp streams a row of every input plane per iteration into a shared-memory ring with TMA bulk copies
//     (cp.async.bulk, one 4-KB copy per plane and row, mbarrier full/empty handshake, 8 / 4 rows in flight); the
//     eight consumer warps own 64 consecutive columns each (one double2 per lane), keep the 3-row stencil window
//     in registers, take left/right neighbours from warp shuffles and the two warp-edge values from the ring.
//     Every element is requested from DRAM once per step: 32 B/cell forward (w, wold, c^2 in, u out), 56 B/cell
//     adjoint.
//   * FRAME CTAs cover everything else (absorbing frame, ring, pitch padding) one cell per thread-iteration with
//     the full reference expression (phi/psi traffic, sigma/tau profiles, the fp64 divide).  The frame is ~1 % of
//     a 4096^2 grid, so it is mapped for parallelism, not reuse.
// Sources and receivers are injected / sampled by the CTA that owns their cell, after a CTA barrier, through
// per-CTA CSR lists built on the host (no atomics; duplicates accumulate in the reference's sequential order).
// Forward arithmetic is written in the reference's evaluation order and the library is compiled with
// -fmad=false, so forward wavefields and traces are bit-identical to the CPU op.
#pragma once
#include "common.cuh"

#define AC_WARPS 8
#define AC_THREADS (AC_WARPS * 32)
#define AC_WCOLS 64                         // columns per warp
#define AC_TILE_COLS (AC_WARPS * AC_WCOLS)  // columns per marching CTA
#ifndef AC_FRAME_CPT
#define AC_FRAME_CPT 4                      // frame cells per thread
#endif
#define AC_FRAME_CELLS (AC_THREADS * AC_FRAME_CPT)
#ifndef AC_NST_ADJ
#define AC_NST_ADJ 4                        // adjoint: ring depth (rows of the CTA tile in flight)
#endif
#ifndef AC_MINB_ADJ
#define AC_MINB_ADJ 2                       // __launch_bounds__ min CTAs/SM, adjoint (bounded by shared memory)
#endif
#define AC_ADJ_THREADS (AC_THREADS + 32)    // + one producer warp
#define AC_HCOLS (AC_TILE_COLS + 4)         // staged columns of an array with y-neighbours: 16-byte halo on each side
// Shared-memory footprint of a halo'd stage row.  Padded to a multiple of 128 bytes so that no two bulk copies in
// flight ever write into the same 128-byte shared-memory line (the copy of a 516-double row would otherwise end 32
// bytes into the line the next array's copy starts in).  -DADSEIS_NO_SMEM_PAD restores the packed round-1 layout for
// the A/B determinism experiment (scripts/ring_race_ab.sh).
#ifdef ADSEIS_NO_SMEM_PAD
#define AC_HPAD AC_HCOLS
#else
#define AC_HPAD ((AC_HCOLS + 15) / 16 * 16)
#endif

// One ring stage of an adjoint marching CTA: row `li` of its 512-column tile (iteration li - r0).  Filled by five
// 4-KB bulk copies (one UBLKCP per array: the TMA unit retires ~one op per 45 cycles whatever its size, so the
// copies must be CTA-wide, not per warp), completion on the stage's `full` mbarrier; the eight consumer warps
// release it through the `empty` mbarrier.
struct __align__(128) AcAdjStage {
  double ub1[AC_HPAD], c2[AC_HPAD], wf[AC_HPAD];     // row li+1, columns c0-2 .. c0+513
  double ub2[AC_TILE_COLS], G[AC_TILE_COLS];         // row li,   columns c0   .. c0+511
};
#ifndef AC_MINB_FWD
#define AC_MINB_FWD 2                       // __launch_bounds__ min CTAs/SM, forward (same wave size as the adjoint)
#endif
#ifndef AC_NST_FWD
#define AC_NST_FWD 4   // power of two (slot = it % depth).  Depth 8 (198 KB per SM) is 2.7 % faster (92.1 vs 94.7 us at 4096^2)
                       // but made the forward sweep NON-DETERMINISTIC at full size (scripts/determinism_probe.py: 3 of 11
                       // repeat runs differ in small patches of the precursor zone); depth 4 has a clean record
#endif
#define AC_FWD_THREADS (AC_THREADS + 32)
struct __align__(128) AcFwdStage {
  double w[AC_HPAD];                                 // row li+1, columns c0-2 .. c0+513
  double wold[AC_TILE_COLS], c2[AC_TILE_COLS];       // row li
};
#ifndef AC_FWD_SMEM_EXTRA
#define AC_FWD_SMEM_EXTRA 0                 // experiments: unused shared memory (lowers the CTAs per SM)
#endif
#define AC_FWD_SMEM ((int)(AC_NST_FWD * (sizeof(AcFwdStage) + 16)) + AC_FWD_SMEM_EXTRA)
#define AC_ADJ_SMEM ((int)(AC_NST_ADJ * (sizeof(AcAdjStage) + 16)))

#ifdef AC_RING_DEBUG
// Debug build only (scripts/ring_race_ab.sh): every consumer lane re-reads its ring data straight from global memory
// and records mismatches -- tells "stage read before the copy landed" (old row) from "overwritten early" (later row).
__device__ unsigned int g_ring_dbg_n;
__device__ double g_ring_dbg[64 * 12];
#endif

struct AcGeom {
  int H, W;    // global padded rows (NX+2) and columns (NY+2)
  int Hl, ld;  // local rows held by this GPU (incl. halo rows) and pitch in doubles
  int goff;    // global row index of local row 0
  double dt, hx, hy;
  double kx2, ky2;  // 2*dt*dt/hx/hx , 2*dt*dt/hy/hy
  double rx, ry;    // dt/hx , dt/hy
  double px, py;    // dt*dt/(2.0*hx) , dt*dt/(2.0*hy)
  double dt2;       // dt*dt
  double rhx, rhy;  // RN(1/hx), RN(1/hy)
  i64 plane;        // Hl*ld
};

// CTA -> work mapping of one launch (built on the host, passed by value)
struct AcTiling {
  int mr0, mr1;       // marched local rows [mr0, mr1)
  int mc0, mc_end;    // marched columns [mc0, mc_end), mc0 % 16 == 0, both even
  int rb;             // rows per marching CTA (uniform tiles of the middle rows)
  int elo, ehi;       // slab plans: rows of the thin first / last row tile next to a neighbour (0 = none)
  int nct, ntr;       // marching column tiles / row tiles; marching CTA id = tr*nct + ct
  int nmarch;         // nct*ntr
  int nrect;          // frame rectangles (local rows [rr0,rr1) x columns [rc0,rc1))
  int rr0[4], rr1[4], rc0[4], rc1[4];
  int rblk[5];        // CTA-id prefix of the rectangles, relative to nmarch
  int bx_r0, bx_r1, bx_c0, bx_c1;  // frame cells inside this box (the rim ring a WIDE frame-only launch recomputes so that the
                      // next frame launch does not depend on the concurrent box-pair launch) store their state only: the
                      // box kernel owns their Gbar / phibar / psibar updates
  int fcpt;           // frame cells per thread (a frame CTA covers fthr * fcpt consecutive cells of its rectangle):
                      // AC_FRAME_CPT when frame CTAs share a launch with marching CTAs, 1 in frame-only launches
  int fthr;           // threads of a frame CTA that own cells: AC_THREADS in a full launch, AC_FO_THREADS in a frame-only one
};
// Frame-only launches of the two-step path run CONCURRENTLY with the box-pair kernel (second stream).  That only works if
// their CTAs fit into what two resident box-pair CTAs leave of an SM (10 K registers next to ac_fwd2_kernel, 13 K next to
// ac_adj2_kernel): 128 threads at <= 80 registers.  With the 288-thread CTAs of the full kernels the frame launches sat
// in the queue until the box kernel drained (measured: 7 us between the first and the last CTA entering).
#define AC_FO_THREADS 128

// Row tile tr of the marched rows -> [r0, r1).  Slab plans put a thin tile on the rows next to a neighbour: those CTAs
// are launched first and finish within a couple of microseconds, so the halo rows they push travel over NVLink while
// the rest of the (single-wave) launch is still streaming, instead of leaving at the very end of it.
__host__ __device__ __forceinline__ void ac_row_tile(const AcTiling& t, int tr, int* r0, int* r1) {
  const int m0 = t.mr0 + t.elo, m1 = t.mr1 - t.ehi;  // uniform middle
  if (t.elo && tr == 0) { *r0 = t.mr0; *r1 = m0; return; }
  const int a = m0 + (tr - (t.elo ? 1 : 0)) * t.rb;
  if (a >= m1) { *r0 = m1; *r1 = t.mr1; return; }
  *r0 = a; *r1 = (a + t.rb < m1) ? a + t.rb : m1;
}
__host__ __device__ __forceinline__ int ac_row_tile_of(const AcTiling& t, int li) {
  const int m0 = t.mr0 + t.elo, m1 = t.mr1 - t.ehi;
  if (li < m0) return 0;
  if (li >= m1) return (t.elo ? 1 : 0) + (m1 - m0 + t.rb - 1) / t.rb;
  return (t.elo ? 1 : 0) + (li - m0) / t.rb;
}

// Per-CTA point lists: points (sources or receivers) owned by CTA b are entries [blk[b], blk[b+1]) of the
// unique-cell arrays; start/perm give the original point indices on each cell (see PointSet in common.cuh).
struct AcPoints {
  const int* blk;    // [nblocks+1] or null when the set is empty
  const int* cell;
  const int* start;
  const int* perm;
};

// first entry of a per-CTA point list (sorted by cell) whose cell is >= key
__device__ __forceinline__ int ac_lower_bound(const int* __restrict__ cell, int a, int b, int key) {
  while (a < b) {
    const int m = (a + b) >> 1;
    if (cell[m] < key) a = m + 1; else b = m;
  }
  return a;
}
// v + sum of val[perm[m]] * scale over the points on `key` (sequential, original point order); v if none
__device__ __noinline__ double ac_add_points(double v, AcPoints ps, int a, int b, int key,
                                             const double* __restrict__ val, double scale) {
  const int k = ac_lower_bound(ps.cell, a, b, key);
  if (k < b && ps.cell[k] == key)
    for (int m = ps.start[k]; m < ps.start[k + 1]; m++) v += val[ps.perm[m]] * scale;
  return v;
}

// Slab decomposition, fused into the step kernels (all null / zero on a single GPU).  The CTAs whose cells include
// my first (last) owned row are scheduled first (perm), wait until the neighbour's previous step has delivered the
// halo row they read, and -- after their epilogue -- store their piece of the new edge row straight into the
// neighbour's halo row over NVLink and publish it with a system-scope fence + atomic on the neighbour's flag.
// Interior CTAs never wait, so the exchange latency is hidden behind the bulk of the step.
struct AcFuse {
  const int* perm;              // launch order -> logical CTA id (edge CTAs first), or null
  int own0, own_last;           // my first / last owned local row
  int has_lo, has_hi;
  double *lo_u, *hi_u;          // neighbour halo rows (peer pointers) of the array this launch produces
  double *lo_p, *hi_p;          // ... and of its phi / phibar companion
  unsigned long long *sig_lo, *sig_hi;   // neighbour flags to bump (peer pointers)
  unsigned long long* my_flags;          // [3] bumped by rank-1's step kernels, [4] by rank+1's, [2] error
  unsigned long long expect_lo, expect_hi;
  // wide frame-only launches: the injected points (sources forward, receiver residuals adjoint) that sit on the rim
  // ring of the box.  They are added in registers BEFORE the cell's single store -- the concurrent box-pair launch
  // stores the same final value, so neither launch may read-modify-write such a cell
  AcPoints rim;
  // Packed halo rows ("LL": every 8-byte word carries 32 data bits and the 32-bit epoch of the launch that sent it, so
  // the data is its own flag).  The sender stores its edge row into the neighbour's LL row and is done -- no
  // system-scope fence, no flag, nothing on its critical path; the receiving CTA (next fused launch of the neighbour)
  // polls the words of the columns it owns, writes the doubles into the halo row of the array it is about to read
  // and goes on.  One NVLink one-way latency per step instead of store + fence round trip + flag + poll.
  int ll;
  unsigned ep_send, ep_recv;                               // epoch stamped on my rows; epoch to wait for (0: nothing to receive)
  ulonglong2 *tx_lo_u, *tx_lo_p, *tx_hi_u, *tx_hi_p;       // the neighbours' LL rows (peer pointers), one entry per column
  const ulonglong2 *rx_lo_u, *rx_lo_p, *rx_hi_u, *rx_hi_p; // my LL rows, filled by the neighbours' previous fused launch
  double *in_u, *in_p;                                     // the arrays this launch READS with halo rows (unpack targets)
#ifdef ADSEIS_TIMELINE
  unsigned long long* tl;   // [5]: min entry, max entry, max after-wait (edge CTAs), max end of compute, max after signal
#endif
};
#ifdef ADSEIS_TIMELINE
__device__ __forceinline__ unsigned long long tl_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TL_DEV(f, k) do { if ((f).tl && threadIdx.x == 0) { if ((k) == 0) atomicMin((f).tl, tl_now()); atomicMax((f).tl + ((k) == 0 ? 1 : (k) + 1), tl_now()); } } while (0)
#else
#define TL_DEV(f, k) do {} while (0)
#endif

// The helpers below are out of line and take plain values: the step kernels keep their register budget, and the
// parameter block is not copied to local memory (which passing `const AcFuse&` to a real call would force).
struct AcLLRx { const ulonglong2 *lo_u, *lo_p, *hi_u, *hi_p; double *in_u, *in_p; unsigned long long* my_flags; int own0, own_last; unsigned ep; };
struct AcLLTx { ulonglong2 *lo_u, *lo_p, *hi_u, *hi_p; unsigned ep; };
__device__ __forceinline__ AcLLRx ac_ll_rx(const AcFuse& f) {
  return AcLLRx{f.rx_lo_u, f.rx_lo_p, f.rx_hi_u, f.rx_hi_p, f.in_u, f.in_p, f.my_flags, f.own0, f.own_last, f.ep_recv};
}
__device__ __forceinline__ AcLLTx ac_ll_tx(const AcFuse& f) { return AcLLTx{f.tx_lo_u, f.tx_lo_p, f.tx_hi_u, f.tx_hi_p, f.ep_send}; }
// frame cell (li, j) on an edge row: fetch the halo values next to it (the thread that unpacks a value is the thread
// that reads it afterwards)
__device__ __noinline__ void ac_ll_recv_cell(AcLLRx r, bool lo, bool hi, int j, int ld) {
  if (lo) {
    const double a = ll_get(r.lo_u, j, r.ep, r.my_flags), b = ll_get(r.lo_p, j, r.ep, r.my_flags);
    r.in_u[(i64)(r.own0 - 1) * ld + j] = a; r.in_p[(i64)(r.own0 - 1) * ld + j] = b;
  }
  if (hi) {
    const double a = ll_get(r.hi_u, j, r.ep, r.my_flags), b = ll_get(r.hi_p, j, r.ep, r.my_flags);
    r.in_u[(i64)(r.own_last + 1) * ld + j] = a; r.in_p[(i64)(r.own_last + 1) * ld + j] = b;
  }
}
// marching CTA: columns j, j+1 of the halo row(s); the rows are then read by bulk copies (async proxy) of this CTA, so
// the writes are fenced to that proxy here and ordered before the producer by the __syncthreads() that follows the call
__device__ __noinline__ void ac_ll_recv_march(AcLLRx r, int j, bool mine, bool t_lo, bool t_hi, int ld) {
  if (mine) {
    if (t_lo) {
      double2 v;
      v.x = ll_get(r.lo_u, j, r.ep, r.my_flags); v.y = ll_get(r.lo_u, j + 1, r.ep, r.my_flags);
      st2(r.in_u + (i64)(r.own0 - 1) * ld + j, v);
    }
    if (t_hi) {
      double2 v;
      v.x = ll_get(r.hi_u, j, r.ep, r.my_flags); v.y = ll_get(r.hi_u, j + 1, r.ep, r.my_flags);
      st2(r.in_u + (i64)(r.own_last + 1) * ld + j, v);
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
// the matching sends (frame cell: u and its phi companion; marching cells are PML-free: companion = 0)
__device__ __noinline__ void ac_ll_send_cell(AcLLTx x, bool lo, bool hi, int j, double u, double p) {
  if (lo) { ll_put(x.lo_u, j, u, x.ep); ll_put(x.lo_p, j, p, x.ep); }
  if (hi) { ll_put(x.hi_u, j, u, x.ep); ll_put(x.hi_p, j, p, x.ep); }
}

__device__ __forceinline__ void ac_fuse_wait(const AcFuse& f, bool t_lo, bool t_hi) {
  pdl_wait();  // (programmatic dependent launch) everything the previous launch wrote is visible from here on
  if (f.ll || !(t_lo || t_hi)) return;  // CTA-uniform
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

__device__ __forceinline__ void ac_fuse_signal(const AcFuse& f, bool t_lo, bool t_hi) {
  if (f.ll || !(t_lo || t_hi)) return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (t_lo) atomicAdd_system(f.sig_lo, 1ULL);
    if (t_hi) atomicAdd_system(f.sig_hi, 1ULL);
  }
}

// ------------------------------------------------------------------------------------------------------------
// forward, general (PML / ring / pad) cell: literal AcousticOneStepCpu.h:27-44
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ac_fwd_general_cell(const AcGeom& g, int li, int j, const double* __restrict__ w,
                                                    const double* __restrict__ wold, const double* __restrict__ c2,
                                                    const double* __restrict__ phi, const double* __restrict__ psi,
                                                    const double* __restrict__ sigx, const double* __restrict__ tauy,
                                                    double* __restrict__ u, double* __restrict__ phio,
                                                    double* __restrict__ psio, AcPoints rim = AcPoints{},
                                                    int rim_a = 0, int rim_n = 0, const double* __restrict__ rim_val = nullptr,
                                                    double rim_scale = 0.0) {
  const int gi = g.goff + li;
  const i64 IJ = (i64)li * g.ld + j;
  if (j >= g.W) { u[IJ] = 0.0; return; }
  if (gi == 0 || gi == g.H - 1 || j == 0 || j == g.W - 1) {
    u[IJ] = 0.0; phio[IJ] = 0.0; psio[IJ] = 0.0;
    return;
  }
  const i64 IpJ = IJ + g.ld, InJ = IJ - g.ld, IJp = IJ + 1, IJn = IJ - 1;
  const double sg = sigx[gi], ta = tauy[j], c = c2[IJ], dt = g.dt;
  double v = (2 - sg * ta * dt * dt - g.kx2 * c - g.ky2 * c) * w[IJ] +
             c * g.rx * g.rx * (w[IpJ] + w[InJ]) +
             c * g.ry * g.ry * (w[IJp] + w[IJn]) +
             g.px * (phi[IpJ] - phi[InJ]) +
             g.py * (psi[IJp] - psi[IJn]) -
             (1 - (sg + ta) * dt / 2) * wold[IJ];
  v = (v == 0.0) ? v : v / (1 + (sg + ta) / 2 * dt);  // zero numerators would take the divide's slow path
  if (rim_n > 0) v = ac_add_points(v, rim, rim_a, rim_a + rim_n, (int)IJ, rim_val, rim_scale);
  u[IJ] = v;
  phio[IJ] = (1. - dt * sg) * phi[IJ] + div_exact(dt * c * (ta - sg) / 2.0, g.hx, g.rhx) * (w[IpJ] - w[InJ]);
  psio[IJ] = (1. - dt * ta) * psi[IJ] + div_exact(dt * c * (sg - ta) / 2.0, g.hy, g.rhy) * (w[IJp] - w[IJn]);
}

// ------------------------------------------------------------------------------------------------------------
// PropagatorKernel = 0 (src/Core.jl:528-549; MPI twin src/MPIAcoustic.jl:212-246): phi', psi' are driven by the NEW
// wavefield u' (before this step's source injection).  Only frame cells carry phi/psi, and a frame cell recomputes
// u' of its four neighbours from the launch's read-only inputs with the very expression their owner CTA uses -- the
// values are bit-identical to the stored ones, no second pass and no inter-CTA dependency is needed.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ac_uprime(const AcGeom& g, int gi, int j, i64 IJ, const double* __restrict__ w,
                                            const double* __restrict__ wold, const double* __restrict__ c2,
                                            const double* __restrict__ phi, const double* __restrict__ psi,
                                            const double* __restrict__ sigx, const double* __restrict__ tauy) {
  // (gi, j) must be an interior cell: callers clamp ring neighbours onto the centre cell and discard the value, so
  // that the loads of all five evaluations of a frame cell can be issued together (no branches in between)
  const double sg = sigx[gi], ta = tauy[j], c = c2[IJ], dt = g.dt;
  const double v = (2 - sg * ta * dt * dt - g.kx2 * c - g.ky2 * c) * w[IJ] +
                   c * g.rx * g.rx * (w[IJ + g.ld] + w[IJ - g.ld]) +
                   c * g.ry * g.ry * (w[IJ + 1] + w[IJ - 1]) +
                   g.px * (phi[IJ + g.ld] - phi[IJ - g.ld]) +
                   g.py * (psi[IJ + 1] - psi[IJ - 1]) -
                   (1 - (sg + ta) * dt / 2) * wold[IJ];
  return (v == 0.0) ? v : v / (1 + (sg + ta) / 2 * dt);
}

__device__ __forceinline__ void ac_fwd_general_cell_k0(const AcGeom& g, int li, int j, const double* __restrict__ w,
                                                       const double* __restrict__ wold, const double* __restrict__ c2,
                                                       const double* __restrict__ phi, const double* __restrict__ psi,
                                                       const double* __restrict__ sigx, const double* __restrict__ tauy,
                                                       double* __restrict__ u, double* __restrict__ phio,
                                                       double* __restrict__ psio, AcPoints rim = AcPoints{},
                                                    int rim_a = 0, int rim_n = 0, const double* __restrict__ rim_val = nullptr,
                                                    double rim_scale = 0.0) {
  const int gi = g.goff + li;
  const i64 IJ = (i64)li * g.ld + j;
  if (j >= g.W) { u[IJ] = 0.0; return; }
  if (gi == 0 || gi == g.H - 1 || j == 0 || j == g.W - 1) {
    u[IJ] = 0.0; phio[IJ] = 0.0; psio[IJ] = 0.0;
    return;
  }
  const double sg = sigx[gi], ta = tauy[j], c = c2[IJ], dt = g.dt;
  // u' is zero on the ring (scatter_nd onto the interior): ring neighbours are evaluated on the centre cell, then dropped
  const bool vxm = gi - 1 >= 1, vxp = gi + 1 <= g.H - 2, vym = j - 1 >= 1, vyp = j + 1 <= g.W - 2;
  const double uP = ac_uprime(g, gi, j, IJ, w, wold, c2, phi, psi, sigx, tauy);
  const double uxp = ac_uprime(g, vxp ? gi + 1 : gi, j, vxp ? IJ + g.ld : IJ, w, wold, c2, phi, psi, sigx, tauy);
  const double uxm = ac_uprime(g, vxm ? gi - 1 : gi, j, vxm ? IJ - g.ld : IJ, w, wold, c2, phi, psi, sigx, tauy);
  const double uyp = ac_uprime(g, gi, vyp ? j + 1 : j, vyp ? IJ + 1 : IJ, w, wold, c2, phi, psi, sigx, tauy);
  const double uym = ac_uprime(g, gi, vym ? j - 1 : j, vym ? IJ - 1 : IJ, w, wold, c2, phi, psi, sigx, tauy);
  u[IJ] = (rim_n > 0) ? ac_add_points(uP, rim, rim_a, rim_a + rim_n, (int)IJ, rim_val, rim_scale) : uP;
  const double a = div_exact(dt * c * (ta - sg) / 2.0, g.hx, g.rhx);
  const double b = div_exact(dt * c * (sg - ta) / 2.0, g.hy, g.rhy);
  phio[IJ] = (1. - dt * sg) * phi[IJ] + a * ((vxp ? uxp : 0.0) - (vxm ? uxm : 0.0));
  psio[IJ] = (1. - dt * ta) * psi[IJ] + b * ((vyp ? uyp : 0.0) - (vym ? uym : 0.0));
}

// CTA epilogue shared by both kernels: add `scale * val[perm]` into field[cell] for the injected points this CTA
// owns (sequentially per cell, in original point order), then sample field[cell]*scale into out[perm].
__device__ __forceinline__ void ac_cta_epilogue(int bid, double* __restrict__ field, const AcPoints& inj,
                                                const double* __restrict__ inj_val, double inj_scale,
                                                const AcPoints& smp, double* __restrict__ smp_out, double smp_scale) {
  int ia = 0, ib = 0, sa = 0, sb = 0;
  if (inj.blk != nullptr && inj_val != nullptr) { ia = inj.blk[bid]; ib = inj.blk[bid + 1]; }
  if (smp.blk != nullptr && smp_out != nullptr) { sa = smp.blk[bid]; sb = smp.blk[bid + 1]; }
  if (ib == ia && sb == sa) return;  // CTA-uniform
  __syncthreads();                   // all cells of this CTA are written
  for (int k = ia + threadIdx.x; k < ib; k += blockDim.x) {
    const int cell = inj.cell[k];
    double v = field[cell];
    for (int m = inj.start[k]; m < inj.start[k + 1]; m++) v += inj_val[inj.perm[m]] * inj_scale;
    field[cell] = v;
  }
  if (sb > sa) {
    __syncthreads();
    for (int k = sa + threadIdx.x; k < sb; k += blockDim.x) {
      const double v = field[smp.cell[k]] * smp_scale;
      for (int m = smp.start[k]; m < smp.start[k + 1]; m++) smp_out[smp.perm[m]] = v;
    }
  }
}

// frame CTA -> (rect, first linear cell)
__device__ __forceinline__ void ac_frame_locate(const AcTiling& t, int fb, int* rect, int* idx0) {
  int r = 0;
#pragma unroll
  for (int k = 1; k < 4; k++)
    if (k < t.nrect && fb >= t.rblk[k]) r = k;
  *rect = r;
  *idx0 = (fb - t.rblk[r]) * (t.fthr * t.fcpt);
}

// ------------------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------------------
template <int PK, int FO = 0>  // PropagatorKernel: 1 = custom-op scheme (phi, psi from the old wavefield), 0 = TF-op scheme; FO: frame-only launch
__global__ void __launch_bounds__(FO ? AC_FO_THREADS : AC_FWD_THREADS, FO ? 6 : AC_MINB_FWD)
ac_fwd_kernel(AcGeom g, AcTiling t, const double* __restrict__ w, const double* __restrict__ wold,
              const double* __restrict__ c2, const double* __restrict__ phi, const double* __restrict__ psi,
              const double* __restrict__ sigx, const double* __restrict__ tauy, double* __restrict__ u,
              double* __restrict__ phio, double* __restrict__ psio, AcPoints src,
              const double* __restrict__ srcv_row, AcPoints rcv, double* __restrict__ rcvv_row, AcFuse f) {
  pdl_launch_dependents();
  TL_DEV(f, 0);
  const int bid = f.perm ? f.perm[blockIdx.x] : blockIdx.x;
  const int ld = g.ld;
  bool t_lo = false, t_hi = false;  // this CTA owns cells of my first / last owned row next to a neighbour
  int rect = 0, idx0 = 0, wdt = 1, ncell = 0;
  if (FO || bid >= t.nmarch) {
    // ---------------- frame CTA ----------------
    ac_frame_locate(t, bid - t.nmarch, &rect, &idx0);
    wdt = t.rc1[rect] - t.rc0[rect];
    ncell = (t.rr1[rect] - t.rr0[rect]) * wdt;
    const int rlo = t.rr0[rect] + idx0 / wdt, rhi = t.rr0[rect] + (min(ncell, idx0 + t.fthr * t.fcpt) - 1) / wdt;
    t_lo = f.has_lo && rlo <= f.own0 && f.own0 <= rhi;
    t_hi = f.has_hi && rlo <= f.own_last && f.own_last <= rhi;
    ac_fuse_wait(f, t_lo, t_hi);
    if (f.ll && f.ep_recv != 0u && (t_lo || t_hi)) {
      for (int k = 0; k < t.fcpt; k++) {
        const int idx = idx0 + k * t.fthr + threadIdx.x;
        if (idx < ncell && threadIdx.x < t.fthr) {
          const int li = t.rr0[rect] + idx / wdt, j = t.rc0[rect] + idx % wdt;
          const bool lo = t_lo && li == f.own0, hi = t_hi && li == f.own_last;
          if (lo || hi) ac_ll_recv_cell(ac_ll_rx(f), lo, hi, j, ld);
        }
      }
    }
    if (t_lo || t_hi) TL_DEV(f, 1);
    int rim_a = 0, rim_n = 0;
    if (f.rim.blk != nullptr && srcv_row != nullptr) { rim_a = f.rim.blk[bid]; rim_n = f.rim.blk[bid + 1] - rim_a; }
#pragma unroll 4
    for (int k = 0; k < t.fcpt; k++) {
      const int idx = idx0 + k * t.fthr + threadIdx.x;
      if (idx < ncell && threadIdx.x < t.fthr) {
        const int li = t.rr0[rect] + idx / wdt, j = t.rc0[rect] + idx % wdt;
        const bool inbox = rim_n > 0 && li >= t.bx_r0 && li < t.bx_r1 && j >= t.bx_c0 && j < t.bx_c1;
        if constexpr (PK == 0) ac_fwd_general_cell_k0(g, li, j, w, wold, c2, phi, psi, sigx, tauy, u, phio, psio, f.rim, rim_a, inbox ? rim_n : 0, srcv_row, g.dt2);
        else ac_fwd_general_cell(g, li, j, w, wold, c2, phi, psi, sigx, tauy, u, phio, psio, f.rim, rim_a, inbox ? rim_n : 0, srcv_row, g.dt2);
      }
    }
  } else {
    // ---------------- marching CTA ----------------
    const int ct = bid % t.nct, tr = bid / t.nct;
    int r0, r1;
    ac_row_tile(t, tr, &r0, &r1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = t.mc0 + ct * AC_TILE_COLS;
    const int jb = c0 + warp * AC_WCOLS;
    const int j = jb + 2 * lane;
    const bool ldok = j < ld;       // may load
    const bool act = j < t.mc_end;  // computes and stores (both columns j, j+1 are inside the box)
    t_lo = f.has_lo && r0 <= f.own0 && f.own0 < r1;
    t_hi = f.has_hi && r0 <= f.own_last && f.own_last < r1;
    ac_fuse_wait(f, t_lo, t_hi);
    if (f.ll && f.ep_recv != 0u && (t_lo || t_hi)) ac_ll_recv_march(ac_ll_rx(f), j, act && threadIdx.x < AC_THREADS, t_lo, t_hi, ld);
    if (t_lo || t_hi) TL_DEV(f, 1);
    extern __shared__ __align__(128) unsigned char ac_smem[];
    AcFwdStage* stg = reinterpret_cast<AcFwdStage*>(ac_smem);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ac_smem + AC_NST_FWD * sizeof(AcFwdStage));
    unsigned long long* empty = full + AC_NST_FWD;
    const int nrows = r1 - r0;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < AC_NST_FWD; k++) { mbar_init(full + k, 1); mbar_init(empty + k, AC_WARPS); }
      mbar_init_fence();
    }
    __syncthreads();
    if (warp == AC_WARPS) {
      // producer warp: one 4-KB bulk copy per streamed array and row, AC_NST_FWD rows ahead of the consumers
      if (lane == 0) {
        const unsigned hbytes = (unsigned)min(AC_HCOLS, ld - (c0 - 2)) * 8u;  // never read past the end of a row
        const unsigned cbytes = (unsigned)min(AC_TILE_COLS, ld - c0) * 8u;
        for (int it = 0; it < nrows; it++) {
          const int sidx = it % AC_NST_FWD;
          if (it >= AC_NST_FWD) { mbar_wait(empty + sidx, (unsigned)(it / AC_NST_FWD - 1) & 1u); ring_refill_fence(); }
          AcFwdStage& s = stg[sidx];
          const i64 ro = (i64)(r0 + it) * ld + c0;
          mbar_arrive_expect_tx(full + sidx, hbytes + 2u * cbytes);
          bulk_g2s(s.w, w + ro + ld - 2, hbytes, full + sidx);  // li+1 <= Hl-1: the box never contains the last row
          bulk_g2s(s.wold, wold + ro, cbytes, full + sidx);
          bulk_g2s(s.c2, c2 + ro, cbytes, full + sidx);
        }
      }
    } else {
      const bool wact = jb < t.mc_end;  // this warp's strip intersects the box
      const double2 z2 = make_double2(0.0, 0.0);
      const double kx2 = g.kx2, ky2 = g.ky2, rx = g.rx, ry = g.ry;
      double2 wm = z2, wc = z2;
      double we = 0.0;  // warp-edge neighbour of the centre row (lanes 0 and 31)
      if (wact && ldok) {
        wm = ld2(w + (i64)(r0 - 1) * ld + j);  // r0 >= 1: the box never contains row 0
        wc = ld2(w + (i64)r0 * ld + j);
        if (act) {
          if (lane == 0) we = w[(i64)r0 * ld + jb - 1];
          if (lane == 31) we = w[(i64)r0 * ld + jb + AC_WCOLS];
        }
      }
      const int so = 2 + warp * AC_WCOLS + 2 * lane;
      for (int it = 0; it < nrows; it++) {
        const int li = r0 + it, sidx = it % AC_NST_FWD;
        mbar_wait(full + sidx, (unsigned)(it / AC_NST_FWD) & 1u);
        const AcFwdStage& s = stg[sidx];
        double2 wn = z2, wo = z2, cc = z2;
        double we_n = 0.0;
        if (wact) {
          wn = ld2(s.w + so); wo = ld2(s.wold + so - 2); cc = ld2(s.c2 + so - 2);
          if (lane == 0) we_n = s.w[so - 1];
          if (lane == 31) we_n = s.w[so + 2];
        }
#ifdef AC_RING_DEBUG
        if (wact && ldok) {
          const i64 gro = (i64)li * ld + j;
          const double2 gn = ld2(w + gro + ld), go = ld2(wold + gro), gc2 = ld2(c2 + gro);
          const bool bad_w = __double_as_longlong(gn.x) != __double_as_longlong(wn.x) || __double_as_longlong(gn.y) != __double_as_longlong(wn.y);
          const bool bad_o = __double_as_longlong(go.x) != __double_as_longlong(wo.x) || __double_as_longlong(go.y) != __double_as_longlong(wo.y);
          const bool bad_c = __double_as_longlong(gc2.x) != __double_as_longlong(cc.x) || __double_as_longlong(gc2.y) != __double_as_longlong(cc.y);
          if (bad_w || bad_o || bad_c) {
            const unsigned k = atomicAdd(&g_ring_dbg_n, 1u);
            if (k < 64) {
              double* r = g_ring_dbg + k * 12;
              r[0] = bid; r[1] = it; r[2] = nrows; r[3] = warp * 32 + lane; r[4] = li; r[5] = j;
              r[6] = (bad_w ? 1 : 0) + (bad_o ? 2 : 0) + (bad_c ? 4 : 0);
              const double* arr = bad_w ? w + ld : (bad_o ? wold : c2);
              r[7] = bad_w ? wn.x : (bad_o ? wo.x : cc.x);              // what the stage held
              r[8] = arr[gro];                                          // what it should hold
              r[9] = (it >= AC_NST_FWD) ? arr[gro - (i64)AC_NST_FWD * ld] : -1.0;             // the row AC_NST_FWD iterations earlier
              r[10] = (it + AC_NST_FWD < nrows) ? arr[gro + (i64)AC_NST_FWD * ld] : -1.0;      // ... later
              r[11] = (double)clock64();
            }
          }
        }
#endif
        __syncwarp();  // every lane has read the stage
        if (lane == 0) mbar_arrive(empty + sidx);
        if (wact) {
          double lft = __shfl_up_sync(0xffffffffu, wc.y, 1);
          double rgt = __shfl_down_sync(0xffffffffu, wc.x, 1);
          if (lane == 0) lft = we;
          if (lane == 31) rgt = we;
          if (act) {
            double2 o;
            {
              const double c = cc.x;
              o.x = (2 - kx2 * c - ky2 * c) * wc.x + c * rx * rx * (wn.x + wm.x) + c * ry * ry * (wc.y + lft) - wo.x;
            }
            {
              const double c = cc.y;
              o.y = (2 - kx2 * c - ky2 * c) * wc.y + c * rx * rx * (wn.y + wm.y) + c * ry * ry * (rgt + wc.x) - wo.y;
            }
            st2(u + (i64)li * ld + j, o);
          }
          wm = wc;
          wc = wn;
          we = we_n;
        }
      }
    }
  }
  ac_cta_epilogue(bid, u, src, srcv_row, g.dt2, rcv, rcvv_row, 1.0);
  TL_DEV(f, 2);
  if (t_lo || t_hi) {  // push my piece of the new edge row(s) into the neighbours' halo rows, then publish
    __syncthreads();
    if (FO || bid >= t.nmarch) {
#pragma unroll 4
      for (int k = 0; k < t.fcpt; k++) {
        const int idx = idx0 + k * t.fthr + threadIdx.x;
        if (idx < ncell && threadIdx.x < t.fthr) {
          const int li = t.rr0[rect] + idx / wdt, j = t.rc0[rect] + idx % wdt;
          const i64 IJ = (i64)li * ld + j;
          if (f.ll) {
            ac_ll_send_cell(ac_ll_tx(f), t_lo && li == f.own0, t_hi && li == f.own_last, j, u[IJ], phio[IJ]);
          } else {
            if (t_lo && li == f.own0) { f.lo_u[j] = u[IJ]; f.lo_p[j] = phio[IJ]; }
            if (t_hi && li == f.own_last) { f.hi_u[j] = u[IJ]; f.hi_p[j] = phio[IJ]; }
          }
        }
      }
    } else {
      const int ct = bid % t.nct;
      const int j = t.mc0 + ct * AC_TILE_COLS + 2 * threadIdx.x;
      if (j < t.mc_end && threadIdx.x < AC_THREADS) {
        if (f.ll) {
          if (t_lo) { const double2 v = ld2(u + (i64)f.own0 * ld + j); ac_ll_send_cell(ac_ll_tx(f), true, false, j, v.x, 0.0); ac_ll_send_cell(ac_ll_tx(f), true, false, j + 1, v.y, 0.0); }
          if (t_hi) { const double2 v = ld2(u + (i64)f.own_last * ld + j); ac_ll_send_cell(ac_ll_tx(f), false, true, j, v.x, 0.0); ac_ll_send_cell(ac_ll_tx(f), false, true, j + 1, v.y, 0.0); }
        } else {
          if (t_lo) st2(f.lo_u + j, ld2(u + (i64)f.own0 * ld + j));
          if (t_hi) st2(f.hi_u + j, ld2(u + (i64)f.own_last * ld + j));
        }
      }
    }
    ac_fuse_signal(f, t_lo, t_hi);
    TL_DEV(f, 3);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Adjoint arithmetic of a cell whose whole stencil is PML-free (D = 1, no phibar / psibar terms).  The adjoint is
// compared with the reference at 1e-10, not bit for bit, so these use explicit fused multiply-adds (the library is
// compiled with -fmad=false for the forward kernels).  EVERY code path that evaluates such a cell -- the marching CTAs
// of ac_adj_kernel, the frame CTAs' ac_adj_general_cell, ac_adj2_kernel -- goes through these three functions, so the
// result does not depend on which path owns the cell (checkpoint segments shift the pairing of steps).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ac_a0(double c, double kx2, double ky2) {      // 2 - kx2 c - ky2 c
  return __fma_rn(-ky2, c, __fma_rn(-kx2, c, 2.0));
}
// ubar[s-1](P) = a0(P) g(P) + rx2 (cg(P+ex) + cg(P-ex)) + ry2 (cg(P+ey) + cg(P-ey)) - ubar[s+1](P),  cg = c^2 ubar[s]
__device__ __forceinline__ double ac_adj_cell(double a0, double rx2, double ry2, double g, double cgn, double cgm,
                                              double cgr, double cgl, double u2) {
  return __fma_rn(a0, g, __fma_rn(rx2, cgn + cgm, __fma_rn(ry2, cgr + cgl, -u2)));
}
// cbar(P) = ((-kx2 - ky2) w(P) + rx2 (w(P+ex) + w(P-ex)) + ry2 (w(P+ey) + w(P-ey))) g(P),  w = u[s-1]
__device__ __forceinline__ double ac_corr_cell(double kk, double rx2, double ry2, double wc, double wn, double wm,
                                               double wr, double wl, double g) {
  return __fma_rn(kk, wc, __fma_rn(rx2, wn + wm, ry2 * (wr + wl))) * g;
}

// ------------------------------------------------------------------------------------------------------------
// adjoint, general cell (gather form of AcousticOneStepCpu.h:76-124)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ac_adj_general_cell(const AcGeom& g, int li, int j, const double* __restrict__ ub1,
                                                    const double* __restrict__ ub2, const double* __restrict__ wf,
                                                    const double* __restrict__ c2, const double* __restrict__ phib,
                                                    const double* __restrict__ psib,
                                                    const double* __restrict__ sigx, const double* __restrict__ tauy,
                                                    double* __restrict__ ub0, double* __restrict__ phibo,
                                                    double* __restrict__ psibo, double* __restrict__ G,
                                                    bool state_only = false, AcPoints rim = AcPoints{}, int rim_a = 0,
                                                    int rim_n = 0, const double* __restrict__ rim_val = nullptr) {
  // All loads are issued up front from always-valid (clamped) addresses and the validity predicates only select
  // terms afterwards: a frame cell then costs ONE memory round trip instead of one per neighbour branch (the
  // frame is ~2 % of the cells but its dependent DRAM round trips used to form the tail of every launch).
  const int gi = g.goff + li;
  const i64 IJ = (i64)li * g.ld + j;
  if (j >= g.W) { ub0[IJ] = 0.0; return; }
  const double dt = g.dt;
  const bool colok = (j >= 1 && j <= g.W - 2);
  const bool rowok = (gi >= 1 && gi <= g.H - 2);
  const bool intP = rowok && colok;
  const bool vxm = colok && gi - 1 >= 1 && gi - 1 <= g.H - 2;  // Q = P - e_x is an interior cell
  const bool vxp = colok && gi + 1 >= 1 && gi + 1 <= g.H - 2;  // Q = P + e_x
  const bool vym = rowok && j - 1 >= 1 && j - 1 <= g.W - 2;    // Q = P - e_y
  const bool vyp = rowok && j + 1 >= 1 && j + 1 <= g.W - 2;    // Q = P + e_y
  const i64 Qxm = vxm ? IJ - g.ld : IJ, Qxp = vxp ? IJ + g.ld : IJ, Qym = vym ? IJ - 1 : IJ, Qyp = vyp ? IJ + 1 : IJ;
  const i64 Wu = intP ? IJ + g.ld : IJ, Wd = intP ? IJ - g.ld : IJ, Wr = intP ? IJ + 1 : IJ, Wl = intP ? IJ - 1 : IJ;
  const double sg = sigx[gi], sgm = sigx[vxm ? gi - 1 : gi], sgp = sigx[vxp ? gi + 1 : gi];
  const double ta = tauy[j], tam = tauy[vym ? j - 1 : j], tap = tauy[vyp ? j + 1 : j];
  const double cP = c2[IJ], uP = ub1[IJ], u2P = ub2[IJ], pb = phib[IJ], qb = psib[IJ], GP = G[IJ];
  const double cxm = c2[Qxm], uxm = ub1[Qxm], pxm = phib[Qxm];
  const double cxp = c2[Qxp], uxp = ub1[Qxp], pxp = phib[Qxp];
  const double cym = c2[Qym], uym = ub1[Qym], qym = psib[Qym];
  const double cyp = c2[Qyp], uyp = ub1[Qyp], qyp = psib[Qyp];
  const double wC = wf[IJ], wU = wf[Wu], wD = wf[Wd], wR = wf[Wr], wL = wf[Wl];
  // the adjoint is compared with the reference at 1e-10, not bit for bit: D in [1,2) is inverted once per cell
  // (fast path of the reciprocal sequence) and the constant divisors are folded into rhx, rhy
  const double kx = dt * 0.5 * g.rhx, ky = dt * 0.5 * g.rhy;
  // A cell whose whole stencil is PML-free (sigma = tau = 0 on it and on its four interior neighbours) is evaluated
  // with the MARCHING CTAs' expression and summation order, whoever owns it: the frame-only launches of the two-step
  // path own the rim of the box, which a one-step launch marches -- both must give the same bits, or the result would
  // depend on how the steps of a sweep happen to be paired (checkpoint segments shift the pairing).
  if (intP && vxm && vxp && vym && vyp && sg == 0.0 && sgm == 0.0 && sgp == 0.0 && ta == 0.0 && tam == 0.0 && tap == 0.0) {
    const double rx2 = g.rx * g.rx, ry2 = g.ry * g.ry;
    double a_ = ac_adj_cell(ac_a0(cP, g.kx2, g.ky2), rx2, ry2, uP, cxp * uxp, cxm * uxm, cyp * uyp, cym * uym, u2P);
    if (state_only && rim_n > 0) a_ = ac_add_points(a_, rim, rim_a, rim_a + rim_n, (int)IJ, rim_val, 1.0);
    ub0[IJ] = a_;
    if (state_only) return;   // a rim cell of the box recomputed by a wide frame launch: the box kernel accumulates Gbar
    phibo[IJ] = (1. - dt * sg) * pb + g.px * (uxm - uxp);   // never used (its coefficient is tau - sigma = 0); kept finite
    psibo[IJ] = (1. - dt * ta) * qb + g.py * (uym - uyp);
    G[IJ] = GP + ac_corr_cell(-g.kx2 - g.ky2, rx2, ry2, wC, wU, wD, wR, wL, uP);
    return;
  }
  double acc = 0.0, gP = 0.0;
  if (intP) {
    gP = uP * (1.0 / (1 + (sg + ta) / 2 * dt));
    acc = (2 - sg * ta * dt * dt - g.kx2 * cP - g.ky2 * cP) * gP;
  }
  double gxm = 0.0, gxp = 0.0, gym = 0.0, gyp = 0.0;
  if (vxm) {  // grad_w[IpJ] of Q = P - e_x
    gxm = uxm * (1.0 / (1 + (sgm + ta) / 2 * dt));
    acc += cxm * g.rx * g.rx * gxm + cxm * (ta - sgm) * kx * pxm;
  }
  if (vxp) {  // grad_w[InJ] of Q = P + e_x
    gxp = uxp * (1.0 / (1 + (sgp + ta) / 2 * dt));
    acc += cxp * g.rx * g.rx * gxp - cxp * (ta - sgp) * kx * pxp;
  }
  if (vym) {  // grad_w[IJp] of Q = P - e_y
    gym = uym * (1.0 / (1 + (sg + tam) / 2 * dt));
    acc += cym * g.ry * g.ry * gym + cym * (sg - tam) * ky * qym;
  }
  if (vyp) {  // grad_w[IJn] of Q = P + e_y
    gyp = uyp * (1.0 / (1 + (sg + tap) / 2 * dt));
    acc += cyp * g.ry * g.ry * gyp - cyp * (sg - tap) * ky * qyp;
  }
  if (intP) {
    acc += -(1 - (sg + ta) * dt / 2) * (u2P * (1.0 / (1 + (sg + ta) / 2 * dt)));  // grad_wold of step s+1
    phibo[IJ] = (1. - dt * sg) * pb + g.px * (gxm - gxp);
    psibo[IJ] = (1. - dt * ta) * qb + g.py * (gym - gyp);
    const double cb = ((-g.kx2 - g.ky2) * wC + g.rx * g.rx * (wU + wD) + g.ry * g.ry * (wR + wL)) * gP +
                      (ta - sg) * kx * (wU - wD) * pb + (sg - ta) * ky * (wR - wL) * qb;
    G[IJ] = GP + cb;
  }
  ub0[IJ] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// adjoint, general cell, PropagatorKernel = 0 (transpose of Core.jl:528-549 as tf.gradients builds it).  With
//   a_R = dt c_R (tau_R - sigma_R) / (2 hx),  b_R = dt c_R (sigma_R - tau_R) / (2 hy)   (interior R)
// the adjoint of the pre-injection output u'_s is
//   utilde_s[Q] = ubar_s[Q] + a_{Q-ex} phibar_s[Q-ex] - a_{Q+ex} phibar_s[Q+ex] + b_{Q-ey} psibar_s[Q-ey] - b_{Q+ey} psibar_s[Q+ey]
// and g_s = utilde_s / D replaces ubar_s / D everywhere; w receives no phibar / psibar terms; the c-gradient's
// phibar / psibar terms use u'_s (here: the stored post-injection u[s]; the injected part is removed by
// k_ac_k0_src_corr).  utilde_s of the frame cells is kept in a side plane for the wold-term of the next launch
// (inside the marching box utilde == ubar).
// ------------------------------------------------------------------------------------------------------------
struct AcK0 {
  const double* wnew;   // u[s] (post-injection), history slot s
  const double* ut_in;  // utilde_{s+1} (frame cells)
  double* ut_out;       // utilde_s
  const double* ub2;    // ubar_{s+1} (== utilde_{s+1} wherever every cell within two rows / columns is PML-free)
};

__device__ __forceinline__ bool ac_interior(const AcGeom& g, int gi, int j) {
  return gi >= 1 && gi <= g.H - 2 && j >= 1 && j <= g.W - 2;
}

__device__ __forceinline__ double ac_utilde(const AcGeom& g, int gi, int j, i64 Q, const double* __restrict__ ub1,
                                            const double* __restrict__ c2, const double* __restrict__ phib,
                                            const double* __restrict__ psib, const double* __restrict__ sigx,
                                            const double* __restrict__ tauy, double kx, double ky) {
  // (gi, j) must be interior.  Branch-free: ring neighbours are read at Q itself and their term is dropped, so the
  // loads of the five evaluations of a frame cell can be issued together.
  const bool vxm = gi - 1 >= 1, vxp = gi + 1 <= g.H - 2, vym = j - 1 >= 1, vyp = j + 1 <= g.W - 2;
  const i64 Rxm = vxm ? Q - g.ld : Q, Rxp = vxp ? Q + g.ld : Q, Rym = vym ? Q - 1 : Q, Ryp = vyp ? Q + 1 : Q;
  const double sg = sigx[gi], ta = tauy[j];
  const double sgm = sigx[vxm ? gi - 1 : gi], sgp = sigx[vxp ? gi + 1 : gi];
  const double tam = tauy[vym ? j - 1 : j], tap = tauy[vyp ? j + 1 : j];
  const double txm = c2[Rxm] * (ta - sgm) * kx * phib[Rxm], txp = c2[Rxp] * (ta - sgp) * kx * phib[Rxp];
  const double tym = c2[Rym] * (sg - tam) * ky * psib[Rym], typ = c2[Ryp] * (sg - tap) * ky * psib[Ryp];
  double v = ub1[Q];
  v += vxm ? txm : 0.0;
  v -= vxp ? txp : 0.0;
  v += vym ? tym : 0.0;
  v -= vyp ? typ : 0.0;
  return v;
}

__device__ __forceinline__ void ac_adj_general_cell_k0(const AcGeom& g, int li, int j, const double* __restrict__ ub1,
                                                       const double* __restrict__ wf, const double* __restrict__ c2,
                                                       const double* __restrict__ phib, const double* __restrict__ psib,
                                                       const double* __restrict__ sigx, const double* __restrict__ tauy,
                                                       double* __restrict__ ub0, double* __restrict__ phibo,
                                                       double* __restrict__ psibo, double* __restrict__ G, AcK0 k0) {
  const int gi = g.goff + li;
  const i64 IJ = (i64)li * g.ld + j;
  if (j >= g.W) { ub0[IJ] = 0.0; return; }
  const double dt = g.dt;
  const double kx = dt * 0.5 * g.rhx, ky = dt * 0.5 * g.rhy;
  const bool colok = (j >= 1 && j <= g.W - 2), rowok = (gi >= 1 && gi <= g.H - 2);
  const bool intP = rowok && colok;
  const bool vxm = colok && gi - 1 >= 1 && gi - 1 <= g.H - 2;  // Q = P - e_x is an interior cell
  const bool vxp = colok && gi + 1 >= 1 && gi + 1 <= g.H - 2;
  const bool vym = rowok && j - 1 >= 1 && j - 1 <= g.W - 2;
  const bool vyp = rowok && j + 1 >= 1 && j + 1 <= g.W - 2;
  // every evaluation runs on an interior cell: invalid ones are redirected to a valid one and dropped.  A ring cell P
  // has at most one interior neighbour; `safe` is that neighbour (or P itself when P is interior).
  // A cell with nothing but PML-free interior cells within two rows / columns (the region the marching CTAs cover, and a
  // few columns more) has utilde == ubar on itself and its neighbours: it is evaluated with the marching CTAs' expression
  // and summation order, whoever owns it -- a frame CTA of a step launch or the whole-sweep kernel -- so that the bits
  // of the gradient do not depend on the schedule.
  if (gi - 2 >= 1 && gi + 2 <= g.H - 2 && j - 2 >= 1 && j + 2 <= g.W - 2 && sigx[gi - 2] == 0.0 && sigx[gi - 1] == 0.0 &&
      sigx[gi] == 0.0 && sigx[gi + 1] == 0.0 && sigx[gi + 2] == 0.0 && tauy[j - 2] == 0.0 && tauy[j - 1] == 0.0 &&
      tauy[j] == 0.0 && tauy[j + 1] == 0.0 && tauy[j + 2] == 0.0) {
    const double rx2 = g.rx * g.rx, ry2 = g.ry * g.ry, uP = ub1[IJ];
    ub0[IJ] = ac_adj_cell(ac_a0(c2[IJ], g.kx2, g.ky2), rx2, ry2, uP, c2[IJ + g.ld] * ub1[IJ + g.ld], c2[IJ - g.ld] * ub1[IJ - g.ld],
                          c2[IJ + 1] * ub1[IJ + 1], c2[IJ - 1] * ub1[IJ - 1], k0.ub2[IJ]);
    G[IJ] = G[IJ] + ac_corr_cell(-g.kx2 - g.ky2, rx2, ry2, wf[IJ], wf[IJ + g.ld], wf[IJ - g.ld], wf[IJ + 1], wf[IJ - 1], uP);
    return;
  }
  const int sgi = intP ? gi : (vxm ? gi - 1 : (vxp ? gi + 1 : gi)), sj = intP ? j : (vym ? j - 1 : (vyp ? j + 1 : j));
  const bool any = intP || vxm || vxp || vym || vyp;
  if (!any) { ub0[IJ] = 0.0; return; }  // ring corners: no interior neighbour at all
  const i64 S = IJ + (i64)(sgi - gi) * g.ld + (sj - j);
  auto gof = [&](bool ok, int qi, int qj, i64 Q) -> double {
    const int ei = ok ? qi : sgi, ej = ok ? qj : sj;
    const i64 E = ok ? Q : S;
    const double ut = ac_utilde(g, ei, ej, E, ub1, c2, phib, psib, sigx, tauy, kx, ky);
    return ok ? ut * (1.0 / (1 + (sigx[ei] + tauy[ej]) / 2 * dt)) : 0.0;
  };
  const double utP = ac_utilde(g, sgi, sj, S, ub1, c2, phib, psib, sigx, tauy, kx, ky);  // == utilde[P] when intP
  const double gxm = gof(vxm, gi - 1, j, IJ - g.ld), gxp = gof(vxp, gi + 1, j, IJ + g.ld);
  const double gym = gof(vym, gi, j - 1, IJ - 1), gyp = gof(vyp, gi, j + 1, IJ + 1);
  double acc = 0.0;
  acc += vxm ? c2[IJ - g.ld] * g.rx * g.rx * gxm : 0.0;
  acc += vxp ? c2[IJ + g.ld] * g.rx * g.rx * gxp : 0.0;
  acc += vym ? c2[IJ - 1] * g.ry * g.ry * gym : 0.0;
  acc += vyp ? c2[IJ + 1] * g.ry * g.ry * gyp : 0.0;
  if (intP) {
    const double sg = sigx[gi], ta = tauy[j], pb = phib[IJ], qb = psib[IJ];
    const double rD = 1.0 / (1 + (sg + ta) / 2 * dt);
    const double gP = utP * rD;
    k0.ut_out[IJ] = utP;
    acc += (2 - sg * ta * dt * dt - g.kx2 * c2[IJ] - g.ky2 * c2[IJ]) * gP;
    acc += -(1 - (sg + ta) * dt / 2) * (k0.ut_in[IJ] * rD);  // grad_wold of step s+1
    phibo[IJ] = (1. - dt * sg) * pb + g.px * (gxm - gxp);
    psibo[IJ] = (1. - dt * ta) * qb + g.py * (gym - gyp);
    const double* un = k0.wnew;
    const double cb = ((-g.kx2 - g.ky2) * wf[IJ] + g.rx * g.rx * (wf[IJ + g.ld] + wf[IJ - g.ld]) +
                       g.ry * g.ry * (wf[IJ + 1] + wf[IJ - 1])) * gP +
                      (ta - sg) * kx * (un[IJ + g.ld] - un[IJ - g.ld]) * pb + (sg - ta) * ky * (un[IJ + 1] - un[IJ - 1]) * qb;
    G[IJ] = G[IJ] + cb;
  }
  ub0[IJ] = acc;
}

// PropagatorKernel = 0: remove the injected part of u[s] from the c-gradient's phibar / psibar terms (they are
// defined on the pre-injection u'_s).  One thread per unique source cell S with v = sum of its srcv[s-1, .] dt^2:
//   G[S-ex] -= kxc(S-ex) v phibar_s[S-ex],  G[S+ex] += kxc(S+ex) v phibar_s[S+ex],  same in y with psibar_s,
//   kxc(P) = dt (tau_P - sigma_P) / (2 hx), kyc(P) = dt (sigma_P - tau_P) / (2 hy), interior P only.
// Runs between two adjoint launches on the same stream (plain launch: fully ordered); fp64 atomics because several
// sources may share a neighbour.
__global__ void k_ac_k0_src_corr(AcGeom g, const int* __restrict__ cell, const int* __restrict__ start,
                                 const int* __restrict__ perm, int nu, const double* __restrict__ srcv_row,
                                 const double* __restrict__ phib, const double* __restrict__ psib,
                                 const double* __restrict__ sigx, const double* __restrict__ tauy,
                                 double* __restrict__ G, int own0, int own1) {
  // slab plans: the point list holds the sources within one row of my rows; only cells of rows [own0, own1) are mine
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nu) return;
  double v = 0.0;
  for (int m = start[k]; m < start[k + 1]; m++) v += srcv_row[perm[m]] * g.dt2;
  if (v == 0.0) return;
  const int li = cell[k] / g.ld, j = cell[k] % g.ld, gi = g.goff + li;
  const i64 S = cell[k];
  const double kx = g.dt * 0.5 * g.rhx, ky = g.dt * 0.5 * g.rhy;
  const bool mine = li >= own0 && li < own1;
  if (ac_interior(g, gi - 1, j) && li - 1 >= own0 && li - 1 < own1) {
    const double cf = (tauy[j] - sigx[gi - 1]) * kx;
    if (cf != 0.0) atomicAdd(&G[S - g.ld], -(cf * v * phib[S - g.ld]));
  }
  if (ac_interior(g, gi + 1, j) && li + 1 >= own0 && li + 1 < own1) {
    const double cf = (tauy[j] - sigx[gi + 1]) * kx;
    if (cf != 0.0) atomicAdd(&G[S + g.ld], cf * v * phib[S + g.ld]);
  }
  if (ac_interior(g, gi, j - 1) && mine) {
    const double cf = (sigx[gi] - tauy[j - 1]) * ky;
    if (cf != 0.0) atomicAdd(&G[S - 1], -(cf * v * psib[S - 1]));
  }
  if (ac_interior(g, gi, j + 1) && mine) {
    const double cf = (sigx[gi] - tauy[j + 1]) * ky;
    if (cf != 0.0) atomicAdd(&G[S + 1], cf * v * psib[S + 1]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// adjoint kernel: ub0 = ubar[s-1] from ub1 = ubar[s], ub2 = ubar[s+1], wf = u[s-1]
// ------------------------------------------------------------------------------------------------------------
template <int PK, int FO = 0>
__global__ void __launch_bounds__(FO ? AC_FO_THREADS : AC_ADJ_THREADS, FO ? 5 : AC_MINB_ADJ)
ac_adj_kernel(AcGeom g, AcTiling t, const double* __restrict__ ub1, const double* __restrict__ ub2,
              const double* __restrict__ wf, const double* __restrict__ c2, const double* __restrict__ phib,
              const double* __restrict__ psib, const double* __restrict__ sigx, const double* __restrict__ tauy,
              double* __restrict__ ub0, double* __restrict__ phibo, double* __restrict__ psibo,
              double* __restrict__ G, AcPoints rcv, const double* __restrict__ res_row, AcPoints src,
              double* __restrict__ gsrcv_row, AcFuse f, AcK0 k0) {
  pdl_launch_dependents();
  TL_DEV(f, 0);
  const int bid = f.perm ? f.perm[blockIdx.x] : blockIdx.x;
#ifdef AC_DEBUG_SKIP_FRAME   // timing experiments only (wrong results)
  if (bid >= t.nmarch) return;
#endif
#ifdef AC_DEBUG_SKIP_MARCH
  if (bid < t.nmarch) return;
#endif
  const int ld = g.ld;
  bool t_lo = false, t_hi = false;
  int rect = 0, idx0 = 0, wdt = 1, ncell = 0;
  if (FO || bid >= t.nmarch) {
    ac_frame_locate(t, bid - t.nmarch, &rect, &idx0);
#ifdef AC_DEBUG_RECTS
    if (!((AC_DEBUG_RECTS >> rect) & 1)) return;
#endif
    wdt = t.rc1[rect] - t.rc0[rect];
    ncell = (t.rr1[rect] - t.rr0[rect]) * wdt;
    const int rlo = t.rr0[rect] + idx0 / wdt, rhi = t.rr0[rect] + (min(ncell, idx0 + t.fthr * t.fcpt) - 1) / wdt;
    t_lo = f.has_lo && rlo <= f.own0 && f.own0 <= rhi;
    t_hi = f.has_hi && rlo <= f.own_last && f.own_last <= rhi;
    ac_fuse_wait(f, t_lo, t_hi);
    if (f.ll && f.ep_recv != 0u && (t_lo || t_hi)) {
      for (int k = 0; k < t.fcpt; k++) {
        const int idx = idx0 + k * t.fthr + threadIdx.x;
        if (idx < ncell && threadIdx.x < t.fthr) {
          const int li = t.rr0[rect] + idx / wdt, j = t.rc0[rect] + idx % wdt;
          const bool lo = t_lo && li == f.own0, hi = t_hi && li == f.own_last;
          if (lo || hi) ac_ll_recv_cell(ac_ll_rx(f), lo, hi, j, ld);
        }
      }
    }
    if (t_lo || t_hi) TL_DEV(f, 1);
    int rim_a = 0, rim_n = 0;
    if (f.rim.blk != nullptr && res_row != nullptr) { rim_a = f.rim.blk[bid]; rim_n = f.rim.blk[bid + 1] - rim_a; }
#pragma unroll 4
    for (int k = 0; k < t.fcpt; k++) {
      const int idx = idx0 + k * t.fthr + threadIdx.x;
      if (idx < ncell && threadIdx.x < t.fthr) {
        const int li = t.rr0[rect] + idx / wdt, j = t.rc0[rect] + idx % wdt;
        if constexpr (PK == 0) ac_adj_general_cell_k0(g, li, j, ub1, wf, c2, phib, psib, sigx, tauy, ub0, phibo, psibo, G, k0);
        else ac_adj_general_cell(g, li, j, ub1, ub2, wf, c2, phib, psib, sigx, tauy, ub0, phibo, psibo, G,
                                 li >= t.bx_r0 && li < t.bx_r1 && j >= t.bx_c0 && j < t.bx_c1, f.rim, rim_a, rim_n, res_row);
      }
    }
  } else {
    const int ct = bid % t.nct, tr = bid / t.nct;
    int r0, r1;
    ac_row_tile(t, tr, &r0, &r1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = t.mc0 + ct * AC_TILE_COLS;
    const int jb = c0 + warp * AC_WCOLS;
    const int j = jb + 2 * lane;
    const bool ldok = j < ld;
    const bool act = j < t.mc_end;
    t_lo = f.has_lo && r0 <= f.own0 && f.own0 < r1;
    t_hi = f.has_hi && r0 <= f.own_last && f.own_last < r1;
    ac_fuse_wait(f, t_lo, t_hi);
    if (f.ll && f.ep_recv != 0u && (t_lo || t_hi)) ac_ll_recv_march(ac_ll_rx(f), j, act && threadIdx.x < AC_THREADS, t_lo, t_hi, ld);
    if (t_lo || t_hi) TL_DEV(f, 1);
    extern __shared__ __align__(128) unsigned char ac_smem[];
    AcAdjStage* stg = reinterpret_cast<AcAdjStage*>(ac_smem);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ac_smem + AC_NST_ADJ * sizeof(AcAdjStage));
    unsigned long long* empty = full + AC_NST_ADJ;
    const int nrows = r1 - r0;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < AC_NST_ADJ; k++) { mbar_init(full + k, 1); mbar_init(empty + k, AC_WARPS); }
      mbar_init_fence();
    }
    __syncthreads();
    if (warp == AC_WARPS) {
      // ---------------- producer warp: keeps AC_NST_ADJ rows of the five streamed arrays in flight ----------------
      if (lane == 0) {
        const unsigned hbytes = (unsigned)min(AC_HCOLS, ld - (c0 - 2)) * 8u;  // never read past the end of a row
        const unsigned cbytes = (unsigned)min(AC_TILE_COLS, ld - c0) * 8u;
        for (int it = 0; it < nrows; it++) {
          const int sidx = it % AC_NST_ADJ;
          if (it >= AC_NST_ADJ) { mbar_wait(empty + sidx, (unsigned)(it / AC_NST_ADJ - 1) & 1u); ring_refill_fence(); }
          AcAdjStage& s = stg[sidx];
          const i64 ro = (i64)(r0 + it) * ld + c0;
          mbar_arrive_expect_tx(full + sidx, 3u * hbytes + 2u * cbytes);
          bulk_g2s(s.ub1, ub1 + ro + ld - 2, hbytes, full + sidx);
          bulk_g2s(s.c2, c2 + ro + ld - 2, hbytes, full + sidx);
          bulk_g2s(s.wf, wf + ro + ld - 2, hbytes, full + sidx);
          bulk_g2s(s.ub2, ub2 + ro, cbytes, full + sidx);
          bulk_g2s(s.G, G + ro, cbytes, full + sidx);
        }
      }
    } else {
      // ---------------- consumer warps: 64 columns each, 3-row windows in registers ----------------
      const bool wact = jb < t.mc_end;  // this warp's strip intersects the box
      const double2 z2 = make_double2(0.0, 0.0);
      const double rx2 = g.rx * g.rx, ry2 = g.ry * g.ry, kk = -g.kx2 - g.ky2, kx2 = g.kx2, ky2 = g.ky2;
      // register windows: cg = c^2 * ubar[s] (D == 1 and no ring cell within one cell of the box), w = u[s-1]
      double2 cgm = z2, cgc = z2, wm = z2, wc = z2, gc = z2, ccen = z2;
      double ecg = 0.0, ew = 0.0;  // warp-edge neighbours of the centre row (lanes 0 and 31)
      if (wact && ldok) {
        i64 ro = (i64)(r0 - 1) * ld + j;
        const double2 c_ = ld2(c2 + ro), u_ = ld2(ub1 + ro);
        cgm = make_double2(c_.x * u_.x, c_.y * u_.y);
        wm = ld2(wf + ro);
        ro += ld;
        ccen = ld2(c2 + ro);
        gc = ld2(ub1 + ro);
        cgc = make_double2(ccen.x * gc.x, ccen.y * gc.y);
        wc = ld2(wf + ro);
        if (act) {
          const i64 rc = (i64)r0 * ld;
          if (lane == 0) { ecg = c2[rc + jb - 1] * ub1[rc + jb - 1]; ew = wf[rc + jb - 1]; }
          if (lane == 31) { ecg = c2[rc + jb + AC_WCOLS] * ub1[rc + jb + AC_WCOLS]; ew = wf[rc + jb + AC_WCOLS]; }
        }
      }
      const int so = 2 + warp * AC_WCOLS + 2 * lane;  // this lane's column pair inside a halo'd stage row
      for (int it = 0; it < nrows; it++) {
        const int li = r0 + it, sidx = it % AC_NST_ADJ;
        mbar_wait(full + sidx, (unsigned)(it / AC_NST_ADJ) & 1u);
        const AcAdjStage& s = stg[sidx];
        double2 un = z2, cn = z2, wn = z2, u2 = z2, Gr = z2;
        double ecg_n = 0.0, ew_n = 0.0;  // edge neighbours of the next centre row
        if (wact) {
          un = ld2(s.ub1 + so); cn = ld2(s.c2 + so); wn = ld2(s.wf + so);
          u2 = ld2(s.ub2 + so - 2); Gr = ld2(s.G + so - 2);
          if (lane == 0) { ecg_n = s.c2[so - 1] * s.ub1[so - 1]; ew_n = s.wf[so - 1]; }
          if (lane == 31) { ecg_n = s.c2[so + 2] * s.ub1[so + 2]; ew_n = s.wf[so + 2]; }
        }
        __syncwarp();  // every lane has read the stage
        if (lane == 0) mbar_arrive(empty + sidx);
        if (wact) {
          const double2 cgp = make_double2(cn.x * un.x, cn.y * un.y);
          double cgl = __shfl_up_sync(0xffffffffu, cgc.y, 1);
          double cgr = __shfl_down_sync(0xffffffffu, cgc.x, 1);
          double wl = __shfl_up_sync(0xffffffffu, wc.y, 1);
          double wr = __shfl_down_sync(0xffffffffu, wc.x, 1);
          if (lane == 0) { cgl = ecg; wl = ew; }
          if (lane == 31) { cgr = ecg; wr = ew; }
          if (act) {
            double2 o, Go;
            o.x = ac_adj_cell(ac_a0(ccen.x, kx2, ky2), rx2, ry2, gc.x, cgp.x, cgm.x, cgc.y, cgl, u2.x);
            o.y = ac_adj_cell(ac_a0(ccen.y, kx2, ky2), rx2, ry2, gc.y, cgp.y, cgm.y, cgr, cgc.x, u2.y);
            Go.x = Gr.x + ac_corr_cell(kk, rx2, ry2, wc.x, wn.x, wm.x, wc.y, wl, gc.x);
            Go.y = Gr.y + ac_corr_cell(kk, rx2, ry2, wc.y, wn.y, wm.y, wr, wc.x, gc.y);
            st2(ub0 + (i64)li * ld + j, o);
            st2(G + (i64)li * ld + j, Go);
          }
          cgm = cgc; cgc = cgp;
          wm = wc; wc = wn;
          gc = un; ccen = cn;
          ecg = ecg_n; ew = ew_n;
        }
      }
    }
  }
  ac_cta_epilogue(bid, ub0, rcv, res_row, 1.0, src, gsrcv_row, g.dt2);
  TL_DEV(f, 2);
  if (t_lo || t_hi) {
    __syncthreads();
    if (FO || bid >= t.nmarch) {
#pragma unroll 4
      for (int k = 0; k < t.fcpt; k++) {
        const int idx = idx0 + k * t.fthr + threadIdx.x;
        if (idx < ncell && threadIdx.x < t.fthr) {
          const int li = t.rr0[rect] + idx / wdt, j = t.rc0[rect] + idx % wdt;
          const i64 IJ = (i64)li * ld + j;
          if (f.ll) {
            ac_ll_send_cell(ac_ll_tx(f), t_lo && li == f.own0, t_hi && li == f.own_last, j, ub0[IJ], phibo[IJ]);
          } else {
            if (t_lo && li == f.own0) { f.lo_u[j] = ub0[IJ]; f.lo_p[j] = phibo[IJ]; }
            if (t_hi && li == f.own_last) { f.hi_u[j] = ub0[IJ]; f.hi_p[j] = phibo[IJ]; }
          }
        }
      }
    } else {
      const int ct = bid % t.nct;
      const int j = t.mc0 + ct * AC_TILE_COLS + 2 * threadIdx.x;
      if (j < t.mc_end && threadIdx.x < AC_THREADS) {
        if (f.ll) {
          if (t_lo) { const double2 v = ld2(ub0 + (i64)f.own0 * ld + j); ac_ll_send_cell(ac_ll_tx(f), true, false, j, v.x, 0.0); ac_ll_send_cell(ac_ll_tx(f), true, false, j + 1, v.y, 0.0); }
          if (t_hi) { const double2 v = ld2(ub0 + (i64)f.own_last * ld + j); ac_ll_send_cell(ac_ll_tx(f), false, true, j, v.x, 0.0); ac_ll_send_cell(ac_ll_tx(f), false, true, j + 1, v.y, 0.0); }
        } else {
          if (t_lo) st2(f.lo_u + j, ld2(ub0 + (i64)f.own0 * ld + j));
          if (t_hi) st2(f.hi_u + j, ld2(ub0 + (i64)f.own_last * ld + j));
        }
      }
    }
    ac_fuse_signal(f, t_lo, t_hi);
    TL_DEV(f, 3);
  }
}

// ============================================================================================================
// WHOLE SWEEP in one launch (small grids).  When a step is a few hundred thousand cells its kernel is all launch latency
// (C1, 401 x 133: 6.5 / 10 us per forward / adjoint launch for 0.3 us of HBM time).  Here one cooperative launch runs
// every step of a sweep: all cells are evaluated as general cells by CTAs that stay resident, own the same cells for
// the whole sweep (so the per-CTA point lists work unchanged) and meet at a grid barrier after every step.  Same
// expressions as the step kernels (interior cells of the adjoint go through the shared FMA helpers), so the results
// are the same bits.  State lives in the same arrays as for the step kernels; between steps it is L2 traffic.
// ============================================================================================================
#define AC_PS_THREADS 256
struct AcPersist {
  double* hist; i64 plane;                 // forward history: slot k at hist + k * plane (single segment, base 0)
  double *phi[2], *psi[2];
  const double *c2, *sigx, *tauy;
  AcPoints src, rcv;
  const double* srcv; int nsrc;
  double* rcvv; int nrcv;                  // forward: traces out (null: no sampling)
  i64 s_first, s_last;
  // adjoint
  double* ub[4]; int nub;
  double *phib[2], *psib[2], *G;
  double* ut[2];                           // PropagatorKernel = 0: utilde side planes
  const double* res; double* gradsrcv;
  unsigned long long* bar;                 // grid barrier counter (zero at launch)
  // Neighbour synchronisation instead of the grid barrier: prog[b] = steps CTA b has finished.  A CTA's cells are a
  // contiguous range of the row-major cell list, and a step reads at most `reach` rows / columns around a cell, so CTA b
  // depends on the CTAs covering [first - reach*W, last + reach*W] only.  The same wait covers the write-after-read
  // hazard of the ping-pong arrays (my dependents are my dependencies).
  unsigned long long* prog;                // null: grid barrier
  int dep_lo, dep_hi;                      // CTA b waits for CTAs b - dep_lo .. b + dep_hi (clipped to the grid)
};

// all CTAs of the (cooperative) launch; `target` = arrivals expected so far
__device__ __forceinline__ void ac_grid_barrier(unsigned long long* bar, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1ULL);
    unsigned long long v;
    do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory"); } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// step `done` of this CTA is finished (all its stores issued): publish, then wait until the CTAs it depends on have
// finished the same step
__device__ __forceinline__ void ac_neighbour_sync(const AcPersist& a, unsigned long long done) {
  __syncthreads();
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(a.prog + b), "l"(done) : "memory");
  }
  const int n = a.dep_lo + a.dep_hi + 1;
  if ((int)threadIdx.x < n) {
    const int q = b - a.dep_lo + (int)threadIdx.x;
    if (q >= 0 && q < (int)gridDim.x && q != b) {
      unsigned long long v;
      do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.prog + q) : "memory"); } while (v < done);
    }
  }
  __syncthreads();
}

template <int PK>
__global__ void __launch_bounds__(AC_PS_THREADS, 2) ac_fwd_persist_kernel(AcGeom g, AcTiling t, AcPersist a) {
  const int bid = blockIdx.x;
  int rect, idx0;
  ac_frame_locate(t, bid, &rect, &idx0);
  const int wdt = t.rc1[rect] - t.rc0[rect];
  const int ncell = (t.rr1[rect] - t.rr0[rect]) * wdt;
  AcPoints none{};
  unsigned long long arrivals = 0;
  for (i64 s = a.s_first; s <= a.s_last; s++) {
    // plain (non-__restrict__) views: these arrays are written by other CTAs during this launch
    double* w = a.hist + (s - 1) * a.plane;
    double* wold = a.hist + (s - 2) * a.plane;
    double* u = a.hist + s * a.plane;
    double *phi = a.phi[(s - 1) & 1], *psi = a.psi[(s - 1) & 1], *phio = a.phi[s & 1], *psio = a.psi[s & 1];
    for (int k = 0; k < t.fcpt; k++) {
      const int idx = idx0 + k * t.fthr + threadIdx.x;
      if (idx < ncell) {
        const int li = t.rr0[rect] + idx / wdt, j = t.rc0[rect] + idx % wdt;
        if constexpr (PK == 0) ac_fwd_general_cell_k0(g, li, j, w, wold, a.c2, phi, psi, a.sigx, a.tauy, u, phio, psio);
        else ac_fwd_general_cell(g, li, j, w, wold, a.c2, phi, psi, a.sigx, a.tauy, u, phio, psio);
      }
    }
    ac_cta_epilogue(bid, u, a.src, a.nsrc > 0 ? a.srcv + (s - 1) * a.nsrc : nullptr, g.dt2, a.rcvv ? a.rcv : none,
                    (a.rcvv && a.nrcv > 0) ? a.rcvv + s * a.nrcv : nullptr, 1.0);
    arrivals += gridDim.x;
    if (a.prog) ac_neighbour_sync(a, arrivals / gridDim.x);
    else ac_grid_barrier(a.bar, arrivals);
  }
}

template <int PK>
__global__ void __launch_bounds__(AC_PS_THREADS, 2) ac_adj_persist_kernel(AcGeom g, AcTiling t, AcPersist a) {
  const int bid = blockIdx.x;
  int rect, idx0;
  ac_frame_locate(t, bid, &rect, &idx0);
  const int wdt = t.rc1[rect] - t.rc0[rect];
  const int ncell = (t.rr1[rect] - t.rr0[rect]) * wdt;
  AcPoints none{};
  const int NUB = a.nub;
  unsigned long long arrivals = 0;
  for (i64 s = a.s_last; s >= a.s_first; s--) {   // ubar[s-1] from ubar[s], ubar[s+1], u[s-1]
    double *ub1 = a.ub[s % NUB], *ub2 = a.ub[(s + 1) % NUB], *ub0 = a.ub[(s - 1 + NUB) % NUB];
    double* wf = a.hist + (s - 1) * a.plane;
    double *phib = a.phib[s & 1], *psib = a.psib[s & 1], *phibo = a.phib[(s - 1) & 1], *psibo = a.psib[(s - 1) & 1];
    AcK0 k0{};
    if constexpr (PK == 0) { k0.wnew = a.hist + s * a.plane; k0.ut_in = a.ut[(s + 1) & 1]; k0.ut_out = a.ut[s & 1]; k0.ub2 = ub2; }
    for (int k = 0; k < t.fcpt; k++) {
      const int idx = idx0 + k * t.fthr + threadIdx.x;
      if (idx < ncell) {
        const int li = t.rr0[rect] + idx / wdt, j = t.rc0[rect] + idx % wdt;
        if constexpr (PK == 0) ac_adj_general_cell_k0(g, li, j, ub1, wf, a.c2, phib, psib, a.sigx, a.tauy, ub0, phibo, psibo, a.G, k0);
        else ac_adj_general_cell(g, li, j, ub1, ub2, wf, a.c2, phib, psib, a.sigx, a.tauy, ub0, phibo, psibo, a.G);
      }
    }
    ac_cta_epilogue(bid, ub0, a.rcv, a.nrcv > 0 ? a.res + (s - 1) * a.nrcv : nullptr, 1.0, (s - 2 >= 1) ? a.src : none,
                    (s - 2 >= 1 && a.nsrc > 0) ? a.gradsrcv + (s - 2) * a.nsrc : nullptr, g.dt2);
    arrivals += gridDim.x;
    if (a.prog) ac_neighbour_sync(a, arrivals / gridDim.x);
    else ac_grid_barrier(a.bar, arrivals);
  }
}

// ============================================================================================================
// TWO time steps per launch (temporal blocking), marching CTAs only.
//
// A marching CTA already holds the rows it needs in shared memory / registers, so it can apply the update twice
// before anything goes back to HBM: u[s+1] = Step(u[s], u[s-1]) is computed on the tile plus a one-cell rim, kept in
// a three-row register window, and u[s+2] = Step(u[s+1], u[s]) follows one row behind.  HBM traffic per cell and
// pair of steps: read u[s], u[s-1], c^2, write u[s+1], u[s+2] = 40 B, i.e. 20 B per cell-step instead of 32 B (both
// snapshots are still written: the reverse sweep needs every time level).  The rim values are recomputed by each
// tile with the same expression (bit-identical to the owner's), so tiles stay independent: no inter-CTA exchange.
// The marched box is the PML-free box shrunk by TWO cells (t2), so that every rim cell is itself a plain interior
// cell (sigma = tau = 0, phi = psi = 0): the launch depends on time levels s and s-1 only.  Everything outside the
// box (absorbing frame, ring, rim) is advanced one step at a time by frame-only launches of ac_fwd_kernel.
// Sources inside (tile + rim) are injected into the register copy of u[s+1] before it is used; injection into
// u[s+2] and receiver sampling of both levels happen in the CTA epilogue as in the one-step kernel.
// Lane layout as in ac_fwd_kernel (8 consumer warps x 64 columns, one double2 per lane); lanes 0 and 31 also carry
// the warp's rim column of u[s+1], computed from a 3-row window of that column.
// ============================================================================================================
#ifndef AC_NST_FWD2
#define AC_NST_FWD2 6
#endif
struct __align__(128) AcFwd2Stage {
  double w[AC_HPAD];      // u[s]   row q+1, columns c0-2 .. c0+513
  double wold[AC_HPAD];   // u[s-1] row q
  double c2[AC_HPAD];     // c^2    row q
};
#define AC_FWD2_SMEM ((int)(AC_NST_FWD2 * (sizeof(AcFwd2Stage) + 16)))

__device__ __forceinline__ double ac_int_cell(double c, double kx2, double ky2, double rx, double ry, double wc,
                                              double wn, double wm, double wr, double wl, double wo) {
  // the marching CTAs' expression of ac_fwd_kernel, term for term (reference order: AcousticOneStepCpu.h:35-44 with
  // sigma = tau = 0, phi = psi = 0)
  return (2 - kx2 * c - ky2 * c) * wc + c * rx * rx * (wn + wm) + c * ry * ry * (wr + wl) - wo;
}

__global__ void __launch_bounds__(AC_FWD_THREADS, 2)
ac_fwd2_kernel(AcGeom g, AcTiling t, const double* __restrict__ w, const double* __restrict__ wold,
               const double* __restrict__ c2, double* __restrict__ u1, double* __restrict__ u2, AcPoints srch,
               const double* __restrict__ srcv_row1, AcPoints src, const double* __restrict__ srcv_row2, AcPoints rcv,
               double* __restrict__ rcvv_row1, double* __restrict__ rcvv_row2
#ifdef ADSEIS_TIMELINE
               , unsigned long long* tl
#endif
               ) {
#ifdef ADSEIS_TIMELINE
  struct TlExit { unsigned long long* p; __device__ ~TlExit() { if (p && threadIdx.x == 0) atomicMax(p + 3, tl_now()); } } tl_exit{tl};
  if (tl && threadIdx.x == 0) { atomicMin(tl, tl_now()); atomicMax(tl + 1, tl_now()); }
#endif
  pdl_launch_dependents();
  const int bid = blockIdx.x;
  const int ld = g.ld;
  const int ct = bid % t.nct, tr = bid / t.nct;
  int r0, r1;
  ac_row_tile(t, tr, &r0, &r1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = t.mc0 + ct * AC_TILE_COLS;
  const int jb = c0 + warp * AC_WCOLS;
  const int j = jb + 2 * lane;
  const bool act = j < t.mc_end;  // stores (both columns j, j+1 are inside the box)
  pdl_wait();
  extern __shared__ __align__(128) unsigned char ac_smem[];
  AcFwd2Stage* stg = reinterpret_cast<AcFwd2Stage*>(ac_smem);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ac_smem + AC_NST_FWD2 * sizeof(AcFwd2Stage));
  unsigned long long* empty = full + AC_NST_FWD2;
  const int nrows = r1 - r0;
  const int nit = nrows + 2;  // intermediate rows q = r0-1 .. r1
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < AC_NST_FWD2; k++) { mbar_init(full + k, 1); mbar_init(empty + k, AC_WARPS); }
    mbar_init_fence();
  }
  __syncthreads();
  if (warp == AC_WARPS) {
    if (lane == 0) {
      const unsigned hbytes = (unsigned)min(AC_HCOLS, ld - (c0 - 2)) * 8u;  // never read past the end of a row
      for (int it = 0; it < nit; it++) {
        const int sidx = it % AC_NST_FWD2;
        if (it >= AC_NST_FWD2) { mbar_wait(empty + sidx, (unsigned)(it / AC_NST_FWD2 - 1) & 1u); ring_refill_fence(); }
        AcFwd2Stage& s = stg[sidx];
        const i64 ro = (i64)(r0 - 1 + it) * ld + c0 - 2;  // row q, first staged column
        mbar_arrive_expect_tx(full + sidx, 3u * hbytes);
        bulk_g2s(s.w, w + ro + ld, hbytes, full + sidx);   // q+1 <= r1+1 <= Hl-1: the box is shrunk by two
        bulk_g2s(s.wold, wold + ro, hbytes, full + sidx);
        bulk_g2s(s.c2, c2 + ro, hbytes, full + sidx);
      }
    }
  } else {
    const bool wact = jb < t.mc_end;
    const double2 z2 = make_double2(0.0, 0.0);
    const double kx2 = g.kx2, ky2 = g.ky2, rx = g.rx, ry = g.ry;
    // u[s] window (rows q-1, q), u[s+1] window (rows q-2, q-1), c^2 of row q-1
    double2 wm = z2, wc = z2, Im = z2, Ic = z2, cprev = z2;
    // rim column of this warp (lane 0: column jb-1, lane 31: column jb+64): u[s] window, outer neighbour, u[s+1]
    double ewm = 0.0, ewc = 0.0, eoc = 0.0, Iec = 0.0;
    const bool edge = (lane == 0) || (lane == 31);
    const int ecol = (lane == 0) ? jb - 1 : jb + AC_WCOLS;      // rim column
    const int ocol = (lane == 0) ? jb - 2 : jb + AC_WCOLS + 1;  // its outer neighbour
    if (wact) {
      const i64 ra = (i64)(r0 - 2) * ld, rb_ = ra + ld;  // r0 >= 2
      if (j + 1 < ld) { wm = ld2(w + ra + j); wc = ld2(w + rb_ + j); }
      if (edge && ocol < ld) { ewm = w[ra + ecol]; ewc = w[rb_ + ecol]; eoc = w[rb_ + ocol]; }
    }
    int ha = 0, hb = 0;
    if (srch.blk != nullptr && srcv_row1 != nullptr) { ha = srch.blk[bid]; hb = srch.blk[bid + 1]; }
    const int so = 2 + warp * AC_WCOLS + 2 * lane;              // own pair inside a staged row
    const int se = (lane == 0) ? so - 1 : so + 2, sx = (lane == 0) ? so - 2 : so + 3;
    for (int it = 0; it < nit; it++) {
      const int q = r0 - 1 + it, sidx = it % AC_NST_FWD2;
      mbar_wait(full + sidx, (unsigned)(it / AC_NST_FWD2) & 1u);
      const AcFwd2Stage& s = stg[sidx];
      double2 wn = z2, wo = z2, cc = z2;
      double ewn = 0.0, eon = 0.0, ewo = 0.0, ecc = 0.0;
      if (wact) {
        wn = ld2(s.w + so); wo = ld2(s.wold + so); cc = ld2(s.c2 + so);
        if (edge) { ewn = s.w[se]; eon = s.w[sx]; ewo = s.wold[se]; ecc = s.c2[se]; }
      }
      __syncwarp();  // every lane has read the stage
      if (lane == 0) mbar_arrive(empty + sidx);
      if (wact) {
        // ---- first step: u[s+1] on row q (own pair and the rim column)
        double lft = __shfl_up_sync(0xffffffffu, wc.y, 1);
        double rgt = __shfl_down_sync(0xffffffffu, wc.x, 1);
        if (lane == 0) lft = ewc;
        if (lane == 31) rgt = ewc;
        double2 In;
        In.x = ac_int_cell(cc.x, kx2, ky2, rx, ry, wc.x, wn.x, wm.x, wc.y, lft, wo.x);
        In.y = ac_int_cell(cc.y, kx2, ky2, rx, ry, wc.y, wn.y, wm.y, rgt, wc.x, wo.y);
        double Ien = 0.0;
        if (edge)
          Ien = (lane == 0) ? ac_int_cell(ecc, kx2, ky2, rx, ry, ewc, ewn, ewm, wc.x, eoc, ewo)
                            : ac_int_cell(ecc, kx2, ky2, rx, ry, ewc, ewn, ewm, eoc, wc.y, ewo);
        // sources on (tile + rim) cells of row q: u[s+1][cell] += srcv * dt^2, sequentially per cell (ScatterAddOps)
        for (int k = ha; k < hb; k++) {
          const int cell = srch.cell[k];
          const int row = cell / ld;
          if (row != q) continue;
          const int col = cell - row * ld;
          const bool mx = col == j, my = col == j + 1, me = edge && col == ecol;
          if (mx || my || me) {
            double v = mx ? In.x : (my ? In.y : Ien);
            for (int m = srch.start[k]; m < srch.start[k + 1]; m++) v += srcv_row1[srch.perm[m]] * g.dt2;
            if (mx) In.x = v; else if (my) In.y = v; else Ien = v;
          }
        }
        if (act && it >= 1 && it <= nrows) st2(u1 + (i64)q * ld + j, In);
        // ---- second step: u[s+2] on row q-1 from u[s+1] rows q-2, q-1, q; "wold" is u[s] row q-1
        if (it >= 2) {
          double l2 = __shfl_up_sync(0xffffffffu, Ic.y, 1);
          double r2 = __shfl_down_sync(0xffffffffu, Ic.x, 1);
          if (lane == 0) l2 = Iec;
          if (lane == 31) r2 = Iec;
          if (act) {
            double2 o;
            o.x = ac_int_cell(cprev.x, kx2, ky2, rx, ry, Ic.x, In.x, Im.x, Ic.y, l2, wm.x);
            o.y = ac_int_cell(cprev.y, kx2, ky2, rx, ry, Ic.y, In.y, Im.y, r2, Ic.x, wm.y);
            st2(u2 + (i64)(q - 1) * ld + j, o);
          }
        }
        Im = Ic; Ic = In; Iec = Ien;
        wm = wc; wc = wn; cprev = cc;
        ewm = ewc; ewc = ewn; eoc = eon;
      }
    }
  }
  const AcPoints none{};
  ac_cta_epilogue(bid, u1, none, nullptr, 0.0, rcv, rcvv_row1, 1.0);          // u[s+1] already carries its sources
  ac_cta_epilogue(bid, u2, src, srcv_row2, g.dt2, rcv, rcvv_row2, 1.0);
}

// ------------------------------------------------------------------------------------------------------------
// Two ADJOINT steps per launch (marching CTAs only), the transpose of ac_fwd2_kernel's pair in reverse order:
//   A = ubar[s-1] = Step^T(ubar[s], ubar[s+1])  on the tile and its one-cell rim (+ the receiver residuals of
//       slot s-1 that fall on those cells, injected into the register copy),
//   B = ubar[s-2] = Step^T(A, ubar[s])          one row behind,
//   Gbar += stencil(u[s-1]) * ubar[s]  +  stencil(u[s-2]) * A     (the same two additions, in the same order, as two
//       one-step launches: bit-identical to them).
// HBM traffic per cell and pair: read ubar[s], ubar[s+1], c^2, u[s-1], u[s-2], Gbar; write A, B, Gbar = 72 B, i.e. 36 B
// per cell-step instead of 56 B.  Four rotating ubar planes (A and B must not overwrite the inputs other tiles read).
// Mapping: the pair carries seven three-row windows, ~170 live registers with the double2-per-lane layout of the
// other kernels -- one CTA of 8 warps per SM then runs at 0.2 IPC (measured: 428 us per launch), and two CTAs spill.
// So here a lane owns ONE column: 16 consumer warps x 32 columns (+ the producer warp) = 544 threads, one CTA per SM,
// <= 120 registers, four consumer warps per scheduler; lanes 0 / 31 also carry the warp's rim columns.
// ------------------------------------------------------------------------------------------------------------
#define AC2_WARPS 16
#define AC2_WCOLS 32
#define AC2_THREADS ((AC2_WARPS + 1) * 32)
#ifndef AC_NST_ADJ2
#define AC_NST_ADJ2 8
#endif
struct __align__(128) AcAdj2Stage {
  double ub1[AC_HPAD];   // ubar[s]   row q+1, columns c0-2 .. c0+513
  double c2[AC_HPAD];    // c^2       row q+1
  double ub2[AC_HPAD];   // ubar[s+1] row q
  double w1[AC_HPAD];    // u[s-1]    row q+1
  double w2[AC_HPAD];    // u[s-2]    row q
  double G[AC_HPAD];     // Gbar      row q-1
};
#define AC_ADJ2_SMEM ((int)(AC_NST_ADJ2 * (sizeof(AcAdj2Stage) + 16)))


// Shared-memory layout after the ring stages: one private row pair per consumer warp holding c^2 * A of its 32 columns
// and its two rim columns (double-buffered: row q is written while row q-1 is read).
#define AC2_PRIV (2 * (AC2_WCOLS + 2))
#undef AC_ADJ2_SMEM
#define AC_ADJ2_SMEM ((int)(AC_NST_ADJ2 * (sizeof(AcAdj2Stage) + 16) + AC2_WARPS * AC2_PRIV * sizeof(double)))

// Stage t of a tile carries row r0-2+t of ubar[s], c^2, u[s-1], row r0-3+t of ubar[s+1], u[s-2] and row r0-4+t of Gbar.
// Iteration `it` (row q = r0-1+it of A, row q-1 of B) reads stages it+2 (rows q+1 / q / q-1), it+1 and it, and releases
// stage `it` at its end: the three-row stencil windows live in the ring, not in registers (the register-window version
// of this kernel spent 28 % of its instructions rotating windows and spilled; the ring costs ~33 shared-memory loads
// per cell pair instead).
__global__ void __launch_bounds__(AC2_THREADS, 1)
ac_adj2_kernel(AcGeom g, AcTiling t, const double* __restrict__ ub1, const double* __restrict__ ub2,
               const double* __restrict__ w1, const double* __restrict__ w2, const double* __restrict__ c2,
               double* __restrict__ ubA, double* __restrict__ ubB, double* __restrict__ G, AcPoints rcvh,
               const double* __restrict__ res_row1, AcPoints rcv, const double* __restrict__ res_row2, AcPoints src,
               double* __restrict__ gsrcv_row1, double* __restrict__ gsrcv_row2
#ifdef ADSEIS_TIMELINE
               , unsigned long long* tl
#endif
               ) {
#ifdef ADSEIS_TIMELINE
  struct TlExit { unsigned long long* p; __device__ ~TlExit() { if (p && threadIdx.x == 0) atomicMax(p + 3, tl_now()); } } tl_exit{tl};
  if (tl && threadIdx.x == 0) { atomicMin(tl, tl_now()); atomicMax(tl + 1, tl_now()); }
#endif
  pdl_launch_dependents();
  const int bid = blockIdx.x;
  const int ld = g.ld;
  const int ct = bid % t.nct, tr = bid / t.nct;
  int r0, r1;
  ac_row_tile(t, tr, &r0, &r1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = t.mc0 + ct * AC_TILE_COLS;
  const int jb = c0 + warp * AC2_WCOLS;
  const int j = jb + lane;
  const bool act = j < t.mc_end;
  pdl_wait();
  extern __shared__ __align__(128) unsigned char ac_smem[];
  AcAdj2Stage* stg = reinterpret_cast<AcAdj2Stage*>(ac_smem);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ac_smem + AC_NST_ADJ2 * sizeof(AcAdj2Stage));
  unsigned long long* empty = full + AC_NST_ADJ2;
  double* priv = reinterpret_cast<double*>(empty + AC_NST_ADJ2) + warp * AC2_PRIV;
  const int nrows = r1 - r0;
  const int nit = nrows + 2;   // rows q = r0-1 .. r1 of A
  const int nst = nit + 2;     // stages (two priming rows)
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < AC_NST_ADJ2; k++) { mbar_init(full + k, 1); mbar_init(empty + k, AC2_WARPS); }
    mbar_init_fence();
  }
  __syncthreads();
  if (warp == AC2_WARPS) {
    if (lane == 0) {
      const unsigned hbytes = (unsigned)min(AC_HCOLS, ld - (c0 - 2)) * 8u;
      for (int k = 0; k < nst; k++) {
        const int sidx = k % AC_NST_ADJ2;
        if (k >= AC_NST_ADJ2) { mbar_wait(empty + sidx, (unsigned)(k / AC_NST_ADJ2 - 1) & 1u); ring_refill_fence(); }
        AcAdj2Stage& s = stg[sidx];
        const i64 ro = (i64)(r0 - 2 + k) * ld + c0 - 2;   // row r0-2+k >= 1
        mbar_arrive_expect_tx(full + sidx, (k >= 2 ? 6u : 3u) * hbytes);
        bulk_g2s(s.ub1, ub1 + ro, hbytes, full + sidx);
        bulk_g2s(s.c2, c2 + ro, hbytes, full + sidx);
        bulk_g2s(s.w1, w1 + ro, hbytes, full + sidx);
        if (k >= 2) {
          bulk_g2s(s.ub2, ub2 + ro - ld, hbytes, full + sidx);
          bulk_g2s(s.w2, w2 + ro - ld, hbytes, full + sidx);
          bulk_g2s(s.G, G + ro - 2 * ld, hbytes, full + sidx);
        }
      }
    }
  } else {
    const bool wact = jb < t.mc_end;
    const double rx2 = g.rx * g.rx, ry2 = g.ry * g.ry, kk = -g.kx2 - g.ky2, kx2 = g.kx2, ky2 = g.ky2;
    double cAm = 0.0, cAc = 0.0, Ac = 0.0, a0prev = 0.0, t1p = 0.0;   // carried from the previous rows
    const bool edge = (lane == 0) || (lane == 31);
    const int so = 2 + warp * AC2_WCOLS + lane;           // own column inside a staged row
    const int sr = (lane == 0) ? so - 1 : so + 1;         // rim column (lanes 0 / 31)
    const int sq = (lane == 0) ? so - 2 : so + 2;         // its outer neighbour
    const int ecol = (lane == 0) ? jb - 1 : jb + AC2_WCOLS;
    int ha = 0, hb = 0;
    if (rcvh.blk != nullptr && res_row1 != nullptr) { ha = rcvh.blk[bid]; hb = rcvh.blk[bid + 1]; }
    mbar_wait(full + 0, 0u);
    mbar_wait(full + 1, 0u);
    for (int it = 0; it < nit; it++) {
      const int q = r0 - 1 + it;
      const int i0 = (it + 2) % AC_NST_ADJ2, i1 = (it + 1) % AC_NST_ADJ2, i2 = it % AC_NST_ADJ2;
      mbar_wait(full + i0, (unsigned)((it + 2) / AC_NST_ADJ2) & 1u);
      if (wact) {
        const AcAdj2Stage &S0 = stg[i0], &S1 = stg[i1], &S2 = stg[i2];
        // ---- A = ubar[s-1] on row q
        const double u_c = S1.ub1[so], c_c = S1.c2[so];
        const double cgl = S1.c2[so - 1] * S1.ub1[so - 1], cgr = S1.c2[so + 1] * S1.ub1[so + 1];
        const double cgn = S0.c2[so] * S0.ub1[so], cgm = S2.c2[so] * S2.ub1[so];
        const double a0 = ac_a0(c_c, kx2, ky2);
        double A = ac_adj_cell(a0, rx2, ry2, u_c, cgn, cgm, cgr, cgl, S0.ub2[so]);
        double Ae = 0.0, ecc = 0.0;
        if (edge) {   // rim column: inner neighbour = own column, outer neighbour = column sq
          ecc = S1.c2[sr];
          const double cgo = S1.c2[sq] * S1.ub1[sq], cgi = c_c * u_c;
          Ae = ac_adj_cell(ac_a0(ecc, kx2, ky2), rx2, ry2, S1.ub1[sr], S0.c2[sr] * S0.ub1[sr], S2.c2[sr] * S2.ub1[sr],
                           lane == 0 ? cgi : cgo, lane == 0 ? cgo : cgi, S0.ub2[sr]);
        }
        // first correlation term of row q: stencil(u[s-1]) * ubar[s]
        const double t1 = ac_corr_cell(kk, rx2, ry2, S1.w1[so], S0.w1[so], S2.w1[so], S1.w1[so + 1], S1.w1[so - 1], u_c);
        // receiver residuals of slot s-1 on (tile + rim) cells of row q
        if (hb > ha) {
          const int key0 = q * ld + c0 - 1;
          const int k0 = ac_lower_bound(rcvh.cell, ha, hb, key0);
          if (k0 < hb && rcvh.cell[k0] <= key0 + AC_TILE_COLS + 1) {   // the row has entries within this tile + rim
            A = ac_add_points(A, rcvh, k0, hb, q * ld + j, res_row1, 1.0);
            if (edge) Ae = ac_add_points(Ae, rcvh, k0, hb, q * ld + ecol, res_row1, 1.0);
          }
        }
        if (act && it >= 1 && it <= nrows) ubA[(i64)q * ld + j] = A;
        const double cAn = c_c * A;
        double* pw = priv + (it & 1) * (AC2_WCOLS + 2);
        pw[1 + lane] = cAn;
        if (edge) pw[lane == 0 ? 0 : AC2_WCOLS + 1] = ecc * Ae;
        // ---- B = ubar[s-2] on row q-1 and both correlation terms of that row
        if (it >= 2) {
          const double* pr = priv + ((it - 1) & 1) * (AC2_WCOLS + 2);   // c^2 A of row q-1 (written one iteration ago)
          if (act) {
            const i64 o = (i64)(q - 1) * ld + j;
            ubB[o] = ac_adj_cell(a0prev, rx2, ry2, Ac, cAn, cAm, pr[lane + 2], pr[lane], S2.ub1[so]);
            G[o] = (S0.G[so] + t1p) +
                   ac_corr_cell(kk, rx2, ry2, S1.w2[so], S0.w2[so], S2.w2[so], S1.w2[so + 1], S1.w2[so - 1], Ac);
          }
        }
        cAm = cAc; cAc = cAn; Ac = A; a0prev = a0; t1p = t1;
      }
      __syncwarp();   // stage reads and the private-row accesses of this iteration are complete in every lane
      if (lane == 0) mbar_arrive(empty + i2);
    }
    // the last two stages were read but never released inside the loop: nothing waits for them
  }
  const AcPoints none{};
  ac_cta_epilogue(bid, ubA, none, nullptr, 0.0, src, gsrcv_row1, g.dt2);   // A already carries its residuals
  ac_cta_epilogue(bid, ubB, rcv, res_row2, 1.0, src, gsrcv_row2, g.dt2);
}
