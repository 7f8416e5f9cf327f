// acoustic_kernels.cuh -- fused acoustic time-step kernels (forward and exact adjoint) for sm_100a.
//
// One launch = one time step over a range of local rows:
//   forward : u[s]   = Step(u[s-1], u[s-2], phi, psi; c^2, sigma_x(i), tau_y(j))  + source injection into u[s]
//                      + receiver sampling of u[s]                       (reference: AcousticOneStepCpu.h:1-48,
//                      ScatterAddOps.h:1-7 via Core.jl:600-601, gather via Core.jl:727-728)
//   adjoint : ubar[s-1] = Step^T in GATHER form (scatter form AcousticOneStepCpu.h:51-125)
//                      + cross-correlation gbar_c2 += cbar_s in the same pass
//                      + receiver-residual injection into ubar[s-1] + grad_srcv sampling
//
// Mapping (HBM-bound fp64 stencil, no tensor cores): a warp owns 64 consecutive columns (one double2 per lane,
// 512 B per row -> fully coalesced 128 B lines) and marches down AC_RB rows keeping the 3-row window of the
// stencil in registers; left/right neighbours come from warp shuffles, the two warp-edge values from one
// predicated 8-byte load that hits L1/L2.  Every array element is therefore requested from DRAM once per step:
// 32 B/cell forward (w, wold, c^2 read + u written), 56 B/cell adjoint.  PML work (phi/psi traffic, the fp64
// divide) is confined to the rows/warps that intersect the absorbing frame; the frame test is warp-uniform.
// Arithmetic in the forward kernel is written in the reference's evaluation order and the library is compiled
// with -fmad=false, so forward wavefields and traces are bit-identical to the CPU op.
#pragma once
#include "common.cuh"

#define AC_WARPS 8
#define AC_THREADS (AC_WARPS * 32)
#define AC_WCOLS 64                       // columns per warp
#define AC_TILE_COLS (AC_WARPS * AC_WCOLS)  // columns per CTA
#define AC_RB 32                          // rows per CTA
#define AC_U 4                            // rows whose loads are issued together (memory-level parallelism)

struct AcGeom {
  int H, W;    // global padded rows (NX+2) and columns (NY+2)
  int Hl, ld;  // local rows held by this GPU (incl. halo rows) and pitch in doubles
  int goff;    // global row index of local row 0
  int fi0, fi1, fj0, fj1;  // inclusive global rows / columns of the PML-free fast region (empty if fi0 > fi1)
  double dt, hx, hy;
  double kx2, ky2;  // 2*dt*dt/hx/hx , 2*dt*dt/hy/hy
  double rx, ry;    // dt/hx , dt/hy
  double px, py;    // dt*dt/(2.0*hx) , dt*dt/(2.0*hy)
  double dt2;       // dt*dt
  i64 plane;        // Hl*ld
};

// Peer-memory halo targets of a slab (all null on a single GPU).  After finishing a tile that contains its first
// (last) owned row, a CTA stores that row into the lower (upper) neighbour's halo row of the same array.
struct AcPeer {
  double* lo_u;    // address of the row in rank-1's array that mirrors my first owned row (its upper halo row)
  double* hi_u;    // address of the row in rank+1's array that mirrors my last owned row (its lower halo row)
  double* lo_phi;
  double* hi_phi;
};

// ------------------------------------------------------------------------------------------------------------
// forward, general (PML / ring / pad) cell: literal AcousticOneStepCpu.h:27-44
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ac_fwd_general_cell(const AcGeom& g, int li, int j, const double* __restrict__ w,
                                                    const double* __restrict__ wold, const double* __restrict__ c2,
                                                    const double* __restrict__ phi, const double* __restrict__ psi,
                                                    const double* __restrict__ sigx, const double* __restrict__ tauy,
                                                    double* __restrict__ u, double* __restrict__ phio,
                                                    double* __restrict__ psio) {
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
  u[IJ] = v / (1 + (sg + ta) / 2 * dt);
  phio[IJ] = (1. - dt * sg) * phi[IJ] + dt * c * (ta - sg) / 2.0 / g.hx * (w[IpJ] - w[InJ]);
  psio[IJ] = (1. - dt * ta) * psi[IJ] + dt * c * (sg - ta) / 2.0 / g.hy * (w[IJp] - w[IJn]);
}

// Tile epilogue shared by both kernels: add `scale * val[perm]` into field[cell] for the points of set `inj` that
// lie in this tile (sequentially per cell, in original point order), then sample field[cell] * sscale into
// out[perm] for the points of set `smp`.
__device__ __forceinline__ void ac_tile_epilogue(double* __restrict__ field, const PointSetDev& inj,
                                                 const double* __restrict__ inj_val, double inj_scale, int ia, int ib,
                                                 const PointSetDev& smp, double* __restrict__ smp_out,
                                                 double smp_scale, int sa, int sb) {
  if (inj_val != nullptr) {
    for (int k = ia + threadIdx.x; k < ib; k += blockDim.x) {
      const int cell = inj.cell[k];
      double v = field[cell];
      for (int m = inj.start[k]; m < inj.start[k + 1]; m++) v += inj_val[inj.perm[m]] * inj_scale;
      field[cell] = v;
    }
  }
  if (smp_out != nullptr && sb > sa) {
    __syncthreads();  // sb>sa is CTA-uniform
    for (int k = sa + threadIdx.x; k < sb; k += blockDim.x) {
      const double v = field[smp.cell[k]] * smp_scale;
      for (int m = smp.start[k]; m < smp.start[k + 1]; m++) smp_out[smp.perm[m]] = v;
    }
  }
}

// Store one finished row into a neighbour GPU's halo row over NVLink (peer pointer), 16 B per lane.
__device__ __forceinline__ void ac_push_row(double* __restrict__ peer_row, const double* __restrict__ my_row, int ct,
                                            int ld) {
  const int j = ct * AC_TILE_COLS + 2 * threadIdx.x;
  if (j < ld) st2(peer_row + j, ld2(my_row + j));
}

// ------------------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AC_THREADS)
ac_fwd_kernel(AcGeom g, int l0, int l1, const double* __restrict__ w, const double* __restrict__ wold,
              const double* __restrict__ c2, const double* __restrict__ phi, const double* __restrict__ psi,
              const double* __restrict__ sigx, const double* __restrict__ tauy, double* __restrict__ u,
              double* __restrict__ phio, double* __restrict__ psio, PointSetDev src,
              const double* __restrict__ srcv_row, PointSetDev rcv, double* __restrict__ rcvv_row) {
  __shared__ int s_rng[4];
  const int ct = blockIdx.x;
  const int r0 = l0 + blockIdx.y * AC_RB;
  const int r1 = min(l1, r0 + AC_RB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jb = ct * AC_TILE_COLS + warp * AC_WCOLS;
  const int j = jb + 2 * lane;
  const int ld = g.ld;
  if (threadIdx.x == 0) ps_range(src, ct, r0, r1, ld, g.plane, &s_rng[0], &s_rng[1]);
  if (threadIdx.x == 32) ps_range(rcv, ct, r0, r1, ld, g.plane, &s_rng[2], &s_rng[3]);

  if (j < ld) {
    const bool fastcols = (jb >= g.fj0) && (jb + AC_WCOLS - 1 <= g.fj1) && (g.fi0 <= g.fi1);
    if (!fastcols) {
      for (int li = r0; li < r1; li++) {
        ac_fwd_general_cell(g, li, j, w, wold, c2, phi, psi, sigx, tauy, u, phio, psio);
        ac_fwd_general_cell(g, li, j + 1, w, wold, c2, phi, psi, sigx, tauy, u, phio, psio);
      }
    } else {
      const double2 z2 = make_double2(0.0, 0.0);
      double2 wm = (r0 > 0) ? ld2(w + (i64)(r0 - 1) * ld + j) : z2;
      double2 wc = ld2(w + (i64)r0 * ld + j);
      for (int rb = r0; rb < r1; rb += AC_U) {
        double2 wn[AC_U], wo[AC_U], cc[AC_U];
        double we[AC_U];
#pragma unroll
        for (int k = 0; k < AC_U; k++) {
          const int li = rb + k;
          if (li < r1) {
            const i64 ro = (i64)li * ld;
            wn[k] = (li + 1 < g.Hl) ? ld2(w + ro + ld + j) : z2;
            wo[k] = ld2_stream(wold + ro + j);
            cc[k] = ld2(c2 + ro + j);
            we[k] = 0.0;
            if (lane == 0) we[k] = w[ro + jb - 1];
            if (lane == 31) we[k] = w[ro + jb + AC_WCOLS];
          }
        }
#pragma unroll
        for (int k = 0; k < AC_U; k++) {
          const int li = rb + k;
          if (li < r1) {
            const int gi = g.goff + li;
            if (gi < g.fi0 || gi > g.fi1) {
              ac_fwd_general_cell(g, li, j, w, wold, c2, phi, psi, sigx, tauy, u, phio, psio);
              ac_fwd_general_cell(g, li, j + 1, w, wold, c2, phi, psi, sigx, tauy, u, phio, psio);
            } else {
              double lft = __shfl_up_sync(0xffffffffu, wc.y, 1);
              double rgt = __shfl_down_sync(0xffffffffu, wc.x, 1);
              if (lane == 0) lft = we[k];
              if (lane == 31) rgt = we[k];
              double2 o;
              {
                const double c = cc[k].x;
                o.x = (2 - g.kx2 * c - g.ky2 * c) * wc.x + c * g.rx * g.rx * (wn[k].x + wm.x) +
                      c * g.ry * g.ry * (wc.y + lft) - wo[k].x;
              }
              {
                const double c = cc[k].y;
                o.y = (2 - g.kx2 * c - g.ky2 * c) * wc.y + c * g.rx * g.rx * (wn[k].y + wm.y) +
                      c * g.ry * g.ry * (rgt + wc.x) - wo[k].y;
              }
              st2(u + (i64)li * ld + j, o);
            }
            wm = wc;
            wc = wn[k];
          }
        }
      }
    }
  }
  __syncthreads();
  ac_tile_epilogue(u, src, srcv_row, g.dt2, s_rng[0], s_rng[1], rcv, rcvv_row, 1.0, s_rng[2], s_rng[3]);
}

// ------------------------------------------------------------------------------------------------------------
// adjoint helpers
// ------------------------------------------------------------------------------------------------------------
// g(Q) = ubar(Q) / D(Q) for interior Q, 0 otherwise (the ring is a constant output of the forward step)
__device__ __forceinline__ double ac_g_at(const AcGeom& g, const double* __restrict__ ub,
                                          const double* __restrict__ sigx, const double* __restrict__ tauy, int li,
                                          int j) {
  const int gi = g.goff + li;
  if (gi < 1 || gi > g.H - 2 || j < 1 || j > g.W - 2) return 0.0;
  return ub[(i64)li * g.ld + j] / (1 + (sigx[gi] + tauy[j]) / 2 * g.dt);
}

__device__ __forceinline__ void ac_adj_general_cell(const AcGeom& g, int li, int j, const double* __restrict__ ub1,
                                                    const double* __restrict__ ub2, const double* __restrict__ wf,
                                                    const double* __restrict__ c2, const double* __restrict__ phib,
                                                    const double* __restrict__ psib,
                                                    const double* __restrict__ sigx, const double* __restrict__ tauy,
                                                    double* __restrict__ ub0, double* __restrict__ phibo,
                                                    double* __restrict__ psibo, double* __restrict__ G) {
  const int gi = g.goff + li;
  const i64 IJ = (i64)li * g.ld + j;
  if (j >= g.W) { ub0[IJ] = 0.0; return; }
  const double dt = g.dt;
  const bool intP = (gi >= 1 && gi <= g.H - 2 && j >= 1 && j <= g.W - 2);
  const bool colok = (j >= 1 && j <= g.W - 2);
  double acc = 0.0, sg = 0.0, ta = 0.0, gP = 0.0;
  if (intP) {
    sg = sigx[gi]; ta = tauy[j];
    const double c = c2[IJ];
    gP = ub1[IJ] / (1 + (sg + ta) / 2 * dt);
    acc = (2 - sg * ta * dt * dt - g.kx2 * c - g.ky2 * c) * gP;
  }
  double gxm = 0.0, gxp = 0.0, gym = 0.0, gyp = 0.0;
  if (colok && gi - 1 >= 1 && gi - 1 <= g.H - 2) {  // Q = P - e_x : grad_w[IpJ] of Q
    const i64 Q = IJ - g.ld;
    const double cQ = c2[Q], sQ = sigx[gi - 1], tQ = tauy[j];
    gxm = ub1[Q] / (1 + (sQ + tQ) / 2 * dt);
    acc += cQ * g.rx * g.rx * gxm + dt * cQ * (tQ - sQ) / 2.0 / g.hx * phib[Q];
  }
  if (colok && gi + 1 >= 1 && gi + 1 <= g.H - 2) {  // Q = P + e_x : grad_w[InJ] of Q
    const i64 Q = IJ + g.ld;
    const double cQ = c2[Q], sQ = sigx[gi + 1], tQ = tauy[j];
    gxp = ub1[Q] / (1 + (sQ + tQ) / 2 * dt);
    acc += cQ * g.rx * g.rx * gxp - dt * cQ * (tQ - sQ) / 2.0 / g.hx * phib[Q];
  }
  const bool rowok = (gi >= 1 && gi <= g.H - 2);
  if (rowok && j - 1 >= 1) {  // Q = P - e_y : grad_w[IJp] of Q
    const i64 Q = IJ - 1;
    const double cQ = c2[Q], sQ = sigx[gi], tQ = tauy[j - 1];
    gym = ub1[Q] / (1 + (sQ + tQ) / 2 * dt);
    acc += cQ * g.ry * g.ry * gym + dt * cQ * (sQ - tQ) / 2.0 / g.hy * psib[Q];
  }
  if (rowok && j + 1 <= g.W - 2) {  // Q = P + e_y : grad_w[IJn] of Q
    const i64 Q = IJ + 1;
    const double cQ = c2[Q], sQ = sigx[gi], tQ = tauy[j + 1];
    gyp = ub1[Q] / (1 + (sQ + tQ) / 2 * dt);
    acc += cQ * g.ry * g.ry * gyp - dt * cQ * (sQ - tQ) / 2.0 / g.hy * psib[Q];
  }
  if (intP) {
    acc += -(1 - (sg + ta) * dt / 2) * (ub2[IJ] / (1 + (sg + ta) / 2 * dt));  // grad_wold of step s+1
    const double pb = phib[IJ], qb = psib[IJ];
    phibo[IJ] = (1. - dt * sg) * pb + g.px * (gxm - gxp);
    psibo[IJ] = (1. - dt * ta) * qb + g.py * (gym - gyp);
    const double wC = wf[IJ], wU = wf[IJ + g.ld], wD = wf[IJ - g.ld], wR = wf[IJ + 1], wL = wf[IJ - 1];
    const double cb = ((-g.kx2 - g.ky2) * wC + g.rx * g.rx * (wU + wD) + g.ry * g.ry * (wR + wL)) * gP +
                      dt * (ta - sg) / 2.0 / g.hx * (wU - wD) * pb + dt * (sg - ta) / 2.0 / g.hy * (wR - wL) * qb;
    G[IJ] += cb;
  }
  ub0[IJ] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// adjoint kernel: ub0 = ubar[s-1] from ub1 = ubar[s], ub2 = ubar[s+1], wf = u[s-1]
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AC_THREADS)
ac_adj_kernel(AcGeom g, int l0, int l1, const double* __restrict__ ub1, const double* __restrict__ ub2,
              const double* __restrict__ wf, const double* __restrict__ c2, const double* __restrict__ phib,
              const double* __restrict__ psib, const double* __restrict__ sigx, const double* __restrict__ tauy,
              double* __restrict__ ub0, double* __restrict__ phibo, double* __restrict__ psibo,
              double* __restrict__ G, PointSetDev rcv, const double* __restrict__ res_row, PointSetDev src,
              double* __restrict__ gsrcv_row) {
  __shared__ int s_rng[4];
  const int ct = blockIdx.x;
  const int r0 = l0 + blockIdx.y * AC_RB;
  const int r1 = min(l1, r0 + AC_RB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jb = ct * AC_TILE_COLS + warp * AC_WCOLS;
  const int j = jb + 2 * lane;
  const int ld = g.ld;
  if (threadIdx.x == 0) ps_range(rcv, ct, r0, r1, ld, g.plane, &s_rng[0], &s_rng[1]);
  if (threadIdx.x == 32) ps_range(src, ct, r0, r1, ld, g.plane, &s_rng[2], &s_rng[3]);

  if (j < ld) {
    const bool fastcols = (jb >= g.fj0) && (jb + AC_WCOLS - 1 <= g.fj1) && (g.fi0 <= g.fi1);
    if (!fastcols) {
      for (int li = r0; li < r1; li++) {
        ac_adj_general_cell(g, li, j, ub1, ub2, wf, c2, phib, psib, sigx, tauy, ub0, phibo, psibo, G);
        ac_adj_general_cell(g, li, j + 1, ub1, ub2, wf, c2, phib, psib, sigx, tauy, ub0, phibo, psibo, G);
      }
    } else {
      const double2 z2 = make_double2(0.0, 0.0);
      const double rx2 = g.rx * g.rx, ry2 = g.ry * g.ry, kk = -g.kx2 - g.ky2;
      // windows: cg = c^2 * ubar[s] (D == 1 on every row a fast row can see), w = u[s-1]
      double2 cgm = z2, cgc, wm = z2, wc, gc, ccen;
      if (r0 > 0) {
        const i64 ro = (i64)(r0 - 1) * ld + j;
        const double2 c_ = ld2(c2 + ro), u_ = ld2(ub1 + ro);
        cgm = make_double2(c_.x * u_.x, c_.y * u_.y);
        wm = ld2(wf + ro);
      }
      {
        const i64 ro = (i64)r0 * ld + j;
        ccen = ld2(c2 + ro);
        gc = ld2(ub1 + ro);
        cgc = make_double2(ccen.x * gc.x, ccen.y * gc.y);
        wc = ld2(wf + ro);
      }
      for (int rb = r0; rb < r1; rb += AC_U) {
        double2 un[AC_U], cn[AC_U], wn[AC_U], u2[AC_U], Gr[AC_U];
        double ecg[AC_U], ew[AC_U];
#pragma unroll
        for (int k = 0; k < AC_U; k++) {
          const int li = rb + k;
          if (li < r1) {
            const i64 ro = (i64)li * ld;
            if (li + 1 < g.Hl) {
              un[k] = ld2(ub1 + ro + ld + j);
              cn[k] = ld2(c2 + ro + ld + j);
              wn[k] = ld2(wf + ro + ld + j);
            } else {
              un[k] = z2; cn[k] = z2; wn[k] = z2;
            }
            u2[k] = ld2_stream(ub2 + ro + j);
            Gr[k] = ld2_stream(G + ro + j);
            ecg[k] = 0.0; ew[k] = 0.0;
            if (lane == 0) { ecg[k] = c2[ro + jb - 1] * ub1[ro + jb - 1]; ew[k] = wf[ro + jb - 1]; }
            if (lane == 31) { ecg[k] = c2[ro + jb + AC_WCOLS] * ub1[ro + jb + AC_WCOLS]; ew[k] = wf[ro + jb + AC_WCOLS]; }
          }
        }
#pragma unroll
        for (int k = 0; k < AC_U; k++) {
          const int li = rb + k;
          if (li < r1) {
            const int gi = g.goff + li;
            const double2 cgp = make_double2(cn[k].x * un[k].x, cn[k].y * un[k].y);
            if (gi < g.fi0 || gi > g.fi1) {
              ac_adj_general_cell(g, li, j, ub1, ub2, wf, c2, phib, psib, sigx, tauy, ub0, phibo, psibo, G);
              ac_adj_general_cell(g, li, j + 1, ub1, ub2, wf, c2, phib, psib, sigx, tauy, ub0, phibo, psibo, G);
            } else {
              double cgl = __shfl_up_sync(0xffffffffu, cgc.y, 1);
              double cgr = __shfl_down_sync(0xffffffffu, cgc.x, 1);
              double wl = __shfl_up_sync(0xffffffffu, wc.y, 1);
              double wr = __shfl_down_sync(0xffffffffu, wc.x, 1);
              if (lane == 0) { cgl = ecg[k]; wl = ew[k]; }
              if (lane == 31) { cgr = ecg[k]; wr = ew[k]; }
              double2 o, Go;
              {
                const double c = ccen.x, gg = gc.x;
                o.x = (2 - g.kx2 * c - g.ky2 * c) * gg + rx2 * (cgp.x + cgm.x) + ry2 * (cgc.y + cgl) - u2[k].x;
                Go.x = Gr[k].x + (kk * wc.x + rx2 * (wn[k].x + wm.x) + ry2 * (wc.y + wl)) * gg;
              }
              {
                const double c = ccen.y, gg = gc.y;
                o.y = (2 - g.kx2 * c - g.ky2 * c) * gg + rx2 * (cgp.y + cgm.y) + ry2 * (cgr + cgc.x) - u2[k].y;
                Go.y = Gr[k].y + (kk * wc.y + rx2 * (wn[k].y + wm.y) + ry2 * (wr + wc.x)) * gg;
              }
              st2(ub0 + (i64)li * ld + j, o);
              st2(G + (i64)li * ld + j, Go);
            }
            cgm = cgc; cgc = cgp;
            wm = wc; wc = wn[k];
            gc = un[k]; ccen = cn[k];
          }
        }
      }
    }
  }
  __syncthreads();
  ac_tile_epilogue(ub0, rcv, res_row, 1.0, s_rng[0], s_rng[1], src, gsrcv_row, g.dt2, s_rng[2], s_rng[3]);
}
