// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Builds oracle/_ref/libadseis_ref.so from the reference's OWN numerical C++ bodies, which are
// #included in place from /root/reference at build time (nothing is copied into this repo):
//   deps/CustomOps/AcousticOneStepCpu/AcousticOneStepCpu.h   (AcousticOneStepCpuForward/Backward)
//   deps/CustomOps/MPIAcousticOneStepCpu/MpiAcousticOneStep.h (MpiAcousticOneStepCpuForward/Backward)
//   deps/CustomOps/SourceOps/AddSource.cpp                    (forwardCPU/backwardCPU)
//   deps/CustomOps/ReceiveOps/GetReceive.cpp                  (forward/backward)
//   deps/CustomOps/GatherOps/GatherOps.h, ScatterAddOps/ScatterAddOps.h, ScatterNdOps/ScatterNdOps.h
// The TensorFlow #includes of the two .cpp files resolve to the empty headers in oracle/tf_stubs.
//
// What is mine in this file: thin extern "C" wrappers (ref_op_*) and "drivers" (ref_drv_*) that call
// those bodies in the order the reference's Julia graph builders do (src/Core.jl:562-620 for the
// acoustic loop, src/Core.jl:31-228 for the elastic loop, src/MPIAcoustic.jl:251-404 for the block
// decomposed loop).  The drivers are what bench.py times as the "reference" CPU arm and what the
// C oracle (oracle/oracle.c) is pinned against in tests/test_oracle_pinning.py.  The elastic solvers and the acoustic
// PropagatorKernel=0 scheme are TensorFlow graphs over the reference's gather / scatter / add_source / get_receive
// ops; their drivers live in ref_graph.inc (#included at the end), which records and differentiates those graphs
// with the reference's own forward AND backward op bodies.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long long int64;
namespace tensorflow {}

#include "AcousticOneStepCpu/AcousticOneStepCpu.h"
#include "MPIAcousticOneStepCpu/MpiAcousticOneStep.h"
#include "GatherOps/GatherOps.h"
#include "ScatterAddOps/ScatterAddOps.h"
#include "ScatterNdOps/ScatterNdOps.h"
namespace ref_source {
#include "SourceOps/AddSource.cpp"
}
// AddSource.cpp leaves its accessor macros defined; drop them before the next body.
#undef vx
#undef vy
#undef vx_
#undef vy_
#undef sigmaxx
#undef sigmaxx_
#undef sigmayy
#undef sigmayy_
#undef sigmaxy
#undef sigmaxy_
#undef g_vx
#undef g_vy
#undef g_vx_
#undef g_vy_
#undef g_sigmaxx
#undef g_sigmaxx_
#undef g_sigmayy
#undef g_sigmayy_
#undef g_sigmaxy
#undef g_sigmaxy_
namespace ref_receive {
#include "ReceiveOps/GetReceive.cpp"
}

#define API extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------------
// op-level wrappers (argument order = the reference bodies')
// ------------------------------------------------------------------------------------------------
API void ref_op_acoustic_step_fwd(const double* w, const double* wold, const double* phi, const double* psi,
                                  const double* sigma, const double* tau, const double* c, double dt, double hx,
                                  double hy, int64 NX, int64 NY, double* u, double* phiout, double* psiout) {
  AcousticOneStepCpuForward(w, wold, phi, psi, sigma, tau, c, dt, hx, hy, NX, NY, u, phiout, psiout);
}

// Outputs are zeroed first, as AcousticOneStepCpu.cpp:363-367 does before calling the body.
API void ref_op_acoustic_step_bwd(double* grad_w, double* grad_wold, double* grad_phi, double* grad_psi,
                                  double* grad_c, const double* grad_u, const double* grad_phiout,
                                  const double* grad_psiout, const double* w, const double* wold,
                                  const double* phi, const double* psi, const double* sigma, const double* tau,
                                  const double* c, double dt, double hx, double hy, int64 NX, int64 NY) {
  size_t N = (size_t)(NX + 2) * (NY + 2);
  memset(grad_w, 0, N * 8); memset(grad_wold, 0, N * 8); memset(grad_phi, 0, N * 8);
  memset(grad_psi, 0, N * 8); memset(grad_c, 0, N * 8);
  AcousticOneStepCpuBackward(grad_w, grad_wold, grad_phi, grad_psi, grad_c, grad_u, grad_phiout, grad_psiout,
                             w, wold, phi, psi, sigma, tau, c, dt, hx, hy, NX, NY, nullptr, nullptr, nullptr);
}

API void ref_op_mpi_acoustic_step_fwd(const double* w, const double* wold, const double* phi, const double* psi,
                                      const double* sigma, const double* tau, const double* c, double dt,
                                      double hx, double hy, int64 NX, int64 NY, double* u, double* phiout,
                                      double* psiout) {
  MpiAcousticOneStepCpuForward(w, wold, phi, psi, sigma, tau, c, dt, hx, hy, NX, NY, u, phiout, psiout);
}

API void ref_op_add_source_fwd(double* vx_, double* vy_, double* sxx_, double* syy_, double* sxy_, const double* vx,
                               const double* vy, const double* sxx, const double* syy, const double* sxy,
                               const int64* srci, const int64* srcj, const double* srcv, const int64* srctype,
                               int64 nsrc, int64 NX, int64 NY) {
  ref_source::forwardCPU(vx_, vy_, sxx_, syy_, sxy_, vx, vy, sxx, syy, sxy, srci, srcj, srcv, srctype, nsrc, NX, NY);
}

API void ref_op_add_source_bwd(double* g_vx, double* g_vy, double* g_sxx, double* g_syy, double* g_sxy,
                               double* grad_srcv, const double* g_vx_, const double* g_vy_, const double* g_sxx_,
                               const double* g_syy_, const double* g_sxy_, const int64* srci, const int64* srcj,
                               const double* srcv, const int64* srctype, int64 nsrc, int64 NX, int64 NY) {
  ref_source::backwardCPU(g_vx, g_vy, g_sxx, g_syy, g_sxy, grad_srcv, g_vx_, g_vy_, g_sxx_, g_syy_, g_sxy_, srci,
                          srcj, srcv, srctype, nsrc, NX, NY);
}

API void ref_op_get_receive_fwd(double* out, const double* vx, const double* vy, const double* sxx,
                                const double* syy, const double* sxy, int64 nt, const int64* rcvi,
                                const int64* rcvj, const int64* rcvtype, int64 nrcv, int64 NX, int64 NY) {
  ref_receive::forward(out, vx, vy, sxx, syy, sxy, nt, rcvi, rcvj, rcvtype, nrcv, NX, NY);
}

API void ref_op_get_receive_bwd(double* d_vx, double* d_vy, double* d_sxx, double* d_syy, double* d_sxy,
                                const double* d_out, int64 nt, const int64* rcvi, const int64* rcvj,
                                const int64* rcvtype, int64 nrcv, int64 NX, int64 NY) {
  ref_receive::backward(d_vx, d_vy, d_sxx, d_syy, d_sxy, d_out, nt, rcvi, rcvj, rcvtype, nrcv, NX, NY);
}

API void ref_op_gather_fwd(double* out, const double* v, const int64* ii, int n) { GatherOps_forward(out, v, ii, n); }
API void ref_op_scatter_add_fwd(double* out, const double* ipt, const int64* ii, const double* upd, int d, int n) {
  ScatterAddOps_forward(out, ipt, ii, upd, d, n);
}
API void ref_op_scatter_nd_fwd(double* out, const int64* ii, const double* upd, int n, int m) {
  memset(out, 0, sizeof(double) * (size_t)m);  // ScatterNdOps.cpp:92 zero-fills before the body
  ScatterNdOps_forward(out, ii, upd, n);
}

// ------------------------------------------------------------------------------------------------
// Driver: single-process acoustic loop, PropagatorKernel=1 on CPU  (src/Core.jl:562-620, 726-730)
// 0-based slots: u[0]=u[1]=0; s=2..NSTEP: step, then u[s][src] += srcv0[s-1]*dt^2
// ------------------------------------------------------------------------------------------------
API void ref_drv_acoustic_forward(int64 NX, int64 NY, int64 NSTEP, double dt, double hx, double hy,
                                  const double* sigma, const double* tau, const double* c /* velocity */,
                                  int64 nsrc, const int64* srci, const int64* srcj, const double* srcv,
                                  int64 nrcv, const int64* rcvi, const int64* rcvj,
                                  double* u /* (NSTEP+1)*N */, double* rcvv /* (NSTEP+1)*nrcv or NULL */) {
  const int64 N = (NX + 2) * (NY + 2);
  std::vector<double> c2(N), phi(N, 0.0), psi(N, 0.0), phin(N), psin(N), ut(N), upd(nsrc > 0 ? nsrc : 1);
  std::vector<int64> sidx(nsrc > 0 ? nsrc : 1);
  for (int64 k = 0; k < N; k++) c2[k] = c[k] * c[k];                       // Core.jl:564
  for (int64 k = 0; k < nsrc; k++) sidx[k] = (srci[k] - 1) * (NY + 2) + srcj[k];  // Core.jl:600 (1-based)
  memset(u, 0, sizeof(double) * 2 * N);                                     // Core.jl:607-612
  for (int64 s = 2; s <= NSTEP; s++) {
    AcousticOneStepCpuForward(u + (s - 1) * N, u + (s - 2) * N, phi.data(), psi.data(), sigma, tau, c2.data(), dt,
                              hx, hy, NX, NY, ut.data(), phin.data(), psin.data());
    for (int64 k = 0; k < nsrc; k++) upd[k] = srcv[(s - 1) * nsrc + k] * (dt * dt);  // Core.jl:601
    ScatterAddOps_forward(u + s * N, ut.data(), sidx.data(), upd.data(), (int)N, (int)nsrc);
    phi.swap(phin);
    psi.swap(psin);
  }
  if (rcvv)
    for (int64 s = 0; s <= NSTEP; s++)
      for (int64 r = 0; r < nrcv; r++)                                       // Core.jl:727-728
        rcvv[s * nrcv + r] = u[s * N + (rcvi[r] - 1) * (NY + 2) + rcvj[r] - 1];
}

// loss = sum (rcvv-obs)^2 (src/Utils.jl:308) and its gradient w.r.t. c (velocity) and srcv, by composing the
// reference backward bodies in reverse graph order.  u must hold the forward history.
API void ref_drv_acoustic_gradient(int64 NX, int64 NY, int64 NSTEP, double dt, double hx, double hy,
                                   const double* sigma, const double* tau, const double* c, int64 nsrc,
                                   const int64* srci, const int64* srcj, int64 nrcv, const int64* rcvi,
                                   const int64* rcvj, const double* obs, const double* u, double* loss_out,
                                   double* grad_c /* N */, double* grad_srcv /* NSTEP*nsrc or NULL */) {
  const int64 N = (NX + 2) * (NY + 2);
  std::vector<double> c2(N), G(N, 0.0), gphi(N, 0.0), gpsi(N, 0.0);
  std::vector<double> gw(N), gwold(N), gphi_in(N), gpsi_in(N), gc(N);
  std::vector<double> ub[3] = {std::vector<double>(N), std::vector<double>(N), std::vector<double>(N)};
  for (int64 k = 0; k < N; k++) c2[k] = c[k] * c[k];
  double loss = 0.0;
  for (int64 s = 0; s <= NSTEP; s++)
    for (int64 r = 0; r < nrcv; r++) {
      double d = u[s * N + (rcvi[r] - 1) * (NY + 2) + rcvj[r] - 1] - obs[s * nrcv + r];
      loss += d * d;
    }
  *loss_out = loss;
  auto seed = [&](std::vector<double>& b, int64 s) {  // d loss / d u[s] through the receiver gather
    std::fill(b.begin(), b.end(), 0.0);
    for (int64 r = 0; r < nrcv; r++) {
      int64 id = (rcvi[r] - 1) * (NY + 2) + rcvj[r] - 1;
      b[id] += 2.0 * (u[s * N + id] - obs[s * nrcv + r]);
    }
  };
  if (grad_srcv) memset(grad_srcv, 0, sizeof(double) * NSTEP * nsrc);
  seed(ub[NSTEP % 3], NSTEP);
  if (NSTEP >= 1) seed(ub[(NSTEP - 1) % 3], NSTEP - 1);
  for (int64 s = NSTEP; s >= 2; s--) {
    std::vector<double>& gu = ub[s % 3];
    seed(ub[(s - 2) % 3], s - 2);
    // ScatterAddOps_backward: grad_ipt = grad_out (pass-through), grad_update[k] = grad_out[ii[k]-1]
    if (grad_srcv)
      for (int64 k = 0; k < nsrc; k++)
        grad_srcv[(s - 1) * nsrc + k] = gu[(srci[k] - 1) * (NY + 2) + srcj[k] - 1] * (dt * dt);
    std::fill(gw.begin(), gw.end(), 0.0); std::fill(gwold.begin(), gwold.end(), 0.0);
    std::fill(gphi_in.begin(), gphi_in.end(), 0.0); std::fill(gpsi_in.begin(), gpsi_in.end(), 0.0);
    std::fill(gc.begin(), gc.end(), 0.0);
    AcousticOneStepCpuBackward(gw.data(), gwold.data(), gphi_in.data(), gpsi_in.data(), gc.data(), gu.data(),
                               gphi.data(), gpsi.data(), u + (s - 1) * N, u + (s - 2) * N, nullptr, nullptr, sigma,
                               tau, c2.data(), dt, hx, hy, NX, NY, nullptr, nullptr, nullptr);
    std::vector<double>& g1 = ub[(s - 1) % 3];
    std::vector<double>& g2 = ub[(s - 2) % 3];
    for (int64 k = 0; k < N; k++) { g1[k] += gw[k]; g2[k] += gwold[k]; G[k] += gc[k]; }
    gphi.swap(gphi_in);
    gpsi.swap(gpsi_in);
  }
  for (int64 k = 0; k < N; k++) grad_c[k] = 2.0 * c[k] * G[k];  // d(c^2)/dc, Core.jl:564
}

// ------------------------------------------------------------------------------------------------
// Driver: block-decomposed acoustic loop, PropagatorKernel=1 (src/MPIAcoustic.jl:251-294, 334-404) with the
// MPI ranks emulated by OpenMP threads (one n x n block per task) and mpi_halo_exchange emulated by copies out of
// shared global arrays with a zero fill at physical edges.  Global grid NX x NY (unpadded), M=NX/n, N=NY/n blocks.
// c2g is c^2 (the MPI solver does not square, MPIAcoustic.jl:336), global unpadded NX*NY.  sigma/tau are GLOBAL
// (NX+2)*(NY+2) profiles (MPIAcoustic.jl:177-185 evaluates the same pml_helper at global coordinates).
// ug: (NSTEP+1) * NX*NY global unpadded history.  Source / receiver indices are global, 1-based, unpadded.
// ------------------------------------------------------------------------------------------------
struct BlockScratch {
  std::vector<double> pw, pwold, pphi, ppsi, sg, tg, cb, o1, o2, o3, gw, gwold, gphi, gpsi, gc, i1, i2, i3;
};

// (n+2)^2 halo-padded copy of block (I,J) of a global unpadded field (zero outside the global grid)
static void pad_block(const double* f, int64 NX, int64 NY, int64 n, int64 I, int64 J, double* out) {
  for (int64 i = 0; i < n + 2; i++) {
    const int64 gi = I * n + i - 1;
    double* row = out + i * (n + 2);
    if (gi < 0 || gi >= NX) { memset(row, 0, sizeof(double) * (n + 2)); continue; }
    const int64 gj0 = J * n - 1;
    row[0] = (gj0 >= 0) ? f[gi * NY + gj0] : 0.0;
    memcpy(row + 1, f + gi * NY + J * n, sizeof(double) * n);
    row[n + 1] = (gj0 + n + 1 < NY) ? f[gi * NY + gj0 + n + 1] : 0.0;
  }
}
static void get_block(const double* f, int64 NY, int64 n, int64 I, int64 J, double* out) {
  for (int64 i = 0; i < n; i++) memcpy(out + i * n, f + (I * n + i) * NY + J * n, sizeof(double) * n);
}
static void put_block(double* f, int64 NY, int64 n, int64 I, int64 J, const double* in) {
  for (int64 i = 0; i < n; i++) memcpy(f + (I * n + i) * NY + J * n, in + i * n, sizeof(double) * n);
}

static void init_blocks(std::vector<BlockScratch>& S, int64 NX, int64 NY, int64 n, const double* sigma,
                        const double* tau, const double* c2g, bool backward) {
  const int64 Nb = NY / n, P = (n + 2) * (n + 2), B = (NX / n) * Nb;
  S.resize(B);
  for (int64 b = 0; b < B; b++) {
    const int64 I = b / Nb, J = b % Nb;
    BlockScratch& s = S[b];
    s.pw.resize(P); s.pwold.resize(P); s.pphi.resize(P); s.ppsi.resize(P); s.sg.resize(P); s.tg.resize(P);
    s.cb.resize(n * n); s.o1.resize(n * n); s.o2.resize(n * n); s.o3.resize(n * n);
    if (backward) {
      s.gw.resize(P); s.gwold.resize(P); s.gphi.resize(P); s.gpsi.resize(P); s.gc.resize(n * n);
      s.i1.resize(n * n); s.i2.resize(n * n); s.i3.resize(n * n);
    }
    get_block(c2g, NY, n, I, J, s.cb.data());
    for (int64 i = 0; i < n + 2; i++)
      for (int64 j = 0; j < n + 2; j++) {
        s.sg[i * (n + 2) + j] = sigma[(I * n + i) * (NY + 2) + J * n + j];
        s.tg[i * (n + 2) + j] = tau[(I * n + i) * (NY + 2) + J * n + j];
      }
  }
}

API void ref_drv_mpi_acoustic_forward(int64 NX, int64 NY, int64 n, int64 NSTEP, double dt, double hx, double hy,
                                      const double* sigma, const double* tau, const double* c2g, int64 nsrc,
                                      const int64* srci, const int64* srcj, const double* srcv, double* ug,
                                      int nthreads) {
  const int64 Nb = NY / n, B = (NX / n) * Nb, NG = NX * NY;
  std::vector<BlockScratch> S;
  init_blocks(S, NX, NY, n, sigma, tau, c2g, false);
  std::vector<double> phi[2] = {std::vector<double>(NG, 0.0), std::vector<double>(NG, 0.0)};
  std::vector<double> psi[2] = {std::vector<double>(NG, 0.0), std::vector<double>(NG, 0.0)};
  memset(ug, 0, sizeof(double) * NG * 2);
  if (nthreads < 1) nthreads = 1;
  for (int64 s = 2; s <= NSTEP; s++) {
    const double *w = ug + (s - 1) * NG, *wold = ug + (s - 2) * NG;
    double* un = ug + s * NG;
    const std::vector<double>&ph = phi[(s - 1) & 1], &ps = psi[(s - 1) & 1];
    std::vector<double>&pho = phi[s & 1], &pso = psi[s & 1];
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int64 b = 0; b < B; b++) {
      const int64 I = b / Nb, J = b % Nb;
      BlockScratch& k = S[b];
      pad_block(w, NX, NY, n, I, J, k.pw.data());       // 4 halo exchanges (MPIAcoustic.jl:260-263)
      pad_block(wold, NX, NY, n, I, J, k.pwold.data());
      pad_block(ph.data(), NX, NY, n, I, J, k.pphi.data());
      pad_block(ps.data(), NX, NY, n, I, J, k.ppsi.data());
      MpiAcousticOneStepCpuForward(k.pw.data(), k.pwold.data(), k.pphi.data(), k.ppsi.data(), k.sg.data(), k.tg.data(),
                                   k.cb.data(), dt, hx, hy, n, n, k.o1.data(), k.o2.data(), k.o3.data());
      for (int64 q = 0; q < nsrc; q++) {  // MPIAcoustic.jl:71-78, 377-381
        const int64 li = srci[q] - I * n, lj = srcj[q] - J * n;
        if (li >= 1 && li <= n && lj >= 1 && lj <= n) k.o1[(li - 1) * n + lj - 1] += srcv[(s - 1) * nsrc + q] * (dt * dt);
      }
      put_block(un, NY, n, I, J, k.o1.data());
      put_block(pho.data(), NY, n, I, J, k.o2.data());
      put_block(pso.data(), NY, n, I, J, k.o3.data());
    }
  }
}

// loss = sum (rcvv-obs)^2 over all blocks (mpi_sum of the local losses) and its gradient w.r.t. c^2 [NX*NY] and
// srcv [NSTEP*nsrc], composing MpiAcousticOneStepCpuBackward (MpiAcousticOneStep.h:45-122) per block with the
// transpose of the halo exchange (halo contributions are added into the neighbour's edge cells).  NOTE: the
// reference registers an EMPTY gradient op for this kernel (MpiAcousticOneStep.cpp:347-350); the body used here is
// the one it ships but never calls, so this driver is an optimistic stand-in for "the reference on all cores".
API void ref_drv_mpi_acoustic_gradient(int64 NX, int64 NY, int64 n, int64 NSTEP, double dt, double hx, double hy,
                                       const double* sigma, const double* tau, const double* c2g, int64 nsrc,
                                       const int64* srci, const int64* srcj, int64 nrcv, const int64* rcvi,
                                       const int64* rcvj, const double* obs, const double* ug, double* loss_out,
                                       double* grad_c2, double* grad_srcv, int nthreads) {
  const int64 Nb = NY / n, M = NX / n, B = M * Nb, NG = NX * NY, n2 = n + 2;
  std::vector<BlockScratch> S;
  init_blocks(S, NX, NY, n, sigma, tau, c2g, true);
  std::vector<double> ub[3] = {std::vector<double>(NG, 0.0), std::vector<double>(NG, 0.0), std::vector<double>(NG, 0.0)};
  std::vector<double> gphi[2] = {std::vector<double>(NG, 0.0), std::vector<double>(NG, 0.0)};
  std::vector<double> gpsi[2] = {std::vector<double>(NG, 0.0), std::vector<double>(NG, 0.0)};
  std::vector<double> add1(NG), add2(NG);
  memset(grad_c2, 0, sizeof(double) * NG);
  if (grad_srcv) memset(grad_srcv, 0, sizeof(double) * NSTEP * nsrc);
  if (nthreads < 1) nthreads = 1;
  double loss = 0.0;
  for (int64 s = 0; s <= NSTEP; s++)
    for (int64 r = 0; r < nrcv; r++) {
      const double d = ug[s * NG + (rcvi[r] - 1) * NY + rcvj[r] - 1] - obs[s * nrcv + r];
      loss += d * d;
    }
  *loss_out = loss;
  auto seed = [&](std::vector<double>& bfr, int64 s) {
    std::fill(bfr.begin(), bfr.end(), 0.0);
    for (int64 r = 0; r < nrcv; r++) {
      const int64 id = (rcvi[r] - 1) * NY + rcvj[r] - 1;
      bfr[id] += 2.0 * (ug[s * NG + id] - obs[s * nrcv + r]);
    }
  };
  seed(ub[NSTEP % 3], NSTEP);
  seed(ub[(NSTEP - 1) % 3], NSTEP - 1);
  for (int64 s = NSTEP; s >= 2; s--) {
    std::vector<double>& gu = ub[s % 3];
    seed(ub[(s - 2) % 3], s - 2);
    if (grad_srcv)
      for (int64 q = 0; q < nsrc; q++) grad_srcv[(s - 1) * nsrc + q] = gu[(srci[q] - 1) * NY + srcj[q] - 1] * (dt * dt);
    const double* w = ug + (s - 1) * NG;
    const std::vector<double>&gpo = gphi[s & 1], &gso = gpsi[s & 1];
    std::vector<double>&gpn = gphi[(s - 1) & 1], &gsn = gpsi[(s - 1) & 1];
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int64 b = 0; b < B; b++) {
      const int64 I = b / Nb, J = b % Nb;
      BlockScratch& k = S[b];
      pad_block(w, NX, NY, n, I, J, k.pw.data());
      get_block(gu.data(), NY, n, I, J, k.i1.data());
      get_block(gpo.data(), NY, n, I, J, k.i2.data());
      get_block(gso.data(), NY, n, I, J, k.i3.data());
      std::fill(k.gw.begin(), k.gw.end(), 0.0); std::fill(k.gwold.begin(), k.gwold.end(), 0.0);
      std::fill(k.gphi.begin(), k.gphi.end(), 0.0); std::fill(k.gpsi.begin(), k.gpsi.end(), 0.0);
      std::fill(k.gc.begin(), k.gc.end(), 0.0);
      MpiAcousticOneStepCpuBackward(k.gw.data(), k.gwold.data(), k.gphi.data(), k.gpsi.data(), k.gc.data(), k.i1.data(),
                                    k.i2.data(), k.i3.data(), k.pw.data(), nullptr, nullptr, nullptr, k.sg.data(),
                                    k.tg.data(), k.cb.data(), dt, hx, hy, n, n, nullptr, nullptr, nullptr);
    }
    // transpose of the halo exchanges: own interior + the neighbours' halo cells that mirror my edge cells
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int64 b = 0; b < B; b++) {
      const int64 I = b / Nb, J = b % Nb;
      auto gather = [&](std::vector<double> BlockScratch::*fld, double* out_nn) {
        const std::vector<double>& me = S[b].*fld;
        for (int64 i = 0; i < n; i++)
          for (int64 j = 0; j < n; j++) out_nn[i * n + j] = me[(i + 1) * n2 + j + 1];
        if (I > 0) { const std::vector<double>& o = S[b - Nb].*fld; for (int64 j = 0; j < n; j++) out_nn[j] += o[(n + 1) * n2 + j + 1]; }
        if (I < M - 1) { const std::vector<double>& o = S[b + Nb].*fld; for (int64 j = 0; j < n; j++) out_nn[(n - 1) * n + j] += o[j + 1]; }
        if (J > 0) { const std::vector<double>& o = S[b - 1].*fld; for (int64 i = 0; i < n; i++) out_nn[i * n] += o[(i + 1) * n2 + n + 1]; }
        if (J < Nb - 1) { const std::vector<double>& o = S[b + 1].*fld; for (int64 i = 0; i < n; i++) out_nn[i * n + n - 1] += o[(i + 1) * n2]; }
      };
      BlockScratch& k = S[b];
      gather(&BlockScratch::gw, k.o1.data());
      put_block(add1.data(), NY, n, I, J, k.o1.data());
      gather(&BlockScratch::gwold, k.o1.data());
      put_block(add2.data(), NY, n, I, J, k.o1.data());
      gather(&BlockScratch::gphi, k.o2.data());
      put_block(gpn.data(), NY, n, I, J, k.o2.data());
      gather(&BlockScratch::gpsi, k.o3.data());
      put_block(gsn.data(), NY, n, I, J, k.o3.data());
      for (int64 i = 0; i < n; i++)
        for (int64 j = 0; j < n; j++) grad_c2[(I * n + i) * NY + J * n + j] += k.gc[i * n + j];
    }
    std::vector<double>&g1 = ub[(s - 1) % 3], &g2 = ub[(s - 2) % 3];
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64 q = 0; q < NG; q++) { g1[q] += add1[q]; g2[q] += add2[q]; }
  }
}

API int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#include "ref_graph.inc"
