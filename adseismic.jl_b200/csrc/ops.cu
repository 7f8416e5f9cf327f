// ops.cu -- op-level entry points with the reference ops' own argument lists (device pointers, dense arrays).
// These are what a maintainer binds in place of `load_op_and_grad(libADSeismic, "acoustic_one_step")` etc. when
// keeping the reference's per-step graph (src/Core.jl:475-502, :215-228, :701-712).  The whole-loop entry points
// in acoustic.cu / elastic.cu are the fast path; these exist for drop-in parity at the op boundary.
#include "common.cuh"

// Forward: AcousticOneStepCpu.h:1-48 (sigma/tau are full arrays here, as in the op signature).
__global__ void k_op_ac_fwd(const double* __restrict__ w, const double* __restrict__ wold,
                            const double* __restrict__ phi, const double* __restrict__ psi,
                            const double* __restrict__ sigma, const double* __restrict__ tau,
                            const double* __restrict__ c, double dt, double hx, double hy, int NX, int NY,
                            double* __restrict__ u, double* __restrict__ phiout, double* __restrict__ psiout) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  const int W = NY + 2;
  if (j >= W) return;
  const i64 IJ = (i64)i * W + j;
  if (i == 0 || i == NX + 1 || j == 0 || j == NY + 1) {
    u[IJ] = 0.0; phiout[IJ] = 0.0; psiout[IJ] = 0.0;
    return;
  }
  const i64 IpJ = IJ + W, InJ = IJ - W, IJp = IJ + 1, IJn = IJ - 1;
  double v = (2 - sigma[IJ] * tau[IJ] * dt * dt - 2 * dt * dt / hx / hx * c[IJ] - 2 * dt * dt / hy / hy * c[IJ]) * w[IJ] +
             c[IJ] * (dt / hx) * (dt / hx) * (w[IpJ] + w[InJ]) +
             c[IJ] * (dt / hy) * (dt / hy) * (w[IJp] + w[IJn]) +
             (dt * dt / (2.0 * hx)) * (phi[IpJ] - phi[InJ]) +
             (dt * dt / (2.0 * hy)) * (psi[IJp] - psi[IJn]) -
             (1 - (sigma[IJ] + tau[IJ]) * dt / 2) * wold[IJ];
  u[IJ] = v / (1 + (sigma[IJ] + tau[IJ]) / 2 * dt);
  phiout[IJ] = (1. - dt * sigma[IJ]) * phi[IJ] + dt * c[IJ] * (tau[IJ] - sigma[IJ]) / 2.0 / hx * (w[IpJ] - w[InJ]);
  psiout[IJ] = (1. - dt * tau[IJ]) * psi[IJ] + dt * c[IJ] * (sigma[IJ] - tau[IJ]) / 2.0 / hy * (w[IJp] - w[IJn]);
}

// Backward in gather form (transpose of AcousticOneStepCpu.h:76-124): every output cell is written exactly once.
__global__ void k_op_ac_bwd(double* __restrict__ gw, double* __restrict__ gwold, double* __restrict__ gphi,
                            double* __restrict__ gpsi, double* __restrict__ gc, const double* __restrict__ gu,
                            const double* __restrict__ gphiout, const double* __restrict__ gpsiout,
                            const double* __restrict__ w, const double* __restrict__ sigma,
                            const double* __restrict__ tau, const double* __restrict__ c, double dt, double hx,
                            double hy, int NX, int NY) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  const int W = NY + 2;
  if (j >= W) return;
  const i64 IJ = (i64)i * W + j;
  auto interior = [&](int a, int b) { return a >= 1 && a <= NX && b >= 1 && b <= NY; };
  auto gdiv = [&](i64 Q) { return gu[Q] / (1 + (sigma[Q] + tau[Q]) / 2 * dt); };
  double aw = 0.0, awold = 0.0, aphi = 0.0, apsi = 0.0, ac = 0.0;
  if (interior(i, j)) {
    const double g = gdiv(IJ);
    ac = ((-2 * dt * dt / hx / hx - 2 * dt * dt / hy / hy) * w[IJ] + (dt / hx) * (dt / hx) * (w[IJ + W] + w[IJ - W]) +
          (dt / hy) * (dt / hy) * (w[IJ + 1] + w[IJ - 1])) * g +
         dt * (tau[IJ] - sigma[IJ]) / 2.0 / hx * (w[IJ + W] - w[IJ - W]) * gphiout[IJ] +
         dt * (sigma[IJ] - tau[IJ]) / 2.0 / hy * (w[IJ + 1] - w[IJ - 1]) * gpsiout[IJ];
    aw = (2 - sigma[IJ] * tau[IJ] * dt * dt - 2 * dt * dt / hx / hx * c[IJ] - 2 * dt * dt / hy / hy * c[IJ]) * g;
    awold = -(1 - (sigma[IJ] + tau[IJ]) * dt / 2) * g;
    aphi = (1. - dt * sigma[IJ]) * gphiout[IJ];
    apsi = (1. - dt * tau[IJ]) * gpsiout[IJ];
  }
  if (interior(i - 1, j)) {  // this cell is IpJ of Q
    const i64 Q = IJ - W;
    const double g = gdiv(Q);
    aw += c[Q] * (dt / hx) * (dt / hx) * g + dt * c[Q] * (tau[Q] - sigma[Q]) / 2.0 / hx * gphiout[Q];
    aphi += (dt * dt / (2.0 * hx)) * g;
  }
  if (interior(i + 1, j)) {  // InJ of Q
    const i64 Q = IJ + W;
    const double g = gdiv(Q);
    aw += c[Q] * (dt / hx) * (dt / hx) * g - dt * c[Q] * (tau[Q] - sigma[Q]) / 2.0 / hx * gphiout[Q];
    aphi += -(dt * dt / (2.0 * hx)) * g;
  }
  if (interior(i, j - 1)) {  // IJp of Q
    const i64 Q = IJ - 1;
    const double g = gdiv(Q);
    aw += c[Q] * (dt / hy) * (dt / hy) * g + dt * c[Q] * (sigma[Q] - tau[Q]) / 2.0 / hy * gpsiout[Q];
    apsi += (dt * dt / (2.0 * hy)) * g;
  }
  if (interior(i, j + 1)) {  // IJn of Q
    const i64 Q = IJ + 1;
    const double g = gdiv(Q);
    aw += c[Q] * (dt / hy) * (dt / hy) * g - dt * c[Q] * (sigma[Q] - tau[Q]) / 2.0 / hy * gpsiout[Q];
    apsi += -(dt * dt / (2.0 * hy)) * g;
  }
  gw[IJ] = aw; gwold[IJ] = awold; gphi[IJ] = aphi; gpsi[IJ] = apsi; gc[IJ] = ac;
}

ADSEIS_API int adseis_op_acoustic_step_fwd(adseis_ctx* ctx, const double* w, const double* wold, const double* phi,
                                           const double* psi, const double* sigma, const double* tau, const double* c,
                                           double dt, double hx, double hy, int64_t NX, int64_t NY, double* u,
                                           double* phiout, double* psiout, void* stream) {
  REQUIRE(ctx && w && wold && phi && psi && sigma && tau && c && u && phiout && psiout, "op_acoustic_step_fwd: null");
  REQUIRE(NX >= 1 && NY >= 1 && (NX + 2) * (NY + 2) < 2147483647LL, "op_acoustic_step_fwd: bad sizes");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  dim3 grid((unsigned)((NY + 2 + 127) / 128), (unsigned)(NX + 2));
  k_op_ac_fwd<<<grid, 128, 0, st>>>(w, wold, phi, psi, sigma, tau, c, dt, hx, hy, (int)NX, (int)NY, u, phiout, psiout);
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return ADSEIS_OK;
}

ADSEIS_API int adseis_op_acoustic_step_bwd(adseis_ctx* ctx, double* grad_w, double* grad_wold, double* grad_phi,
                                           double* grad_psi, double* grad_c, const double* grad_u,
                                           const double* grad_phiout, const double* grad_psiout, const double* w,
                                           const double* sigma, const double* tau, const double* c, double dt,
                                           double hx, double hy, int64_t NX, int64_t NY, void* stream) {
  REQUIRE(ctx && grad_w && grad_wold && grad_phi && grad_psi && grad_c && grad_u && grad_phiout && grad_psiout && w &&
              sigma && tau && c, "op_acoustic_step_bwd: null");
  REQUIRE(NX >= 1 && NY >= 1 && (NX + 2) * (NY + 2) < 2147483647LL, "op_acoustic_step_bwd: bad sizes");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  dim3 grid((unsigned)((NY + 2 + 127) / 128), (unsigned)(NX + 2));
  k_op_ac_bwd<<<grid, 128, 0, st>>>(grad_w, grad_wold, grad_phi, grad_psi, grad_c, grad_u, grad_phiout, grad_psiout, w,
                                    sigma, tau, c, dt, hx, hy, (int)NX, (int)NY);
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return ADSEIS_OK;
}

// ---- AddSource (SourceOps/AddSource.cpp:33-87): copy 5 fields, then field[type](srci-1, srcj-1) += srcv ----
__global__ void k_op_copy5(double* __restrict__ a_, double* __restrict__ b_, double* __restrict__ c_,
                           double* __restrict__ d_, double* __restrict__ e_, const double* __restrict__ a,
                           const double* __restrict__ b, const double* __restrict__ c, const double* __restrict__ d,
                           const double* __restrict__ e, i64 n) {
  i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) { a_[k] = a[k]; b_[k] = b[k]; c_[k] = c[k]; d_[k] = d[k]; e_[k] = e[k]; }
}
// one thread walks the sources in order: duplicates on one cell accumulate sequentially, as on the CPU.
// (nsrc is small in every reference use; the fused elastic kernels in elastic.cu do this per tile instead.)
__global__ void k_op_add_source(double* vx_, double* vy_, double* sxx_, double* syy_, double* sxy_,
                                const int64_t* __restrict__ srci, const int64_t* __restrict__ srcj,
                                const double* __restrict__ srcv, const int64_t* __restrict__ srctype, i64 nsrc,
                                int NY) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (i64 k = 0; k < nsrc; k++) {
    const i64 id = (srci[k] - 1) * (NY + 2) + (srcj[k] - 1);
    switch (srctype[k]) {
      case 0: vx_[id] += srcv[k]; break;
      case 1: vy_[id] += srcv[k]; break;
      case 2: sxx_[id] += srcv[k]; break;
      case 3: syy_[id] += srcv[k]; break;
      case 4: sxy_[id] += srcv[k]; break;
      default: break;
    }
  }
}

ADSEIS_API int adseis_op_add_source_fwd(adseis_ctx* ctx, double* vx_, double* vy_, double* sxx_, double* syy_,
                                        double* sxy_, const double* vx, const double* vy, const double* sxx,
                                        const double* syy, const double* sxy, const int64_t* srci,
                                        const int64_t* srcj, const double* srcv, const int64_t* srctype, int64_t nsrc,
                                        int64_t NX, int64_t NY, void* stream) {
  REQUIRE(ctx && vx_ && vy_ && sxx_ && syy_ && sxy_ && vx && vy && sxx && syy && sxy, "op_add_source_fwd: null");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const i64 n = (NX + 2) * (NY + 2);
  k_op_copy5<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(vx_, vy_, sxx_, syy_, sxy_, vx, vy, sxx, syy, sxy, n);
  ctx->launches++;
  if (nsrc > 0) {
    k_op_add_source<<<1, 32, 0, st>>>(vx_, vy_, sxx_, syy_, sxy_, srci, srcj, srcv, srctype, nsrc, (int)NY);
    ctx->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return ADSEIS_OK;
}

// ---- GetReceive (ReceiveOps/GetReceive.cpp:10-46): out[nt*i + k] = field[type_i][k*N + idx_i] ----
__global__ void k_op_get_receive(double* __restrict__ out, const double* __restrict__ vx, const double* __restrict__ vy,
                                 const double* __restrict__ sxx, const double* __restrict__ syy,
                                 const double* __restrict__ sxy, i64 nt, const int64_t* __restrict__ rcvi,
                                 const int64_t* __restrict__ rcvj, const int64_t* __restrict__ rcvtype, i64 nrcv,
                                 i64 N, int NY) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;  // time index (fast in `out`)
  const i64 i = blockIdx.y;
  if (k >= nt || i >= nrcv) return;
  const i64 idx = (rcvi[i] - 1) * (NY + 2) + rcvj[i] - 1;
  const double* f = nullptr;
  switch (rcvtype[i]) {
    case 0: f = vx; break;
    case 1: f = vy; break;
    case 2: f = sxx; break;
    case 3: f = syy; break;
    case 4: f = sxy; break;
    default: return;
  }
  out[nt * i + k] = f[k * N + idx];
}

ADSEIS_API int adseis_op_get_receive_fwd(adseis_ctx* ctx, double* out, const double* vx, const double* vy,
                                         const double* sxx, const double* syy, const double* sxy, int64_t nt,
                                         const int64_t* rcvi, const int64_t* rcvj, const int64_t* rcvtype,
                                         int64_t nrcv, int64_t NX, int64_t NY, void* stream) {
  REQUIRE(ctx && out && vx && vy && sxx && syy && sxy, "op_get_receive_fwd: null");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  if (nrcv > 0 && nt > 0) {
    dim3 grid((unsigned)((nt + 127) / 128), (unsigned)nrcv);
    k_op_get_receive<<<grid, 128, 0, st>>>(out, vx, vy, sxx, syy, sxy, nt, rcvi, rcvj, rcvtype, nrcv,
                                           (NX + 2) * (NY + 2), (int)NY);
    ctx->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return ADSEIS_OK;
}
