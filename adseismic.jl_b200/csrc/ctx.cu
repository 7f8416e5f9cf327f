// ctx.cu -- context, error text, host-only helpers (PML / CPML profiles, slab partition).
#include <math.h>
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void adseis_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

ADSEIS_API int adseis_version(void) { return 100; }
ADSEIS_API const char* adseis_last_error(void) { return g_err; }

ADSEIS_API int adseis_device_count(int* n) {
  REQUIRE(n, "adseis_device_count: null");
  *n = 0;
  CUDA_TRY(cudaGetDeviceCount(n));
  return ADSEIS_OK;
}

ADSEIS_API int adseis_ctx_create(int device, adseis_ctx** out) {
  REQUIRE(out, "adseis_ctx_create: null out");
  *out = nullptr;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev == 0) {
    adseis_set_error("adseis_ctx_create: no CUDA device (this library has no CPU fallback)");
    return ADSEIS_ECUDA;
  }
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  REQUIRE(device < ndev, "adseis_ctx_create: device %d out of range (%d devices)", device, ndev);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    adseis_set_error("adseis_ctx_create: device %d is sm_%d%d; this build targets sm_100a only", device, prop.major,
                     prop.minor);
    return ADSEIS_ECUDA;
  }
  adseis_ctx* c = new adseis_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->launches = 0;
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&c->ev0));
  CUDA_TRY(cudaEventCreate(&c->ev1));
  *out = c;
  return ADSEIS_OK;
}

static void ctx_free(adseis_ctx* ctx) {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

// A context outlives its plans: destroying it while plans are alive (garbage-collected host languages finalise in
// arbitrary order) only marks it; the last plan's destructor frees it.
ADSEIS_API int adseis_ctx_destroy(adseis_ctx* ctx) {
  if (!ctx) return ADSEIS_OK;
  if (ctx->plans > 0) { ctx->zombie = true; return ADSEIS_OK; }
  ctx_free(ctx);
  return ADSEIS_OK;
}

void adseis_ctx_release_plan(adseis_ctx* ctx) {
  if (!ctx) return;
  if (--ctx->plans <= 0 && ctx->zombie) ctx_free(ctx);
}

ADSEIS_API int adseis_ctx_sync(adseis_ctx* ctx) {
  REQUIRE(ctx, "adseis_ctx_sync: null ctx");
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(cudaGetLastError());
  return ADSEIS_OK;
}

ADSEIS_API int adseis_ctx_stream(adseis_ctx* ctx, void** stream) {
  REQUIRE(ctx && stream, "adseis_ctx_stream: null");
  *stream = (void*)ctx->stream;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_ctx_launch_count(adseis_ctx* ctx, int64_t* n) {
  REQUIRE(ctx && n, "adseis_ctx_launch_count: null");
  *n = ctx->launches;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_ctx_timer_start(adseis_ctx* ctx) {
  REQUIRE(ctx, "null ctx");
  CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
  return ADSEIS_OK;
}

ADSEIS_API int adseis_ctx_timer_stop_ms(adseis_ctx* ctx, double* ms) {
  REQUIRE(ctx && ms, "null");
  CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
  CUDA_TRY(cudaEventSynchronize(ctx->ev1));
  float f = 0;
  CUDA_TRY(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = f;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_ctx_mem_info(adseis_ctx* ctx, size_t* free_bytes, size_t* total_bytes) {
  REQUIRE(ctx && free_bytes && total_bytes, "null");
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaMemGetInfo(free_bytes, total_bytes));
  return ADSEIS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Acoustic PML profile: compute_PML_Params!/pml_helper (src/Core.jl:622-655, 738-761; MPI twin
// src/MPIAcoustic.jl:128-185 evaluates the same function at global coordinates X = i*DELTAX).
// The profile is separable: Sigma_x[i,j] = sigx[i], Sigma_y[i,j] = tauy[j].
// ---------------------------------------------------------------------------------------------------------
static double pml_1d(double x, double h, i64 n, i64 npml, double xi, int use_min, int use_max) {
  const double L = (double)npml * h;
  double out = 0.0;
  if (x < L && use_min) {
    double d = fabs(L - x);
    out = xi * (d / L - sin(2.0 * M_PI * d / L) / (2.0 * M_PI));
  } else if (x > h * (double)(n + 1) - L && use_max) {
    double d = fabs(x - (h * (double)(n + 1) - L));
    out = xi * (d / L - sin(2.0 * M_PI * d / L) / (2.0 * M_PI));
  }
  return out;
}

ADSEIS_API int adseis_acoustic_pml_profiles(const adseis_acoustic_params* p, double* sigx, double* tauy) {
  REQUIRE(p && sigx && tauy, "adseis_acoustic_pml_profiles: null");
  REQUIRE(p->NX > 0 && p->NY > 0 && p->NPOINTS_PML > 0, "adseis_acoustic_pml_profiles: bad sizes");
  const double Lx = (double)p->NPOINTS_PML * p->DELTAX, Ly = (double)p->NPOINTS_PML * p->DELTAY;
  const double xix = p->vp_ref / Lx * log(1.0 / p->Rcoef);
  const double xiy = p->vp_ref / Ly * log(1.0 / p->Rcoef);
  for (i64 i = 0; i < p->NX + 2; i++)
    sigx[i] = pml_1d((double)i * p->DELTAX, p->DELTAX, p->NX, p->NPOINTS_PML, xix, p->USE_PML_XMIN, p->USE_PML_XMAX);
  for (i64 j = 0; j < p->NY + 2; j++)
    tauy[j] = pml_1d((double)j * p->DELTAY, p->DELTAY, p->NY, p->NPOINTS_PML, xiy, p->USE_PML_YMIN, p->USE_PML_YMAX);
  return ADSEIS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Elastic CPML coefficients (Komatitsch-Martin as modified by the reference: alpha = alpha_max * xi,
// src/Core.jl:289): src/Core.jl:231-407, src/MPIElastic.jl:191-372 (global index, II=JJ=1).
// ---------------------------------------------------------------------------------------------------------
ADSEIS_API int adseis_elastic_cpml_profiles(const adseis_elastic_params* p, int axis, double* a, double* b) {
  REQUIRE(p && a && b && (axis == 0 || axis == 1), "adseis_elastic_cpml_profiles: bad argument");
  REQUIRE(p->K_MAX_PML == 1.0, "K_MAX_PML must be 1 (the reference ignores K, Core.jl:686-693)");
  const i64 n = axis == 0 ? p->NX : p->NY;
  const double h = axis == 0 ? p->DELTAX : p->DELTAY;
  const int use_min = axis == 0 ? p->USE_PML_XMIN : p->USE_PML_YMIN;
  const int use_max = axis == 0 ? p->USE_PML_XMAX : p->USE_PML_YMAX;
  const double thick = (double)p->NPOINTS_PML * h;
  const double d0 = -(p->NPOWER + 1.0) * p->vp_ref * log(p->Rcoef) / (2.0 * thick);
  const double oleft = thick, oright = ((double)n - 0.5) * h - thick;
  for (i64 i = 0; i < n; i++) {
    double d = 0, dh = 0, K = 1, Kh = 1, al = 0, alh = 0;
    const double x = h * (double)i;
    if (use_min) {
      double ab = oleft - x;
      if (ab >= 0.0) { double an = ab / thick; d = d0 * pow(an, p->NPOWER); K = 1.0 + (p->K_MAX_PML - 1.0) * pow(an, p->NPOWER); al = p->ALPHA_MAX_PML * an; }
      ab = oleft - (x + h / 2.0);
      if (ab >= 0.0) { double an = ab / thick; dh = d0 * pow(an, p->NPOWER); Kh = 1.0 + (p->K_MAX_PML - 1.0) * pow(an, p->NPOWER); alh = p->ALPHA_MAX_PML * an; }
    }
    if (use_max) {
      double ab = x - oright;
      if (ab >= 0.0) { double an = ab / thick; d = d0 * pow(an, p->NPOWER); K = 1.0 + (p->K_MAX_PML - 1.0) * pow(an, p->NPOWER); al = p->ALPHA_MAX_PML * an; }
      ab = x + h / 2.0 - oright;
      if (ab >= 0.0) { double an = ab / thick; dh = d0 * pow(an, p->NPOWER); Kh = 1.0 + (p->K_MAX_PML - 1.0) * pow(an, p->NPOWER); alh = p->ALPHA_MAX_PML * an; }
    }
    if (al < 0) al = 0;
    if (alh < 0) alh = 0;
    b[i] = exp(-(d / K + al) * p->DELTAT);
    b[n + i] = exp(-(dh / Kh + alh) * p->DELTAT);
    a[i] = 0.0;
    a[n + i] = 0.0;
    if (fabs(d) > 1e-6) a[i] = d * (b[i] - 1.0) / (K * (d + K * al));
    if (fabs(dh) > 1e-6) a[n + i] = dh * (b[n + i] - 1.0) / (Kh * (dh + Kh * alh));
  }
  return ADSEIS_OK;
}

// Balanced 1-D slab partition of the padded rows [0, NX+2): interior rows 1..NX are split as evenly as possible,
// the two ring rows go to the first / last rank.
ADSEIS_API int adseis_slab_partition(int64_t NX, int32_t nranks, int32_t rank, adseis_slab* out) {
  REQUIRE(out && nranks >= 1 && rank >= 0 && rank < nranks && NX >= nranks, "adseis_slab_partition: bad argument");
  i64 base = NX / nranks, rem = NX % nranks;
  i64 start = 1 + rank * base + (rank < rem ? rank : rem);
  i64 cnt = base + (rank < rem ? 1 : 0);
  out->rank = rank;
  out->nranks = nranks;
  out->row0 = (rank == 0) ? 0 : start;
  out->row1 = (rank == nranks - 1) ? NX + 2 : start + cnt;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_slab_partition(const adseis_elastic_params* p, int32_t nranks, int32_t rank,
                                             adseis_slab* out) {
  REQUIRE(p && out && nranks >= 1 && rank >= 0 && rank < nranks && p->NX >= 4 * (i64)nranks,
          "adseis_elastic_slab_partition: bad argument");
  const i64 ghost = p->variant == 0 ? 1 : 2, NX = p->NX;
  i64 base = NX / nranks, rem = NX % nranks;
  i64 start = ghost + rank * base + (rank < rem ? rank : rem);
  i64 cnt = base + (rank < rem ? 1 : 0);
  out->rank = rank;
  out->nranks = nranks;
  out->row0 = (rank == 0) ? 0 : start;
  out->row1 = (rank == nranks - 1) ? NX + 2 * ghost : start + cnt;
  return ADSEIS_OK;
}
