// elastic.cu -- elastic plan: device state, time loops (forward, checkpointed reverse sweep), C ABI.
// Reference behaviour restated: src/Core.jl:31-228 (variant 0), src/MPIElastic.jl:374-682 on the global grid
// (variant 1), CPML coefficients src/Core.jl:231-407, receivers src/Core.jl:701-712 + ReceiveOps/GetReceive.cpp,
// sources SourceOps/AddSource.cpp, misfit src/Utils.jl:308; adjoint = SURVEY Appendix B.
#include <algorithm>
#include <map>

#include "elastic_kernels.cuh"
#include "util_kernels.cuh"

// ------------------------------------------------------------------------------------------------------------
// material pre-processing and gradient post-processing
// ------------------------------------------------------------------------------------------------------------
// Averaged materials exactly as fw1/fw2/fw4 form them (Core.jl:109-111, 141, 200); variant M: no averaging.
__global__ void k_el_materials(int Hl, int ld, int W, int goff, int H, int avg, const double* __restrict__ rho,
                               const double* __restrict__ lam, const double* __restrict__ mu,
                               double* __restrict__ lamb, double* __restrict__ lmb, double* __restrict__ mub2,
                               double* __restrict__ rhob, double* __restrict__ rinv, double* __restrict__ rbinv) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int li = blockIdx.y;
  if (q >= ld || li >= Hl) return;
  const i64 c = (i64)li * ld + q;
  if (q >= W) { lamb[c] = lmb[c] = mub2[c] = rhob[c] = rinv[c] = rbinv[c] = 0.0; return; }
  double l_, m1, m2, rb;
  if (avg) {
    const bool dn = goff + li + 1 < H, rt = q + 1 < W;  // raw planes hold Hl+1 local rows: row li+1 is always there
    l_ = dn ? 0.5 * (lam[c + ld] + lam[c]) : lam[c];
    m1 = dn ? 0.5 * (mu[c + ld] + mu[c]) : mu[c];
    m2 = rt ? 0.5 * (mu[c] + mu[c + 1]) : mu[c];
    rb = (dn && rt) ? 0.25 * (rho[c] + rho[c + ld] + rho[c + ld + 1] + rho[c + 1]) : rho[c];
  } else {
    l_ = lam[c]; m1 = mu[c]; m2 = mu[c]; rb = rho[c];
  }
  lamb[c] = l_;
  lmb[c] = l_ + 2 * m1;
  mub2[c] = m2;
  rhob[c] = rb;
  rinv[c] = rho[c] != 0.0 ? 1.0 / rho[c] : 0.0;
  rbinv[c] = rb != 0.0 ? 1.0 / rb : 0.0;
}

// d loss / d (rho, lambda, mu) in the caller's layout from the accumulators w.r.t. the averaged materials.
__global__ void k_el_grad_finalize(ElGeom g, int avg, int off /* rows/cols to strip: 0 (S) or 2 (M) */,
                                   const double* __restrict__ Gl, const double* __restrict__ Gm1,
                                   const double* __restrict__ Gm2, const double* __restrict__ Gr3,
                                   const double* __restrict__ Gr4, double* __restrict__ grho,
                                   double* __restrict__ glam, double* __restrict__ gmu) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int li = g.own0 + blockIdx.y;
  if (q >= g.W || li >= g.own1) return;
  const int gp = g.goff + li;
  const i64 c = (i64)li * g.ld + q;
  double gl, gm, gr;
  if (avg) {
    const bool up = gp - 1 >= 0, lf = q - 1 >= 0;  // slab plans: row li-1 of an accumulator may be a halo row
    gl = 0.5 * (Gl[c] + (up ? Gl[c - g.ld] : 0.0));
    gm = 0.5 * (Gm1[c] + (up ? Gm1[c - g.ld] : 0.0)) + 0.5 * (Gm2[c] + (lf ? Gm2[c - 1] : 0.0));
    gr = Gr3[c] + 0.25 * (Gr4[c] + (up ? Gr4[c - g.ld] : 0.0) + ((up && lf) ? Gr4[c - g.ld - 1] : 0.0) +
                          (lf ? Gr4[c - 1] : 0.0));
  } else {
    gl = Gl[c]; gm = Gm1[c] + Gm2[c]; gr = Gr3[c] + Gr4[c];
  }
  const int op = gp - off, oq = q - off, OW = g.W - 2 * off, OH = g.H - 2 * off;
  if (op < 0 || op >= OH || oq < 0 || oq >= OW) return;
  const i64 o = (i64)op * OW + oq;
  grho[o] = gr; glam[o] = gl; gmu[o] = gm;
}

// post-injection value of a stress plane for snapshot output: plane += pending stress sources of that slot
__global__ void k_el_apply_stress_sources(double* __restrict__ plane, int field, ElPoints src, int nu,
                                          const double* __restrict__ srcv_row) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nu && src.field[k] == field) {
    double v = plane[src.cell[k]];
    for (int m = src.start[k]; m < src.start[k + 1]; m++) v += srcv_row[src.perm[m]];
    plane[src.cell[k]] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------------------
struct ElPointStore {
  PointSetStorage ps;
  int *xstart = nullptr, *xperm = nullptr;
  ElPoints dev{};
};

struct adseis_elastic_plan {
  adseis_ctx* ctx;
  adseis_elastic_params p;
  adseis_slab slab;
  ElGeom g;
  int nblocks = 0, nmarch = 0;
  ElCta* ctas = nullptr;
  int box[4] = {0, 0, 0, 0};  // marching box: local rows [box0,box1), columns [box2,box3)
  int off = 0;          // 0 (S) / 2 (M): offset between the caller's model array and the internal array
  i64 model_elems = 0;
  i64 slot_sz = 0;      // doubles per slot
  // materials
  double *rho = nullptr, *lam = nullptr, *mu = nullptr, *lamb = nullptr, *lmb = nullptr, *mub2 = nullptr,
         *rhob = nullptr, *rinv = nullptr, *rbinv = nullptr;
  double *ax = nullptr, *bx = nullptr, *ay = nullptr, *by = nullptr;
  // history window + checkpoints
  double* hist = nullptr;
  i64 win = 0;
  std::vector<i64> seg_b, seg_e;
  std::vector<double*> ckpt;
  i64 win_base = -1, win_last = -1;
  bool inplace_last = false;  // last forward ran in place (only the final slot is resident)
  // points
  i64 nsrc = 0, nrcv = 0;
  ElPointStore src, rcv;
  unsigned char* rcv_owned = nullptr;
  double *srcv = nullptr, *rcvv = nullptr, *obs = nullptr, *res = nullptr, *loss = nullptr;
  // adjoint state: fields in place + two memory sides
  double* adj = nullptr;       // 5 planes + 2 * (4 xm + 4 ym)
  double *Gl = nullptr, *Gm1 = nullptr, *Gm2 = nullptr, *Gr3 = nullptr, *Gr4 = nullptr;
  double *grho = nullptr, *glam = nullptr, *gmu = nullptr, *gradsrcv = nullptr;
  bool have_model = false, have_srcv = false, have_obs = false, have_fwd = false, have_grad = false,
       have_matgrad = false;
  i64 last_launches = 0, last_segments = 0, last_recomputed = 0;
  // slab decomposition (nranks > 1): one IPC-exported arena holds every array with halo rows
  double* arena = nullptr;
  size_t arena_bytes = 0;
  struct Desc {
    unsigned long long magic;
    long long Hl, ld, plane, slot_sz, win, own0, own1;
    long long off_flags, off_hist, off_adj, off_gacc;
    long long n_edge_lo, n_edge_hi;
    long long off_ll;   // packed halo rows: [from lo, from hi][parity][field 0..2][row 0..1] rows of ld 16-byte words (0: absent)
  } desc{}, dpeer[2];
  char* peer[2] = {nullptr, nullptr};
  unsigned long long epoch = 0, sepoch = 0;
  bool ll = false;              // packed halo rows instead of fence + flag (ADSEIS_EL_LL=0 switches back)
  int ll_prev_nf = 0;           // planes the previous fused launch produced (= the planes whose halo rows the next one reads)
  struct { int arr; long long idx; int field; } ll_prev[3];
  int* perm = nullptr;
  int n_edge_lo = 0, n_edge_hi = 0;
  bool connected = false;
};
#define EL_DESC_MAGIC 0xAD5E15E1A5ULL
#define EL_HX_BLOCKS 8

#define EL_LAUNCH_CHECK(P)        \
  do {                            \
    (P)->ctx->launches++;         \
    (P)->last_launches++;         \
    CUDA_TRY(cudaGetLastError()); \
  } while (0)

static ElSlot slot_at(const adseis_elastic_plan* P, double* base) {
  const ElGeom& g = P->g;
  ElSlot s;
  s.vx = base; s.vy = base + g.plane; s.sxx = base + 2 * g.plane; s.syy = base + 3 * g.plane;
  s.sxy = base + 4 * g.plane; s.xm = base + 5 * g.plane; s.ym = s.xm + 4 * g.xm_sz;
  return s;
}
static inline double* win_ptr(adseis_elastic_plan* P, i64 base, i64 s) { return P->hist + (s - base) * P->slot_sz; }
static ElMat mat_of(const adseis_elastic_plan* P) {
  return ElMat{P->lamb, P->lmb, P->mub2, P->rho, P->rhob, P->rinv, P->rbinv};
}
static ElCoef coef_of(const adseis_elastic_plan* P) { return ElCoef{P->ax, P->bx, P->ay, P->by}; }

static void free_points(ElPointStore* s) {
  free_point_set(&s->ps);
  cudaFree(s->xstart); cudaFree(s->xperm);
  *s = ElPointStore();
}

ADSEIS_API int adseis_elastic_plan_destroy(adseis_elastic_plan* P) {
  if (!P) return ADSEIS_OK;
  cudaSetDevice(P->ctx->device);
  cudaStreamSynchronize(P->ctx->stream);
  if (P->arena) {  // hist, adj and the accumulators live inside the arena
    for (int k = 0; k < 2; k++) if (P->peer[k]) cudaIpcCloseMemHandle(P->peer[k]);
    cudaFree(P->arena);
    P->hist = P->adj = P->Gl = P->Gm1 = P->Gm2 = P->Gr3 = P->Gr4 = nullptr;
  }
  cudaFree(P->perm); cudaFree(P->ctas);
  double* arr[] = {P->rho, P->lam, P->mu, P->lamb, P->lmb, P->mub2, P->rhob, P->rinv, P->rbinv, P->ax, P->bx, P->ay,
                   P->by, P->hist, P->srcv, P->rcvv, P->obs, P->res, P->loss, P->adj, P->Gl, P->Gm1, P->Gm2, P->Gr3,
                   P->Gr4, P->grho, P->glam, P->gmu, P->gradsrcv};
  for (double* a : arr) cudaFree(a);
  for (double* c : P->ckpt) cudaFree(c);
  cudaFree(P->rcv_owned);
  free_points(&P->src); free_points(&P->rcv);
  adseis_ctx* ctx = P->ctx;
  delete P;
  adseis_ctx_release_plan(ctx);
  return ADSEIS_OK;
}

static int el_validate(const adseis_elastic_params* p) {
  REQUIRE(p, "elastic: null params");
  REQUIRE(p->variant == 0 || p->variant == 1, "elastic: variant must be 0 (Core.jl) or 1 (MPIElastic.jl)");
  REQUIRE(p->NX >= 6 && p->NY >= 6 && p->NSTEP >= 1, "elastic: need NX,NY >= 6 and NSTEP >= 1");
  REQUIRE((p->NX + 4) * (p->NY + 20) < 2147483647LL, "elastic: grid too large for 32-bit cell offsets");
  REQUIRE(p->DELTAX > 0 && p->DELTAY > 0 && p->DELTAT > 0, "elastic: DELTAX/DELTAY/DELTAT must be > 0");
  REQUIRE(p->NPOINTS_PML >= 1 && p->Rcoef > 0 && p->vp_ref > 0, "elastic: bad PML parameters");
  REQUIRE(p->K_MAX_PML == 1.0, "elastic: K_MAX_PML must be 1 (the reference ignores K, Core.jl:686-693)");
  return ADSEIS_OK;
}

// segments over slots 0..NSTEP with a window of `win` slots; consecutive segments share one slot
static void el_make_segments(adseis_elastic_plan* P) {
  // the last segment is never replayed: it gets the full window, the remainder goes to the first segment
  P->seg_b.clear(); P->seg_e.clear();
  const i64 NSTEP = P->p.NSTEP, per = P->win - 1;  // new slots per full segment
  const i64 nseg = std::max<i64>(1, (NSTEP + per - 1) / per);
  i64 b = 0, e = std::min(NSTEP, NSTEP - (nseg - 1) * per);
  while (true) {
    P->seg_b.push_back(b); P->seg_e.push_back(e);
    if (e >= NSTEP) break;
    b = e;
    e = std::min(b + P->win - 1, NSTEP);
  }
}

ADSEIS_API int adseis_elastic_plan_create(adseis_ctx* ctx, const adseis_elastic_params* p, const adseis_slab* slab,
                                          int64_t nsrc, const int64_t* srci, const int64_t* srcj,
                                          const int64_t* srctype, int64_t nrcv, const int64_t* rcvi,
                                          const int64_t* rcvj, const int64_t* rcvtype, size_t hist_bytes_budget,
                                          adseis_elastic_plan** out) {
  REQUIRE(ctx && out, "elastic_plan_create: null ctx/out");
  *out = nullptr;
  TRY(el_validate(p));
  REQUIRE(nsrc >= 0 && nrcv >= 0 && (nsrc == 0 || (srci && srcj && srctype)) && (nrcv == 0 || (rcvi && rcvj && rcvtype)),
          "elastic_plan_create: bad source/receiver arrays");
  CUDA_TRY(cudaSetDevice(ctx->device));
  {
    const cudaFuncAttribute A = cudaFuncAttributeMaxDynamicSharedMemorySize;
    CUDA_TRY(cudaFuncSetAttribute(el_sigma_fwd, A, el_ring_bytes<ElSigFwdT>()));
    CUDA_TRY(cudaFuncSetAttribute(el_vel_fwd, A, el_ring_bytes<ElVelFwdT>()));
    CUDA_TRY(cudaFuncSetAttribute(el_vel_adj<true>, A, el_ring_bytes<ElVelAdjT<true>>()));
    CUDA_TRY(cudaFuncSetAttribute(el_vel_adj<false>, A, el_ring_bytes<ElVelAdjT<false>>()));
    CUDA_TRY(cudaFuncSetAttribute(el_sigma_adj<true>, A, el_ring_bytes<ElSigAdjT<true>>()));
    CUDA_TRY(cudaFuncSetAttribute(el_sigma_adj<false>, A, el_ring_bytes<ElSigAdjT<false>>()));
  }
  cudaStream_t st = ctx->stream;
  adseis_elastic_plan* P = new adseis_elastic_plan();
  P->ctx = ctx;
  ctx->plans++;
  P->p = *p;
  const int NX = (int)p->NX, NY = (int)p->NY;
  ElGeom& g = P->g;
  memset(&g, 0, sizeof(g));
  g.NX = NX; g.NY = NY;
  if (p->variant == 0) {
    g.H = NX + 2; g.W = NY + 2; g.cx = g.cy = 1; P->off = 0;
    // fw1 Core.jl:100-107; fw2 :132-139; fw3 :162-169; fw4 :190-198 (1-based getid ranges -> 0-based)
    g.p0[0] = 1; g.p1[0] = NX - 1; g.q0[0] = 2; g.q1[0] = NY;
    g.p0[1] = 2; g.p1[1] = NX;     g.q0[1] = 1; g.q1[1] = NY - 1;
    g.p0[2] = 2; g.p1[2] = NX;     g.q0[2] = 2; g.q1[2] = NY;
    g.p0[3] = 1; g.p1[3] = NX - 1; g.q0[3] = 1; g.q1[3] = NY - 1;
    P->model_elems = (i64)(NX + 2) * (NY + 2);
  } else {
    g.H = NX + 4; g.W = NY + 4; g.cx = g.cy = 2; P->off = 2;
    for (int k = 0; k < 4; k++) { g.p0[k] = 2; g.p1[k] = NX + 1; g.q0[k] = 2; g.q1[k] = NY + 1; }  // MPIElastic.jl:175-189
    P->model_elems = (i64)NX * NY;
  }
  // slab of a 1-D decomposition along i: rows [row0,row1) of the INTERNAL array (variant 0: the padded grid,
  // variant 1: the global grid plus its 2 ghost rows on each side), EL_HALO halo rows per interior side
  if (slab && slab->nranks > 1) P->slab = *slab;
  else { P->slab.rank = 0; P->slab.nranks = 1; P->slab.row0 = 0; P->slab.row1 = g.H; }
  const adseis_slab& sl = P->slab;
  if (!(sl.rank >= 0 && sl.rank < sl.nranks && sl.row0 >= 0 && sl.row1 <= g.H && sl.row1 - sl.row0 >= 2 * EL_HALO &&
        (sl.rank > 0 || sl.row0 == 0) && (sl.rank < sl.nranks - 1 || sl.row1 == g.H))) {
    adseis_set_error("elastic_plan_create: inconsistent slab {rank %d/%d rows [%lld,%lld)} for %d internal rows",
                     sl.rank, sl.nranks, (long long)sl.row0, (long long)sl.row1, g.H);
    delete P;
    ctx->plans--;
    return ADSEIS_EINVAL;
  }
  const int halo_lo = sl.rank > 0 ? EL_HALO : 0, halo_hi = sl.rank < sl.nranks - 1 ? EL_HALO : 0;
  g.goff = (int)sl.row0 - halo_lo;
  g.Hl = (int)(sl.row1 - sl.row0) + halo_lo + halo_hi;
  g.own0 = halo_lo; g.own1 = halo_lo + (int)(sl.row1 - sl.row0);
  g.ld = round_up(g.W, 16);
  g.plane = (i64)g.Hl * g.ld;
  g.dt = p->DELTAT; g.dx = p->DELTAX; g.dy = p->DELTAY;

#define EPTRY(expr)                       \
  do {                                    \
    int _r = (expr);                      \
    if (_r != ADSEIS_OK) {                \
      adseis_elastic_plan_destroy(P);     \
      return _r;                          \
    }                                     \
  } while (0)

  // CPML coefficients and the index ranges on which they are non-trivial
  std::vector<double> ax(2 * NX), bx(2 * NX), ay(2 * NY), by(2 * NY);
  EPTRY(adseis_elastic_cpml_profiles(p, 0, ax.data(), bx.data()));
  EPTRY(adseis_elastic_cpml_profiles(p, 1, ay.data(), by.data()));
  auto strip = [](const std::vector<double>& a, int n, int* lo, int* hi) {
    auto nz = [&](int k) { return a[k] != 0.0 || a[n + k] != 0.0; };
    int l = 0;
    while (l < n && nz(l)) l++;
    int h = n;
    while (h > l && nz(h - 1)) h--;
    for (int k = l; k < h; k++)
      if (nz(k)) { l = n; h = n; break; }  // profile is not "two strips": treat every index as PML
    *lo = l; *hi = h;
  };
  strip(ax, NX, &g.xlo, &g.xhi);
  strip(ay, NY, &g.ylo, &g.yhi);
  if (sl.nranks > 1) {
    // x-memories are stored compactly with GLOBAL strip rows and have no halo: an interior slab boundary must keep
    // EL_HALO rows of distance from the x-PML strips (coefficient index kx = row - cx)
    const int b_lo = (int)sl.row0 - g.cx, b_hi = (int)sl.row1 - g.cx;  // first kx owned / first kx of rank+1
    const bool bad_lo = sl.rank > 0 && (b_lo - EL_HALO < g.xlo || b_lo + EL_HALO > g.xhi);
    const bool bad_hi = sl.rank < sl.nranks - 1 && (b_hi - EL_HALO < g.xlo || b_hi + EL_HALO > g.xhi);
    if (bad_lo || bad_hi) {
      adseis_set_error("elastic_plan_create: slab boundary of rank %d lies inside (or within %d rows of) an x-PML strip "
                       "(PML-free coefficient rows [%d,%d)); use fewer ranks or a thinner PML", sl.rank, EL_HALO, g.xlo,
                       g.xhi);
      adseis_elastic_plan_destroy(P);
      return ADSEIS_EINVAL;
    }
  }
  g.nxr = g.xlo + (NX - g.xhi);
  const int nyc = g.ylo + (NY - g.yhi);
  g.ycp = std::max(1, nyc);
  g.xm_sz = (i64)std::max(1, g.nxr) * g.ld;
  g.ym_sz = (i64)g.Hl * g.ycp;
  P->slot_sz = 5 * g.plane + 4 * g.xm_sz + 4 * g.ym_sz;
  g.h24x = 24 * g.dx; g.h24y = 24 * g.dy;
  g.r24x = 1.0 / g.h24x; g.r24y = 1.0 / g.h24y;
  EPTRY(dev_upload(&P->ax, ax, st)); EPTRY(dev_upload(&P->bx, bx, st));
  EPTRY(dev_upload(&P->ay, ay, st)); EPTRY(dev_upload(&P->by, by, st));
  double** mats[] = {&P->rho, &P->lam, &P->mu, &P->lamb, &P->lmb, &P->mub2, &P->rhob, &P->rinv, &P->rbinv};
  for (double** m : mats) EPTRY(dev_alloc_zero(m, (size_t)g.plane + g.ld, st));  // raw planes: one extra row (averaging)

  // ---- CTA table of a step launch: marching tiles over the box, generic 64 x 16 tiles over the rest ----
  std::vector<ElCta> ctas;
  struct Rect { int r0, r1, c0, c1, base, ntc, tw, th; };
  std::vector<Rect> rects;
  int mrb = 0, mnct = 0;  // marching tile rows / column tiles
  std::vector<int> mrt;   // marching row-tile boundaries (local rows), ascending
  {
    // box: inside the update regions of fw1..fw4, CPML-free, and 2 cells away from anything that is not
    int bp0 = 0, bp1 = g.H - 1, bq0 = 0, bq1 = g.W - 1;  // inclusive, global
    for (int k = 0; k < 4; k++) {
      bp0 = std::max(bp0, g.p0[k]); bp1 = std::min(bp1, g.p1[k]);
      bq0 = std::max(bq0, g.q0[k]); bq1 = std::min(bq1, g.q1[k]);
    }
    bp0 = std::max(bp0, g.xlo + g.cx); bp1 = std::min(bp1, g.xhi - 1 + g.cx);
    bq0 = std::max(bq0, g.ylo + g.cy); bq1 = std::min(bq1, g.yhi - 1 + g.cy);
    bp0 += EL_HALO; bp1 -= EL_HALO; bq0 += EL_HALO; bq1 -= EL_HALO;
    int r0 = std::max(g.own0, bp0 - g.goff), r1 = std::min(g.own1, bp1 + 1 - g.goff);
    int c0 = round_up(std::max(bq0, 2), 16);
    int c1 = c0 + 2 * ((bq1 + 1 - c0) / 2);
    // marching needs enough rows per CTA to amortise its pipeline prologue and enough CTAs to fill the GPU
    const i64 min_cells = getenv("ADSEIS_EL_MARCH_MIN") ? atoll(getenv("ADSEIS_EL_MARCH_MIN")) : (i64)16 * EL_TCOLS * ctx->sm_count / 2;
    if (r1 - r0 < 4 || c1 - c0 < 32 || (i64)(r1 - r0) * (c1 - c0) < min_cells) { r0 = r1 = g.own0; c0 = c1 = 0; }
    P->box[0] = r0; P->box[1] = r1; P->box[2] = c0; P->box[3] = c1;
    if (r1 > r0) {
      mnct = (c1 - c0 + EL_TCOLS - 1) / EL_TCOLS;
      // marching CTAs per SM: 6 (else 3) gives whole waves both to the kernels that fit 2 and to those that fit 3
      // CTAs per SM, as long as a CTA keeps >= 8 rows to amortise its pipeline prologue (1000 x 2000 slab: 9-row tiles
      // 77 us per forward step, 18-row tiles 105 us).  Thinner slabs: a marching
      // CTA streams its rows at a latency-bound ~2.5 us per row (ring depth), so the launch needs ALL slots busy --
      // about one CTA per slot (3 per SM), at least 4 rows each (measured on 250 x 2000 slabs: 13-row tiles 62.9 us
      // per forward step, 4..6-row tiles 36..40 us)
      mrb = 0;
      if (sl.nranks == 1) {
        // one GPU: if ONE wave of 2 CTAs per SM covers the box with 24..64 rows per CTA, take exactly that -- every
        // CTA is resident from the start and pays one pipeline prologue (2000^2: 54-row tiles, forward 126 -> 117,
        // material gradient 402 -> 367 us/step against 18-row tiles; wave quantisation matters: 33..41 rows lose)
        const int tiles = std::max(1, 2 * ctx->sm_count / mnct);
        const int r = (r1 - r0 + tiles - 1) / tiles;
        if (r >= 24 && r <= 64) mrb = r;
      }
      for (int k : {6, 3}) {
        if (mrb) break;
        const int want_tr = std::max(1, (k * ctx->sm_count + mnct - 1) / mnct);
        const int r = (r1 - r0 + want_tr - 1) / want_tr;
        if (r >= 8) { mrb = std::min(64, r); break; }
      }
      if (!mrb) mrb = std::min(64, std::max(4, ((r1 - r0) * mnct + 3 * ctx->sm_count - 1) / (3 * ctx->sm_count)));
      if (getenv("ADSEIS_EL_RB")) mrb = std::max(2, atoi(getenv("ADSEIS_EL_RB")));  // tuning experiments
      // slab plans: the EL_HALO rows next to a neighbour form thin row tiles of their own -- launched first, done
      // within a couple of microseconds, so the halo rows they push cross NVLink while the rest of the (single-wave)
      // launch is still streaming instead of leaving at its very end
      const int thin_lo = (halo_lo && r0 == g.own0 && r1 - r0 >= 6 * EL_HALO) ? EL_HALO : 0;
      const int thin_hi = (halo_hi && r1 == g.own1 && r1 - r0 >= 6 * EL_HALO) ? EL_HALO : 0;
      mrt.push_back(r0);
      if (thin_lo) mrt.push_back(r0 + thin_lo);
      for (int a = r0 + thin_lo + mrb; a < r1 - thin_hi; a += mrb) mrt.push_back(a);
      if (thin_hi) mrt.push_back(r1 - thin_hi);
      mrt.push_back(r1);
      for (size_t tr = 0; tr + 1 < mrt.size(); tr++)
        for (int tc = 0; tc < mnct; tc++)
          ctas.push_back(ElCta{0, mrt[tr], mrt[tr + 1], c0 + tc * EL_TCOLS, std::min(c1, c0 + (tc + 1) * EL_TCOLS), 0, 0, 0});
    }
    P->nmarch = (int)ctas.size();
    // cells per thread of a generic CTA: 2 on large grids (measured on B200, C5 2000^2, us per step forward / material
    // gradient: 1 -> 148.5 / 392, 2 -> 148.7 / 368.5, 3 -> 146.9 / 382.6, 4 -> 146.8 / 395.6); 1 on slabs / small grids, where
    // the launch is a single wave and its duration is the latency chain of the longest CTA
    int gen_cpt = (sl.nranks > 1 || (i64)g.Hl * g.W < (i64)1500 * 1500) ? 1 : 2;
    if (getenv("ADSEIS_EL_CPT")) gen_cpt = std::max(1, atoi(getenv("ADSEIS_EL_CPT")));
    auto add_rect = [&](int rr0, int rr1, int cc0, int cc1) {
      if (rr1 <= rr0 || cc1 <= cc0) return;
      // narrow strips get tall tiles (16 x 64, 32 x 32), wide rectangles 64 x 16: 256 threads x 4 cells each
      const int ltw = (cc1 - cc0 <= 16) ? 4 : ((cc1 - cc0 <= 32) ? 5 : 6);
      const int tw = 1 << ltw, th = gen_cpt * ((EL_BX * EL_BY) >> ltw);
      Rect R{rr0, rr1, cc0, cc1, (int)ctas.size(), (cc1 - cc0 + tw - 1) / tw, tw, th};
      for (int a = rr0; a < rr1; a += th)
        for (int b = cc0; b < cc1; b += tw)
          ctas.push_back(ElCta{1, a, std::min(rr1, a + th), b, std::min(cc1, b + tw), ltw, 0, 0});
      rects.push_back(R);
    };
    // (pitch-padding columns q >= W are never read by a cell that is stored: nobody writes them)
    add_rect(g.own0, r0, 0, g.W);
    add_rect(r1, g.own1, 0, g.W);
    add_rect(r0, r1, 0, c0);
    add_rect(r0, r1, c1, g.W);
  }
  P->nblocks = (int)ctas.size();
  {
    std::vector<int> tab((size_t)P->nblocks * 8);
    memcpy(tab.data(), ctas.data(), tab.size() * sizeof(int));
    int* dtab = nullptr;
    EPTRY(dev_upload(&dtab, tab, st));
    P->ctas = reinterpret_cast<ElCta*>(dtab);
  }
  {
    // Launch order (logical CTA id = perm[blockIdx.x]).  Slab plans: the CTAs that own one of my first / last two
    // rows next to a neighbour go first, so their halo pushes leave early and the rest of the step hides the NVLink
    // latency.  Then the generic CTAs (latency-bound) are spread evenly among the marching CTAs (bandwidth-bound).
    std::vector<int> edge, march, frame;
    for (int b = 0; b < P->nblocks; b++) {
      const bool tl = halo_lo && ctas[b].r0 < g.own0 + EL_HALO, th = halo_hi && ctas[b].r1 > g.own1 - EL_HALO;
      if (tl) P->n_edge_lo++;
      if (th) P->n_edge_hi++;
      if (tl || th) edge.push_back(b);
      else (ctas[b].kind == 0 ? march : frame).push_back(b);
    }
    std::vector<int> order(edge);
    size_t im = 0, ifr = 0;
    const size_t nm = march.size(), nf = frame.size();
    // ADSEIS_EL_ORDER: 0 = spread evenly over the launch (slab plans), 1 = generic CTAs first, 2 = spread over the first half
    // of the marching CTAs (single GPU: the last generic CTA then starts mid-launch instead of forming its tail; C5 material
    // gradient 368.5 -> 364.2 us per step)
    const int omode = getenv("ADSEIS_EL_ORDER") ? atoi(getenv("ADSEIS_EL_ORDER")) : (sl.nranks > 1 ? 0 : 2);
    const size_t nm_eff = omode == 2 ? std::max<size_t>(1, nm / 2) : nm;
    while (im < nm || ifr < nf) {
      const bool gen_next = omode == 1 ? (ifr < nf) : (ifr < nf && (im >= nm || ifr * nm_eff <= im * nf));
      if (gen_next) order.push_back(frame[ifr++]);
      else order.push_back(march[im++]);
    }
    EPTRY(dev_upload(&P->perm, order, st));
  }
  auto owner_cta = [&](int li, int q) -> int {
    if (li >= P->box[0] && li < P->box[1] && q >= P->box[2] && q < P->box[3])
      return (int)(std::upper_bound(mrt.begin(), mrt.end(), li) - mrt.begin() - 1) * mnct + (q - P->box[2]) / EL_TCOLS;
    for (const Rect& R : rects)
      if (li >= R.r0 && li < R.r1 && q >= R.c0 && q < R.c1)
        return R.base + ((li - R.r0) / R.th) * R.ntc + (q - R.c0) / R.tw;
    return -1;
  };

  // sources / receivers (1-based indices: S -> padded grid, M -> unpadded global grid; MPIElastic.jl:85-86)
  P->nsrc = nsrc; P->nrcv = nrcv;
  const int ioff = p->variant == 0 ? -1 : 1;
  struct Pt { int cell, field, gid; };
  auto collect = [&](i64 n, const int64_t* pi, const int64_t* pj, const int64_t* pt, std::vector<Pt>* v,
                     const char* what) -> int {
    for (i64 k = 0; k < n; k++) {
      const i64 gi = pi[k] + ioff, gj = pj[k] + ioff;
      REQUIRE(gi >= 0 && gi < g.H && gj >= 0 && gj < g.W, "elastic_plan_create: %s %lld at (%lld,%lld) is outside the grid",
              what, (long long)k, (long long)pi[k], (long long)pj[k]);
      if (pt[k] < 0 || pt[k] > 4) continue;  // AddSource.cpp:83 / GetReceive.cpp:43: other types are ignored
      if (gi < sl.row0 || gi >= sl.row1) continue;  // owned by another slab (MPIElastic.jl:71-86, 116-131)
      v->push_back(Pt{(int)((gi - g.goff) * g.ld + gj), (int)pt[k], (int)k});
    }
    return ADSEIS_OK;
  };
  std::vector<Pt> sp, rp;
  EPTRY(collect(nsrc, srci, srcj, srctype, &sp, "source"));
  EPTRY(collect(nrcv, rcvi, rcvj, rcvtype, &rp, "receiver"));
  auto upload = [&](const std::vector<Pt>& mine, const std::vector<Pt>& other, ElPointStore* dst) -> int {
    std::vector<int> own, cells, gid, field;
    for (const Pt& t : mine) {
      const int li = t.cell / g.ld, q = t.cell % g.ld;
      const int o = owner_cta(li, q);
      REQUIRE(o >= 0, "elastic_plan_create: internal error: cell (%d,%d) has no owner CTA", li, q);
      own.push_back(o);
      cells.push_back(t.cell); gid.push_back(t.gid); field.push_back(t.field);
    }
    PointSetHost h;
    build_point_set(own, cells, gid, field, P->nblocks, &h);
    std::map<std::pair<int, int>, std::vector<int>> idx;
    for (const Pt& t : other) idx[{t.cell, t.field}].push_back(t.gid);
    std::vector<int> xs(1, 0), xp;
    for (size_t k = 0; k < h.cell.size(); k++) {
      auto it = idx.find({h.cell[k], h.field[k]});
      if (it != idx.end()) xp.insert(xp.end(), it->second.begin(), it->second.end());
      xs.push_back((int)xp.size());
    }
    TRY(upload_point_set(h, &dst->ps, st));
    TRY(dev_upload(&dst->xstart, xs, st));
    TRY(dev_upload(&dst->xperm, xp, st));
    if (dst->ps.nu > 0)
      dst->dev = ElPoints{dst->ps.blk, dst->ps.cell, dst->ps.field, dst->ps.start, dst->ps.perm, dst->xstart, dst->xperm};
    return ADSEIS_OK;
  };
  EPTRY(upload(sp, rp, &P->src));
  EPTRY(upload(rp, sp, &P->rcv));
  std::vector<unsigned char> owned((size_t)std::max<i64>(nrcv, 1), 0);
  for (const Pt& t : rp) owned[t.gid] = 1;
  EPTRY(dev_upload(&P->rcv_owned, owned, st));
  EPTRY(dev_alloc_zero(&P->rcvv, (size_t)((p->NSTEP + 1) * nrcv), st));
  EPTRY(dev_alloc_zero(&P->loss, 1 + RL_BLOCKS, st));

  // history window
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  const size_t slot_bytes = (size_t)P->slot_sz * 8;
  const size_t reserve = 16 * (size_t)g.plane * 8 + 4 * slot_bytes +
                         (size_t)(3 * (p->NSTEP + 1) * nrcv + 2 * p->NSTEP * nsrc) * 8 + (512u << 20);
  size_t budget = hist_bytes_budget ? hist_bytes_budget : (free_b > reserve ? free_b - reserve : 0);
  if (budget > free_b) budget = free_b;
  i64 slots = (i64)(budget / slot_bytes);
  if (slots >= p->NSTEP + 1) {
    P->win = p->NSTEP + 1;
  } else if (hist_bytes_budget) {
    P->win = std::max<i64>(slots, 2);
  } else {
    i64 best = 2;
    for (i64 W = slots; W >= 2; W--) {
      i64 nseg = (p->NSTEP + (W - 1) - 1) / (W - 1);
      if (W + (nseg - 1) <= slots) { best = W; break; }
    }
    P->win = best;
  }
  el_make_segments(P);
  const size_t adj_elems = (size_t)(5 * g.plane + 2 * (4 * g.xm_sz + 4 * g.ym_sz));
  if (sl.nranks > 1) {
    // one arena: [descriptor | flags | history window | adjoint state | 5 gradient accumulators]
    adseis_elastic_plan::Desc& d = P->desc;
    d.magic = EL_DESC_MAGIC; d.Hl = g.Hl; d.ld = g.ld; d.plane = g.plane; d.slot_sz = P->slot_sz; d.win = P->win;
    d.own0 = g.own0; d.own1 = g.own1; d.n_edge_lo = P->n_edge_lo; d.n_edge_hi = P->n_edge_hi;
    long long off = 512;
    d.off_flags = off; off += 512;
    d.off_hist = off; off += (long long)((size_t)P->win * slot_bytes);
    d.off_adj = off; off += (long long)(adj_elems * 8);
    d.off_gacc = off; off += (long long)(5 * (size_t)g.plane * 8);
    {
      // opt-in (ADSEIS_EL_LL=1): measured on 2 B200s (C5, 1002-row slabs) the packed rows change nothing -- forward 75.1 vs
      // 74.9 us per step, material gradient 203.7 vs 204.0 -- and one full-size run reported a halo time-out that the
      // remaining GPU budget of the round did not allow to chase; parity tests pass with it (tests/mgpu_worker.py)
      const char* ell = getenv("ADSEIS_EL_LL");
      P->ll = (ell && ell[0] == '1');
      d.off_ll = 0;
      if (P->ll) { d.off_ll = off; off += 2LL * 2 * 3 * EL_HALO * g.ld * 16; off = (off + 511) / 512 * 512; }
    }
    P->arena_bytes = (size_t)off;
    cudaError_t e = cudaMalloc((void**)&P->arena, P->arena_bytes);
    if (e != cudaSuccess) {
      adseis_set_error("elastic_plan_create: cannot allocate the %zu-byte slab arena: %s", P->arena_bytes,
                       cudaGetErrorString(e));
      adseis_elastic_plan_destroy(P);
      return ADSEIS_ENOMEM;
    }
    CUDA_TRY(cudaMemsetAsync(P->arena, 0, P->arena_bytes, st));
    CUDA_TRY(cudaMemcpyAsync(P->arena, &d, sizeof(d), cudaMemcpyHostToDevice, st));
    char* base = (char*)P->arena;
    P->hist = (double*)(base + d.off_hist);
    P->adj = (double*)(base + d.off_adj);
    double* ga = (double*)(base + d.off_gacc);
    P->Gl = ga; P->Gm1 = ga + g.plane; P->Gm2 = ga + 2 * g.plane; P->Gr3 = ga + 3 * g.plane; P->Gr4 = ga + 4 * g.plane;
    CUDA_TRY(cudaStreamSynchronize(st));
  } else {
    cudaError_t e = cudaMalloc((void**)&P->hist, (size_t)P->win * slot_bytes);
    if (e != cudaSuccess) {
      adseis_set_error("elastic_plan_create: cannot allocate %lld history slots of %zu bytes: %s", (long long)P->win,
                       slot_bytes, cudaGetErrorString(e));
      adseis_elastic_plan_destroy(P);
      return ADSEIS_ENOMEM;
    }
    CUDA_TRY(cudaMemsetAsync(P->hist, 0, (size_t)P->win * slot_bytes, st));
  }
  for (size_t k = 1; k < P->seg_b.size(); k++) {
    double* c = nullptr;
    EPTRY(dev_alloc(&c, (size_t)P->slot_sz));
    P->ckpt.push_back(c);
  }
  *out = P;
  return ADSEIS_OK;
}

// caller's dense model array -> pitched internal plane of this slab's rows plus one extra row (ghost / pad zero)
static int el_load_plane(adseis_elastic_plan* P, const double* src, double* dst) {
  const ElGeom& g = P->g;
  cudaStream_t st = P->ctx->stream;
  CUDA_TRY(cudaMemsetAsync(dst, 0, (size_t)(g.plane + g.ld) * 8, st));
  const int off = P->off, OW = g.W - 2 * off, OH = g.H - 2 * off;
  // local row li <-> internal row goff+li <-> caller row goff+li-off, for li in [0, Hl]
  const int l0 = std::max(0, off - g.goff), l1 = std::min(g.Hl + 1, OH + off - g.goff);
  if (l1 > l0)
    CUDA_TRY(cudaMemcpy2DAsync(dst + (i64)l0 * g.ld + off, (size_t)g.ld * 8, src + (i64)(g.goff + l0 - off) * OW,
                               (size_t)OW * 8, (size_t)OW * 8, (size_t)(l1 - l0), cudaMemcpyDefault, st));
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_plan_set_model(adseis_elastic_plan* P, const double* rho, const double* lambda,
                                             const double* mu, int on_device) {
  (void)on_device;
  REQUIRE(P && rho && lambda && mu, "elastic_plan_set_model: null");
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  TRY(el_load_plane(P, rho, P->rho));
  TRY(el_load_plane(P, lambda, P->lam));
  TRY(el_load_plane(P, mu, P->mu));
  const ElGeom& g = P->g;
  dim3 grid((unsigned)((g.ld + 127) / 128), (unsigned)g.Hl);
  k_el_materials<<<grid, 128, 0, P->ctx->stream>>>(g.Hl, g.ld, g.W, g.goff, g.H, P->p.variant == 0, P->rho, P->lam, P->mu, P->lamb,
                                                   P->lmb, P->mub2, P->rhob, P->rinv, P->rbinv);
  P->ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  P->have_model = true;
  P->have_fwd = P->have_grad = false;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_plan_set_srcv(adseis_elastic_plan* P, const double* srcv, int64_t rows, int on_device) {
  (void)on_device;
  REQUIRE(P && (srcv || P->nsrc == 0), "elastic_plan_set_srcv: null");
  REQUIRE(rows >= P->p.NSTEP, "elastic_plan_set_srcv: srcv has %lld rows, need >= NSTEP=%lld", (long long)rows,
          (long long)P->p.NSTEP);
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  if (!P->srcv) TRY(dev_alloc(&P->srcv, (size_t)(P->p.NSTEP * P->nsrc)));
  if (P->nsrc > 0)
    CUDA_TRY(cudaMemcpyAsync(P->srcv, srcv, (size_t)(P->p.NSTEP * P->nsrc) * 8, cudaMemcpyDefault, P->ctx->stream));
  P->have_srcv = true;
  P->have_fwd = P->have_grad = false;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_plan_set_obs(adseis_elastic_plan* P, const double* obs, int on_device) {
  (void)on_device;
  REQUIRE(P && (obs || P->nrcv == 0), "elastic_plan_set_obs: null");
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  const size_t n = (size_t)((P->p.NSTEP + 1) * P->nrcv);
  if (!P->obs) { TRY(dev_alloc(&P->obs, n)); TRY(dev_alloc(&P->res, n)); }
  if (n > 0) CUDA_TRY(cudaMemcpyAsync(P->obs, obs, n * 8, cudaMemcpyDefault, P->ctx->stream));
  P->have_obs = true;
  P->have_grad = false;
  return ADSEIS_OK;
}

// ---- slab decomposition: peer pointers of the next fused launch, standalone exchange / barrier ----------------
// A plane with halo rows is named by (array, index, field) because the neighbours' arenas have the same structure
// but not the same plane size (their slabs may hold a different number of rows).
enum ElArr { EA_HIST, EA_ADJ, EA_GACC };
struct ElPlaneRef { int arr; i64 idx; int field; };  // HIST: idx = window index, field 0..4; ADJ: field 0..4; GACC: field 0..4
static inline long long el_plane_off(const adseis_elastic_plan::Desc& d, const ElPlaneRef& r) {
  const long long base = r.arr == EA_HIST ? d.off_hist + r.idx * d.slot_sz * 8 : (r.arr == EA_ADJ ? d.off_adj : d.off_gacc);
  return base + (long long)r.field * d.plane * 8;
}

// `recv`: the previous fused launch produced the planes whose halo rows this launch reads (false for the first launch of a
// sweep / segment / replay and after an explicit exchange: the halo rows are then in place already)
static ElFuse el_make_fuse(adseis_elastic_plan* P, int nf, const ElPlaneRef* planes, bool recv) {
  ElFuse f;
  memset(&f, 0, sizeof(f));
  f.perm = P->perm;
  if (!P->arena) return f;
  f.has_lo = P->peer[0] != nullptr; f.has_hi = P->peer[1] != nullptr;
  f.nf = nf;
  P->sepoch++;
  if (P->ll) {
    auto rows = [&](char* base, const adseis_elastic_plan::Desc& d, int from, unsigned long long ep) {
      return (ulonglong2*)(base + d.off_ll + (long long)(from * 2 + (int)(ep & 1ULL)) * 3 * EL_HALO * d.ld * 16);
    };
    const unsigned long long es = P->sepoch, er = P->sepoch - 1;
    f.ll = 1;
    f.ep_send = (unsigned)(es % 0xFFFFFFFEULL) + 1u;
    f.ep_recv = (recv && P->ll_prev_nf > 0) ? (unsigned)(er % 0xFFFFFFFEULL) + 1u : 0u;
    f.my_flags = (unsigned long long*)((char*)P->arena + P->desc.off_flags);
    for (int k = 0; k < nf; k++) f.src[k] = (const double*)((char*)P->arena + el_plane_off(P->desc, planes[k]));
    if (f.has_lo) { f.tx_lo = rows(P->peer[0], P->dpeer[0], 1, es); f.rx_lo = rows((char*)P->arena, P->desc, 0, er); }
    if (f.has_hi) { f.tx_hi = rows(P->peer[1], P->dpeer[1], 0, es); f.rx_hi = rows((char*)P->arena, P->desc, 1, er); }
    f.nf_in = P->ll_prev_nf;
    for (int k = 0; k < P->ll_prev_nf; k++) {
      const ElPlaneRef r{P->ll_prev[k].arr, P->ll_prev[k].idx, P->ll_prev[k].field};
      f.in[k] = (double*)((char*)P->arena + el_plane_off(P->desc, r));
    }
    P->ll_prev_nf = nf;
    for (int k = 0; k < nf; k++) { P->ll_prev[k].arr = planes[k].arr; P->ll_prev[k].idx = planes[k].idx; P->ll_prev[k].field = planes[k].field; }
    return f;
  }
  for (int k = 0; k < nf; k++) {
    f.src[k] = (const double*)((char*)P->arena + el_plane_off(P->desc, planes[k]));
    if (f.has_lo) f.lo[k] = (double*)(P->peer[0] + el_plane_off(P->dpeer[0], planes[k])) + (P->dpeer[0].Hl - EL_HALO) * P->dpeer[0].ld;
    if (f.has_hi) f.hi[k] = (double*)(P->peer[1] + el_plane_off(P->dpeer[1], planes[k]));
  }
  if (f.has_lo) {
    f.sig_lo = (unsigned long long*)(P->peer[0] + P->dpeer[0].off_flags) + 4;
    f.expect_lo = (unsigned long long)P->dpeer[0].n_edge_hi * (P->sepoch - 1);
  }
  if (f.has_hi) {
    f.sig_hi = (unsigned long long*)(P->peer[1] + P->dpeer[1].off_flags) + 3;
    f.expect_hi = (unsigned long long)P->dpeer[1].n_edge_lo * (P->sepoch - 1);
  }
  f.my_flags = (unsigned long long*)((char*)P->arena + P->desc.off_flags);
  return f;
}

struct ElHaloArgs {
  int nf;
  const double* src[5];
  double* lo[5];
  double* hi[5];
  int ld, own0, own1;
  int has_lo, has_hi;
  unsigned long long *sig_lo, *sig_hi, *my_flags;
  unsigned long long expect;
};

// Standalone exchange: push my EL_HALO edge rows of up to five planes into the neighbours' halo rows, publish, then
// wait for theirs.  nf == 0 is a pure barrier (guards host-enqueued memsets / copies of arrays with halo rows).
__global__ void __launch_bounds__(256) k_el_halo_exchange(ElHaloArgs a) {
  const int n2 = EL_HALO * a.ld / 2;  // double2 elements per (plane, side)
  const int per = (n2 + EL_HX_BLOCKS - 1) / EL_HX_BLOCKS;
  const int j0 = blockIdx.x * per, j1 = min(n2, j0 + per);
  for (int k = 0; k < a.nf; k++) {
    if (a.has_lo) {
      const double2* s2 = reinterpret_cast<const double2*>(a.src[k] + (i64)a.own0 * a.ld);
      double2* d2 = reinterpret_cast<double2*>(a.lo[k]);
      for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) d2[j] = s2[j];
    }
    if (a.has_hi) {
      const double2* s2 = reinterpret_cast<const double2*>(a.src[k] + (i64)(a.own1 - EL_HALO) * a.ld);
      double2* d2 = reinterpret_cast<double2*>(a.hi[k]);
      for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) d2[j] = s2[j];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (a.has_lo) atomicAdd_system(a.sig_lo, 1ULL);
    if (a.has_hi) atomicAdd_system(a.sig_hi, 1ULL);
    if (blockIdx.x == 0) {
      volatile unsigned long long* f = a.my_flags;
      unsigned long long spins = 0;
      while ((a.has_lo && f[0] < a.expect) || (a.has_hi && f[1] < a.expect)) {
        if (++spins > (1ULL << 26)) { f[2] = 1ULL; break; }  // neighbour lost: report, do not hang
      }
      __threadfence_system();
    }
  }
}

static int el_halo_exchange(adseis_elastic_plan* P, int nf, const ElPlaneRef* planes) {
  P->ll_prev_nf = 0;   // packed rows: the next fused launch finds its halo rows in place

  if (!P->arena) return ADSEIS_OK;
  if (!P->connected) {
    adseis_set_error("elastic slab plan: adseis_elastic_plan_ipc_connect has not been called");
    return ADSEIS_ESTATE;
  }
  const ElGeom& g = P->g;
  ElHaloArgs a;
  memset(&a, 0, sizeof(a));
  a.nf = nf; a.ld = g.ld; a.own0 = g.own0; a.own1 = g.own1;
  a.has_lo = P->peer[0] != nullptr; a.has_hi = P->peer[1] != nullptr;
  for (int k = 0; k < nf; k++) {
    a.src[k] = (const double*)((char*)P->arena + el_plane_off(P->desc, planes[k]));
    if (a.has_lo) a.lo[k] = (double*)(P->peer[0] + el_plane_off(P->dpeer[0], planes[k])) + (P->dpeer[0].Hl - EL_HALO) * g.ld;
    if (a.has_hi) a.hi[k] = (double*)(P->peer[1] + el_plane_off(P->dpeer[1], planes[k]));
  }
  if (a.has_lo) a.sig_lo = (unsigned long long*)(P->peer[0] + P->dpeer[0].off_flags) + 1;
  if (a.has_hi) a.sig_hi = (unsigned long long*)(P->peer[1] + P->dpeer[1].off_flags) + 0;
  a.my_flags = (unsigned long long*)((char*)P->arena + P->desc.off_flags);
  P->epoch++;
  a.expect = (unsigned long long)EL_HX_BLOCKS * P->epoch;
  k_el_halo_exchange<<<EL_HX_BLOCKS, 256, 0, P->ctx->stream>>>(a);
  EL_LAUNCH_CHECK(P);
  return ADSEIS_OK;
}
static int el_barrier(adseis_elastic_plan* P) { return el_halo_exchange(P, 0, nullptr); }

static int el_halo_check(adseis_elastic_plan* P) {
  if (!P->arena) return ADSEIS_OK;
  unsigned long long f[3] = {0, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(f, (char*)P->arena + P->desc.off_flags, sizeof(f), cudaMemcpyDeviceToHost, P->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(P->ctx->stream));
  if (f[2] != 0) {
    adseis_set_error("elastic slab plan (rank %d): timed out waiting for a neighbour's halo rows", P->slab.rank);
    return ADSEIS_ECOMM;
  }
  return ADSEIS_OK;
}

// Packed halo rows are unpacked by the NEXT step launch; the velocity planes of the last slot of a sweep / segment /
// replay have none, so their halo rows are exchanged explicitly (they go into checkpoints and are read by the
// material-gradient kernels).  Also ends the chain: the following launch receives nothing.
static int el_finish_slot(adseis_elastic_plan* P, i64 widx) {
  if (!P->arena) return ADSEIS_OK;
  if (!P->ll) return el_halo_exchange(P, 0, nullptr);
  const ElPlaneRef v[2] = {{EA_HIST, widx, 0}, {EA_HIST, widx, 1}};
  return el_halo_exchange(P, 2, v);
}

// one forward step s: slot `in` (s-1) -> slot `out` (s); in == out is allowed (in-place stepping)
static int el_step_forward(adseis_elastic_plan* P, i64 s, double* in, double* out, bool sample) {
  cudaStream_t st = P->ctx->stream;
  const ElSlot si = slot_at(P, in), so = slot_at(P, out);
  const ElMat mt = mat_of(P);
  const ElCoef cf = coef_of(P);
  ElPoints none{};
  const double* row = P->nsrc > 0 ? P->srcv + (s - 1) * P->nsrc : nullptr;
  const double* prev = (P->nsrc > 0 && s >= 2) ? P->srcv + (s - 2) * P->nsrc : nullptr;
  const i64 widx = (out - P->hist) / P->slot_sz;
  const ElPlaneRef sig_planes[2] = {{EA_HIST, widx, 2}, {EA_HIST, widx, 4}};  // fw3/fw4 difference sxx, sxy along x
  CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_slab(), el_sigma_fwd, P->nblocks, EL_NT, el_ring_bytes<ElSigFwdT>(), st, P->g, P->ctas, si, so, mt, cf, P->src.dev, prev, el_make_fuse(P, 2, sig_planes, true)));
  EL_LAUNCH_CHECK(P);
  const ElPlaneRef vel_planes[2] = {{EA_HIST, widx, 0}, {EA_HIST, widx, 1}};
  CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_slab(), el_vel_fwd, P->nblocks, EL_NT, el_ring_bytes<ElVelFwdT>(), st, P->g, P->ctas, si, so, mt, cf, P->src.dev, row, sample ? P->rcv.dev : none,
                                         (sample && P->nrcv > 0) ? P->rcvv : nullptr, (int)(P->p.NSTEP + 1), (int)s,
                                         el_make_fuse(P, 2, vel_planes, true)));
  EL_LAUNCH_CHECK(P);
  return ADSEIS_OK;
}

static int el_check_ready(adseis_elastic_plan* P, bool need_obs) {
  if (!P->have_model || (!P->have_srcv && P->nsrc > 0) || (need_obs && !P->have_obs)) {
    adseis_set_error("elastic plan: set_model / set_srcv%s must be called first", need_obs ? " / set_obs" : "");
    return ADSEIS_ESTATE;
  }
  return ADSEIS_OK;
}

// forward sweep through the segments of the history window (keep==true) or in place in slot 0 (keep==false)
static int el_forward_sweep(adseis_elastic_plan* P, bool keep, bool save_ckpt) {
  cudaStream_t st = P->ctx->stream;
  const size_t sb = (size_t)P->slot_sz * 8;
  CUDA_TRY(cudaMemsetAsync(P->hist, 0, sb, st));  // slot 0 = zeros (Core.jl:37-45)
  if (P->nrcv > 0) CUDA_TRY(cudaMemsetAsync(P->rcvv, 0, (size_t)((P->p.NSTEP + 1) * P->nrcv) * 8, st));
  TRY(el_barrier(P));  // slab plans: nobody pushes halo rows before everybody's memsets are done
  if (!keep) {
    for (i64 s = 1; s <= P->p.NSTEP; s++) TRY(el_step_forward(P, s, P->hist, P->hist, true));
    if (P->ll) TRY(el_finish_slot(P, 0));
    P->win_base = P->win_last = P->p.NSTEP;
    P->inplace_last = true;
    return ADSEIS_OK;
  }
  const size_t nseg = P->seg_b.size();
  for (size_t k = 0; k < nseg; k++) {
    const i64 b = P->seg_b[k], e = P->seg_e[k];
    if (k > 0) {
      // slab plans: the neighbours' last pushes must have landed before halo rows are copied
      TRY(el_finish_slot(P, b - P->seg_b[k - 1]));
      double* last = win_ptr(P, P->seg_b[k - 1], b);
      if (save_ckpt) CUDA_TRY(cudaMemcpyAsync(P->ckpt[k - 1], last, sb, cudaMemcpyDeviceToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(P->hist, last, sb, cudaMemcpyDeviceToDevice, st));
    }
    // memory strips of a slot that a step does not write must be zero: slots were zero-filled at creation and
    // every step writes the same (region x PML) cells, so nothing else to do.
    for (i64 s = b + 1; s <= e; s++) TRY(el_step_forward(P, s, win_ptr(P, b, s - 1), win_ptr(P, b, s), true));
    P->win_base = b; P->win_last = e;
  }
  if (P->ll && nseg > 0) TRY(el_finish_slot(P, P->seg_e[nseg - 1] - P->seg_b[nseg - 1]));
  P->inplace_last = false;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_plan_forward(adseis_elastic_plan* P) {
  REQUIRE(P, "elastic_plan_forward: null");
  TRY(el_check_ready(P, false));
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  P->last_launches = 0; P->last_recomputed = 0;
  // keep the history when all of it fits (so that snapshots can be fetched); otherwise step in place
  TRY(el_forward_sweep(P, P->win >= P->p.NSTEP + 1, false));
  P->last_segments = 1;
  P->have_fwd = true;
  return ADSEIS_OK;
}

static int el_ensure_adjoint(adseis_elastic_plan* P, bool mat) {
  cudaStream_t st = P->ctx->stream;
  const ElGeom& g = P->g;
  if (!P->adj) TRY(dev_alloc_zero(&P->adj, (size_t)(5 * g.plane + 2 * (4 * g.xm_sz + 4 * g.ym_sz)), st));
  if (!P->gradsrcv) TRY(dev_alloc_zero(&P->gradsrcv, (size_t)(P->p.NSTEP * P->nsrc), st));
  if (mat && !P->Gl) {
    double** a[] = {&P->Gl, &P->Gm1, &P->Gm2, &P->Gr3, &P->Gr4};
    for (double** x : a) TRY(dev_alloc_zero(x, (size_t)g.plane, st));
  }
  if (mat && !P->grho) {
    TRY(dev_alloc_zero(&P->grho, (size_t)P->model_elems, st));
    TRY(dev_alloc_zero(&P->glam, (size_t)P->model_elems, st));
    TRY(dev_alloc_zero(&P->gmu, (size_t)P->model_elems, st));
  }
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_plan_gradient(adseis_elastic_plan* P, int want_material_grads) {
  REQUIRE(P, "elastic_plan_gradient: null");
  TRY(el_check_ready(P, true));
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  const ElGeom& g = P->g;
  cudaStream_t st = P->ctx->stream;
  const i64 NSTEP = P->p.NSTEP;
  const bool mat = want_material_grads != 0;
  P->last_launches = 0; P->last_recomputed = 0;
  TRY(el_ensure_adjoint(P, mat));
  // ---- forward: with history only when material gradients are wanted ----
  TRY(el_forward_sweep(P, mat, mat));
  const size_t nseg = mat ? P->seg_b.size() : 1;
  P->last_segments = (i64)nseg;
  const i64 nr = (NSTEP + 1) * P->nrcv;
  k_residual_partial<<<RL_BLOCKS, 256, 0, st>>>(P->rcvv, P->obs, P->rcv_owned, (int)std::max<i64>(P->nrcv, 1), 0, nr,
                                               P->res, P->loss);
  k_residual_final<<<1, RL_BLOCKS, 0, st>>>(P->loss);
  EL_LAUNCH_CHECK(P);
  // ---- reverse sweep ----
  const size_t msz = (size_t)(4 * g.xm_sz + 4 * g.ym_sz);
  CUDA_TRY(cudaMemsetAsync(P->adj, 0, (size_t)(5 * g.plane + 2 * msz) * 8, st));
  if (P->nsrc > 0) CUDA_TRY(cudaMemsetAsync(P->gradsrcv, 0, (size_t)(NSTEP * P->nsrc) * 8, st));
  if (mat) {
    double* a[] = {P->Gl, P->Gm1, P->Gm2, P->Gr3, P->Gr4};
    for (double* x : a) CUDA_TRY(cudaMemsetAsync(x, 0, (size_t)g.plane * 8, st));
  }
  ElSlot side[2];
  for (int k = 0; k < 2; k++) {
    side[k] = slot_at(P, P->adj);  // fields are shared (updated in place) ...
    side[k].xm = P->adj + 5 * g.plane + k * msz;  // ... the adjoint memories ping-pong
    side[k].ym = side[k].xm + 4 * g.xm_sz;
  }
  const ElMat mt = mat_of(P);
  const ElCoef cf = coef_of(P);
  const int stride = (int)(NSTEP + 1);
  ElPoints none{};
  TRY(el_barrier(P));  // slab plans: the memsets above are done everywhere before anybody pushes adjoint halo rows
  // start: velocity residuals of slot NSTEP into vbar; grad_srcv row NSTEP-1
  el_adj_start<<<P->nblocks, 256, 0, st>>>(g, side[0], P->rcv.dev, P->nrcv > 0 ? P->res : nullptr, stride, (int)NSTEP,
                                           P->src.dev, P->nsrc > 0 ? P->gradsrcv + (NSTEP - 1) * P->nsrc : nullptr);
  EL_LAUNCH_CHECK(P);
  const ElPlaneRef vb_planes[2] = {{EA_ADJ, 0, 0}, {EA_ADJ, 0, 1}};               // vbar_x, vbar_y
  const ElPlaneRef sb_planes[3] = {{EA_ADJ, 0, 2}, {EA_ADJ, 0, 3}, {EA_ADJ, 0, 4}};  // sigma_bar xx, yy, xy
  TRY(el_halo_exchange(P, 2, vb_planes));
  const size_t sb = (size_t)P->slot_sz * 8;
  ElSlot zero{};
  for (i64 k = (i64)nseg - 1; k >= 0; k--) {
    i64 b = 0, e = NSTEP;
    if (mat) {
      b = P->seg_b[k]; e = P->seg_e[k];
      if (k != (i64)nseg - 1) {  // restore the first slot of segment k and replay its forward steps
        if (k == 0) CUDA_TRY(cudaMemsetAsync(P->hist, 0, sb, st));
        else CUDA_TRY(cudaMemcpyAsync(P->hist, P->ckpt[k - 1], sb, cudaMemcpyDeviceToDevice, st));
        if (P->ll) TRY(el_barrier(P));   // handshake: the neighbour is done with the words of my last adjoint launches
        P->ll_prev_nf = 0;   // the previous fused launch was an adjoint one: nothing to receive
        for (i64 s = b + 1; s <= e; s++) TRY(el_step_forward(P, s, win_ptr(P, b, s - 1), win_ptr(P, b, s), false));
        P->last_recomputed += e - b;
        P->win_base = b; P->win_last = e;
        if (P->ll && P->arena) {
          TRY(el_finish_slot(P, e - b));
          TRY(el_halo_exchange(P, 2, vb_planes));   // vbar rows sent by the last adjoint launch before the replay
        }
      }
    }
    for (i64 s = e; s >= b + 1; s--) {
      const ElSlot& bi = side[s & 1];
      const ElSlot& bo = side[(s & 1) ^ 1];
      const ElSlot fs = mat ? slot_at(P, win_ptr(P, b, s)) : zero;
      const ElSlot fp = mat ? slot_at(P, win_ptr(P, b, s - 1)) : zero;
      const double* resp = P->nrcv > 0 ? P->res : nullptr;
      double* grow = (s - 2 >= 0 && P->nsrc > 0 && s >= 2) ? P->gradsrcv + (s - 2) * P->nsrc : nullptr;
      if (mat) {
        CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_slab(), el_vel_adj<true>, P->nblocks, EL_NT, el_ring_bytes<ElVelAdjT<true>>(), st, g, P->ctas, bi, bo, fs, mt, cf, P->Gr3, P->Gr4, P->rcv.dev, resp, stride, (int)s,
                                                     el_make_fuse(P, 3, sb_planes, true)));
        EL_LAUNCH_CHECK(P);
        CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_slab(), el_sigma_adj<true>, P->nblocks, EL_NT, el_ring_bytes<ElSigAdjT<true>>(), st, g, P->ctas, bi, bo, fp, fs, mt, cf, P->Gl, P->Gm1, P->Gm2,
                                                       s >= 2 ? P->rcv.dev : none, s >= 2 ? resp : nullptr, stride,
                                                       (int)(s - 1), s >= 2 ? P->src.dev : none, grow,
                                                       el_make_fuse(P, 2, vb_planes, true)));
        EL_LAUNCH_CHECK(P);
      } else {
        CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_slab(), el_vel_adj<false>, P->nblocks, EL_NT, el_ring_bytes<ElVelAdjT<false>>(), st, g, P->ctas, bi, bo, zero, mt, cf, nullptr, nullptr, P->rcv.dev, resp, stride,
                                                      (int)s, el_make_fuse(P, 3, sb_planes, true)));
        EL_LAUNCH_CHECK(P);
        CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_slab(), el_sigma_adj<false>, P->nblocks, EL_NT, el_ring_bytes<ElSigAdjT<false>>(), st, g, P->ctas, bi, bo, zero, zero, mt, cf, nullptr, nullptr, nullptr,
                                                        s >= 2 ? P->rcv.dev : none, s >= 2 ? resp : nullptr, stride,
                                                        (int)(s - 1), s >= 2 ? P->src.dev : none, grow,
                                                        el_make_fuse(P, 2, vb_planes, true)));
        EL_LAUNCH_CHECK(P);
      }
    }
  }
  if (mat) {
    CUDA_TRY(cudaMemsetAsync(P->grho, 0, (size_t)P->model_elems * 8, st));
    CUDA_TRY(cudaMemsetAsync(P->glam, 0, (size_t)P->model_elems * 8, st));
    CUDA_TRY(cudaMemsetAsync(P->gmu, 0, (size_t)P->model_elems * 8, st));
    if (P->p.variant == 0) {  // un-averaging reads the accumulator row above my first owned row
      const ElPlaneRef gp[3] = {{EA_GACC, 0, 0}, {EA_GACC, 0, 1}, {EA_GACC, 0, 4}};
      TRY(el_halo_exchange(P, 3, gp));
    }
    dim3 gg((unsigned)((g.W + 127) / 128), (unsigned)(g.own1 - g.own0));
    k_el_grad_finalize<<<gg, 128, 0, st>>>(g, P->p.variant == 0, P->off, P->Gl, P->Gm1, P->Gm2, P->Gr3, P->Gr4, P->grho,
                                           P->glam, P->gmu);
    EL_LAUNCH_CHECK(P);
  }
  P->have_fwd = true;
  P->have_grad = true;
  P->have_matgrad = mat;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_plan_get(adseis_elastic_plan* P, int what, double* dst, int to_device) {
  (void)to_device;
  REQUIRE(P && dst, "elastic_plan_get: null");
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  const double* src = nullptr;
  size_t n = 0;
  bool need_grad = true, need_mat = false;
  switch (what) {
    case ADSEIS_GET_RCVV: src = P->rcvv; n = (size_t)((P->p.NSTEP + 1) * P->nrcv); need_grad = false; break;
    case ADSEIS_GET_LOSS: src = P->loss; n = 1; break;
    case ADSEIS_GET_GRAD_SRCV: src = P->gradsrcv; n = (size_t)(P->p.NSTEP * P->nsrc); break;
    case ADSEIS_GET_GRAD_RHO: src = P->grho; n = (size_t)P->model_elems; need_mat = true; break;
    case ADSEIS_GET_GRAD_LAMBDA: src = P->glam; n = (size_t)P->model_elems; need_mat = true; break;
    case ADSEIS_GET_GRAD_MU: src = P->gmu; n = (size_t)P->model_elems; need_mat = true; break;
    default: adseis_set_error("elastic_plan_get: unknown item %d", what); return ADSEIS_EINVAL;
  }
  if ((!need_grad && !P->have_fwd) || (need_grad && !P->have_grad) || (need_mat && !P->have_matgrad)) {
    adseis_set_error("elastic_plan_get: item %d has not been computed yet", what);
    return ADSEIS_ESTATE;
  }
  if (n > 0) CUDA_TRY(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDefault, P->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(P->ctx->stream));
  TRY(el_halo_check(P));
  return ADSEIS_OK;
}

// copy one field of a resident slot out in the caller's layout (stress planes get their pending sources added)
static int el_copy_field_out(adseis_elastic_plan* P, const double* slot_base, int field, i64 slot, double* dst,
                             double* scratch) {
  const ElGeom& g = P->g;
  cudaStream_t st = P->ctx->stream;
  const double* plane = slot_base + (i64)field * g.plane;
  if (field >= 2 && P->src.ps.nu > 0 && slot >= 1) {
    CUDA_TRY(cudaMemcpyAsync(scratch, plane, (size_t)g.plane * 8, cudaMemcpyDeviceToDevice, st));
    k_el_apply_stress_sources<<<(P->src.ps.nu + 127) / 128, 128, 0, st>>>(scratch, field, P->src.dev, P->src.ps.nu,
                                                                          P->srcv + (slot - 1) * P->nsrc);
    P->ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    plane = scratch;
  }
  const int off = P->off, OW = g.W - 2 * off, OH = g.H - 2 * off;
  CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)OW * 8, plane + (i64)off * g.ld + off, (size_t)g.ld * 8, (size_t)OW * 8,
                             (size_t)OH, cudaMemcpyDefault, st));
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_plan_get_snapshot(adseis_elastic_plan* P, int field, int64_t slot, double* dst,
                                                int to_device) {
  (void)to_device;
  REQUIRE(P && dst && field >= 0 && field <= 4, "elastic_plan_get_snapshot: bad argument");
  if (!P->have_fwd || slot < P->win_base || slot > P->win_last) {
    adseis_set_error("elastic_plan_get_snapshot: slot %lld is not resident (window holds [%lld,%lld])",
                     (long long)slot, (long long)P->win_base, (long long)P->win_last);
    return ADSEIS_ESTATE;
  }
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  double* scratch = nullptr;
  TRY(dev_alloc(&scratch, (size_t)P->g.plane));
  const double* base = P->inplace_last ? P->hist : win_ptr(P, P->win_base, slot);
  int rc = el_copy_field_out(P, base, field, slot, dst, scratch);
  cudaStreamSynchronize(P->ctx->stream);
  cudaFree(scratch);
  return rc;
}

ADSEIS_API int adseis_elastic_plan_info(adseis_elastic_plan* P, int64_t info[8]) {
  REQUIRE(P && info, "elastic_plan_info: null");
  info[0] = P->win; info[1] = P->last_segments; info[2] = P->last_launches; info[3] = P->g.Hl; info[4] = P->g.ld;
  info[5] = P->last_recomputed; info[6] = (i64)P->seg_b.size(); info[7] = P->slot_sz;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_plan_ipc_export(adseis_elastic_plan* P, void* handle_out) {
  REQUIRE(P && handle_out, "elastic_plan_ipc_export: null");
  if (!P->arena) {
    adseis_set_error("elastic_plan_ipc_export: not a slab plan (nranks == 1)");
    return ADSEIS_ESTATE;
  }
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, P->arena));
  static_assert(sizeof(h) == ADSEIS_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  return ADSEIS_OK;
}

ADSEIS_API int adseis_elastic_plan_ipc_connect(adseis_elastic_plan* P, const void* lo, const void* hi) {
  REQUIRE(P, "elastic_plan_ipc_connect: null");
  if (!P->arena) {
    adseis_set_error("elastic_plan_ipc_connect: not a slab plan (nranks == 1)");
    return ADSEIS_ESTATE;
  }
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  const void* hs[2] = {lo, hi};
  const bool need[2] = {P->slab.rank > 0, P->slab.rank < P->slab.nranks - 1};
  for (int k = 0; k < 2; k++) {
    if (!need[k]) continue;
    REQUIRE(hs[k], "elastic_plan_ipc_connect: missing handle of rank %d", P->slab.rank + (k ? 1 : -1));
    cudaIpcMemHandle_t h;
    memcpy(&h, hs[k], sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      adseis_set_error("elastic_plan_ipc_connect: cudaIpcOpenMemHandle(rank %d) -> %s", P->slab.rank + (k ? 1 : -1),
                       cudaGetErrorString(e));
      return ADSEIS_ECOMM;
    }
    P->peer[k] = (char*)ptr;
    CUDA_TRY(cudaMemcpy(&P->dpeer[k], ptr, sizeof(P->dpeer[k]), cudaMemcpyDeviceToHost));
    const adseis_elastic_plan::Desc& d = P->dpeer[k];
    if (d.magic != EL_DESC_MAGIC || d.ld != P->desc.ld || d.win != P->desc.win) {
      adseis_set_error("elastic_plan_ipc_connect: neighbour %d has an incompatible layout (magic %llx ld %lld win %lld; "
                       "mine ld %lld win %lld)", P->slab.rank + (k ? 1 : -1), d.magic, d.ld, d.win, P->desc.ld, P->desc.win);
      return ADSEIS_ECOMM;
    }
  }
  P->connected = true;
  return ADSEIS_OK;
}

// ------------------------------------------------------------------------------------------------------------
// one-call host-buffer entry points
// ------------------------------------------------------------------------------------------------------------
ADSEIS_API int adseis_elastic_forward(adseis_ctx* ctx, const adseis_elastic_params* p, const double* rho,
                                      const double* lambda, const double* mu, int64_t nsrc, const int64_t* srci,
                                      const int64_t* srcj, const int64_t* srctype, const double* srcv,
                                      int64_t srcv_rows, int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj,
                                      const int64_t* rcvtype, double* rcvv_out, double* hist_out) {
  adseis_elastic_plan* P = nullptr;
  // without a history request two slots are enough (in-place stepping uses one)
  TRY(adseis_elastic_plan_create(ctx, p, nullptr, nsrc, srci, srcj, srctype, nrcv, rcvi, rcvj, rcvtype,
                                 hist_out ? 0 : 1, &P));
  int rc = adseis_elastic_plan_set_model(P, rho, lambda, mu, 0);
  if (!rc) rc = adseis_elastic_plan_set_srcv(P, srcv, srcv_rows, 0);
  if (!rc && !hist_out) rc = adseis_elastic_plan_forward(P);
  if (!rc && hist_out) {
    // field histories for the caller: [5][(NSTEP+1)][model]; step in place and copy each slot out
    cudaStream_t st = ctx->stream;
    double* scratch = nullptr;
    rc = dev_alloc(&scratch, (size_t)P->g.plane);
    if (!rc) rc = el_check_ready(P, false);
    if (!rc) {
      cudaMemsetAsync(P->hist, 0, (size_t)P->slot_sz * 8, st);
      if (P->nrcv > 0) cudaMemsetAsync(P->rcvv, 0, (size_t)((p->NSTEP + 1) * P->nrcv) * 8, st);
      for (i64 s = 0; s <= p->NSTEP && !rc; s++) {
        if (s >= 1) rc = el_step_forward(P, s, P->hist, P->hist, true);
        for (int f = 0; f < 5 && !rc; f++)
          rc = el_copy_field_out(P, P->hist, f, s, hist_out + ((i64)f * (p->NSTEP + 1) + s) * P->model_elems, scratch);
      }
      P->have_fwd = (rc == 0);
    }
    cudaStreamSynchronize(st);
    cudaFree(scratch);
  }
  if (!rc && rcvv_out && nrcv > 0) rc = adseis_elastic_plan_get(P, ADSEIS_GET_RCVV, rcvv_out, 0);
  if (!rc) rc = adseis_ctx_sync(ctx);
  adseis_elastic_plan_destroy(P);
  return rc;
}

ADSEIS_API int adseis_elastic_misfit_grad(adseis_ctx* ctx, const adseis_elastic_params* p, const double* rho,
                                          const double* lambda, const double* mu, int64_t nsrc, const int64_t* srci,
                                          const int64_t* srcj, const int64_t* srctype, const double* srcv,
                                          int64_t srcv_rows, int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj,
                                          const int64_t* rcvtype, const double* obs, double* loss_out,
                                          double* rcvv_out, double* grad_rho_out, double* grad_lambda_out,
                                          double* grad_mu_out, double* grad_srcv_out) {
  adseis_elastic_plan* P = nullptr;
  const bool mat = grad_rho_out || grad_lambda_out || grad_mu_out;
  TRY(adseis_elastic_plan_create(ctx, p, nullptr, nsrc, srci, srcj, srctype, nrcv, rcvi, rcvj, rcvtype, mat ? 0 : 1,
                                 &P));
  int rc = adseis_elastic_plan_set_model(P, rho, lambda, mu, 0);
  if (!rc) rc = adseis_elastic_plan_set_srcv(P, srcv, srcv_rows, 0);
  if (!rc) rc = adseis_elastic_plan_set_obs(P, obs, 0);
  if (!rc) rc = adseis_elastic_plan_gradient(P, mat ? 1 : 0);
  if (!rc && loss_out) rc = adseis_elastic_plan_get(P, ADSEIS_GET_LOSS, loss_out, 0);
  if (!rc && rcvv_out && nrcv > 0) rc = adseis_elastic_plan_get(P, ADSEIS_GET_RCVV, rcvv_out, 0);
  if (!rc && grad_rho_out) rc = adseis_elastic_plan_get(P, ADSEIS_GET_GRAD_RHO, grad_rho_out, 0);
  if (!rc && grad_lambda_out) rc = adseis_elastic_plan_get(P, ADSEIS_GET_GRAD_LAMBDA, grad_lambda_out, 0);
  if (!rc && grad_mu_out) rc = adseis_elastic_plan_get(P, ADSEIS_GET_GRAD_MU, grad_mu_out, 0);
  if (!rc && grad_srcv_out && nsrc > 0) rc = adseis_elastic_plan_get(P, ADSEIS_GET_GRAD_SRCV, grad_srcv_out, 0);
  if (!rc) rc = adseis_ctx_sync(ctx);
  adseis_elastic_plan_destroy(P);
  return rc;
}
