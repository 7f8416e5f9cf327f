// acoustic.cu -- acoustic plan: device state, time loops (forward, checkpointed reverse sweep), C ABI.
// Reference behaviour restated: src/Core.jl:562-620 (loop), :622-655 (PML), :726-730 (receivers),
// src/Utils.jl:308 (misfit), src/MPIAcoustic.jl:334-431 (MPI-convention inputs); adjoint = SURVEY Appendix A.
#include <math.h>

#include "acoustic_kernels.cuh"
#include "util_kernels.cuh"

// ------------------------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------------------------
__global__ void k_square(double* __restrict__ dst, const double* __restrict__ src, i64 n) {
  i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[k] * src[k];
}

// whole-array point injection / sampling (used once per sweep for the last slot)
__global__ void k_points_inject(double* __restrict__ field, AcPoints ps, int nu, const double* __restrict__ val,
                                double scale) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nu) {
    const int cell = ps.cell[k];
    double v = field[cell];
    for (int m = ps.start[k]; m < ps.start[k + 1]; m++) v += val[ps.perm[m]] * scale;
    field[cell] = v;
  }
}
__global__ void k_points_sample(const double* __restrict__ field, AcPoints ps, int nu, double* __restrict__ out,
                                double scale) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nu) {
    const double v = field[ps.cell[k]] * scale;
    for (int m = ps.start[k]; m < ps.start[k + 1]; m++) out[ps.perm[m]] = v;
  }
}

// grad wrt the caller's model array from the accumulated d loss / d c^2 (pitched local rows)
//   mpi_convention=0: out[(gi)*(W)+j] = 2*c*G   (chain rule of Core.jl:564)     (dense padded)
//   mpi_convention=1: out[(gi-1)*NY + j-1] = G                                   (dense unpadded)
__global__ void k_grad_finalize(const double* __restrict__ G, const double* __restrict__ cvel, int ld, int goff,
                                int l0, int l1, int H, int W, int mpi, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int li = l0 + blockIdx.y;
  if (li >= l1 || j >= W) return;
  const int gi = goff + li;
  const i64 IJ = (i64)li * ld + j;
  if (!mpi) {
    out[(i64)gi * W + j] = 2.0 * cvel[IJ] * G[IJ];
  } else if (gi >= 1 && gi <= H - 2 && j >= 1 && j <= W - 2) {
    out[(i64)(gi - 1) * (W - 2) + (j - 1)] = G[IJ];
  }
}

// ------------------------------------------------------------------------------------------------------------
// slab decomposition: peer-memory halo exchange over NVLink (one process per GPU, CUDA IPC)
// ------------------------------------------------------------------------------------------------------------
// Every slab plan keeps the arrays that have halo rows in ONE device arena whose IPC handle is exported; the
// descriptor at the start of the arena tells a neighbour where each array lives.
struct AcDesc {
  unsigned long long magic;
  long long Hl, ld, plane, win, own0, own1;
  long long off_flags, off_hist, off_phi[2], off_psi[2], off_ub[4], off_phib[2], off_psib[2];
  long long n_edge_lo, n_edge_hi;  // CTAs of one step launch that own cells of my first / last owned row
  long long n_edge_lo_f, n_edge_hi_f;  // ... of a frame-only launch of the two-step path (0: that path is off)
  long long n_edge_lo_w, n_edge_hi_w;  // ... of a WIDE frame-only launch (frame + rim ring of the box)
  long long nub;
  long long off_ll;   // packed halo rows: [from lo, from hi][u, p][parity] rows of ld 16-byte words (0: not present)
};
#define AC_DESC_MAGIC 0xAD5E15B200ULL
#define AC_HX_BLOCKS 8
#define AC_HX_SPIN_LIMIT (1ULL << 26)

struct AcHaloArgs {
  int nrow;                 // rows to push (0 = pure barrier)
  const double* src[12];    // my rows (ld doubles each)
  double* dst[12];          // the neighbour's halo rows (peer pointers)
  int ld;
  int has_lo, has_hi;
  unsigned long long* sig_lo;   // neighbour's "flag_from_hi" / "flag_from_lo" (peer pointers)
  unsigned long long* sig_hi;
  unsigned long long* my_flags; // [0] = written by rank-1, [1] = written by rank+1, [2] = error
  unsigned long long expect;    // AC_HX_BLOCKS * epoch
};

// Push my edge rows into the neighbours' halo rows (plain 16-byte stores over NVLink), publish them with a
// system-scope fence + atomic on the neighbour's flag, then wait until both neighbours have published theirs.
// All ranks run the same sequence of exchanges, so a single monotonically increasing epoch orders everything.
__global__ void __launch_bounds__(256) k_halo_exchange(AcHaloArgs a) {
  const int per = (a.ld / 2 + AC_HX_BLOCKS - 1) / AC_HX_BLOCKS;  // double2 elements per CTA
  const int j0 = blockIdx.x * per, j1 = min(a.ld / 2, j0 + per);
  for (int r = 0; r < a.nrow; r++) {
    const double2* s = reinterpret_cast<const double2*>(a.src[r]);
    double2* d = reinterpret_cast<double2*>(a.dst[r]);
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) d[j] = s[j];
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
        if (++spins > AC_HX_SPIN_LIMIT) { f[2] = 1ULL; break; }  // neighbour lost: report, do not hang
      }
      __threadfence_system();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------------------
struct adseis_acoustic_plan {
  adseis_ctx* ctx;
  adseis_acoustic_params p;
  adseis_slab slab;
  AcGeom g;
  AcTiling t;
  int nblocks = 0;   // CTAs per time-step launch
  // temporal blocking (two forward steps per launch; single-GPU plans with a large enough PML-free box):
  //   t2 = marching-only tiling of the box shrunk by two cells, tf = frame-only tiling of everything else
  bool tb = false;
  AcTiling t2{};
  int nblocks2 = 0;
  int box_i0 = 0, box_i1 = -1, box_j0 = 0, box_j1 = -1;  // PML-free box (global padded indices, inclusive)
  // frame-only tilings of the two-step path: [0] = everything outside the box, [1] = the same plus the rim ring of the box
  // (a "wide" launch: the next frame launch then depends on it alone, not on the concurrent box-pair launch)
  struct FrameKind {
    AcTiling t{};
    int nblocks = 0;
    int* perm = nullptr;
    int n_edge_lo = 0, n_edge_hi = 0;
    PointSetStorage src, rcv;       // all points the tiling owns
    AcPoints srcp{}, rcvp{};
    // wide kind only: the points outside the box (epilogue injection) and the points on the rim ring of the box
    // (injected in registers, see AcFuse::rim)
    PointSetStorage srcN, rcvN, srcR, rcvR;
    AcPoints srcNp{}, rcvNp{}, srcRp{}, rcvRp{};
  };
  FrameKind fk[2];
  bool tb_overlap = false;                               // box-pair launches and frame launches on two streams
  cudaStream_t sb = nullptr;                             // the frames' stream
  cudaEvent_t evA[2] = {nullptr, nullptr}, evB[2] = {nullptr, nullptr};
  PointSetStorage srcM, rcvM, srcH;                      // points by owner under t2 (box) / t2 + rim
  AcPoints srcMp{}, rcvMp{}, srcHp{};
  int fast_rows = 0; // rows of the PML-free box owned by this GPU
  int own0, own1;  // owned local rows [own0, own1)
  i64 model_elems; // elements of the caller's model array
  // static fields (pitched, local rows)
  double *c2 = nullptr, *cvel = nullptr, *sigx = nullptr, *tauy = nullptr;
  double *phi[2] = {nullptr, nullptr}, *psi[2] = {nullptr, nullptr};
  // history window + checkpoints
  double* hist = nullptr;
  i64 win = 0;  // window capacity in slots
  std::vector<i64> seg_b, seg_e;
  std::vector<double*> ckpt;  // per segment (index k>=1): 4 planes
  i64 win_base = -1, win_last = -1;  // slots currently valid in the window: [win_base, win_last]
  // points
  i64 nsrc = 0, nrcv = 0;
  i64 points_gen = 0;   // bumped whenever a per-point device buffer is reallocated (invalidates captured graphs)
  // CUDA graphs (single-GPU plans with small step kernels): the whole launch sequence of forward() / gradient() is
  // captured once and replayed -- every kernel argument is a device address that stays fixed between calls (the
  // history window, the per-step rows of srcv / rcvv / res / gradsrcv, the per-CTA point lists, which set_points
  // rewrites in place).  `key` fingerprints those addresses; a mismatch re-captures.
  struct SpanRec { int phase; i64 launches; cudaEvent_t a, b; };
  struct GraphSlot { cudaGraphExec_t exec = nullptr; unsigned long long key = 0; i64 launches = 0, recomputed = 0; size_t ev_used = 0; std::vector<SpanRec> spans_copy; };
  GraphSlot gslot[2];       // 0 = forward(), 1 = gradient()
  int graphs = -1;          // -1 undecided, 0 off, 1 on
  bool capturing = false;
  PointSetStorage src, rcv;
  AcPoints srcp{}, rcvp{};
  unsigned char* rcv_owned = nullptr;
  double *srcv = nullptr, *rcvv = nullptr, *obs = nullptr, *res = nullptr, *loss = nullptr;
  i64 srcv_rows = 0;
  // adjoint state
  double *ub[4] = {nullptr, nullptr, nullptr, nullptr}, *phib[2] = {nullptr, nullptr}, *psib[2] = {nullptr, nullptr};
  int nub = 3;       // rotating ubar planes: 3, or 4 when the two-step adjoint kernel runs (it writes two time levels)
  bool tb_adj = false;  // two adjoint steps per launch (tb && PropagatorKernel != 0)
  PointSetStorage rcvH;  // receivers under every box tile whose one-cell rim holds them (two-step adjoint)
  AcPoints rcvHp{};
  double *G = nullptr, *gradc = nullptr, *gradsrcv = nullptr;
  double* ut[2] = {nullptr, nullptr};  // PropagatorKernel=0: utilde planes (adjoint of the pre-injection outputs)
  bool k0_corr = false;                // ... and some source touches a cell whose phi/psi coefficient is non-zero
  bool have_model = false, have_srcv = false, have_obs = false, have_grad = false, have_fwd = false;
  // slab decomposition (nranks > 1): arena, neighbours
  double* arena = nullptr;
  size_t arena_bytes = 0;
  AcDesc desc{};
  AcDesc dpeer[2];                           // descriptors of rank-1 / rank+1
  char* peer[2] = {nullptr, nullptr};        // their arenas mapped into this process
  unsigned long long epoch = 0;              // exchanges issued so far
  unsigned long long sepoch = 0;             // step kernels launched so far (same sequence on every rank)
  int* perm = nullptr;                       // launch order -> logical CTA id (edge CTAs first)
  int n_edge_lo = 0, n_edge_hi = 0;
  // whole sweep in one cooperative launch (small single-GPU grids whose tape is resident): tiling of ALL cells as general
  // cells, the points grouped by its CTAs, the barrier counter
  bool persist = false;
  AcTiling tp{};
  int nblocksP = 0;
  PointSetStorage srcP{}, rcvP{};
  unsigned long long* pbar = nullptr;        // [0] grid barrier counter, [1 ...] per-CTA progress
  int hd = 1;                                // halo rows per neighbour (2 for PropagatorKernel = 0, whose phi'/psi' read u' one row further)
  bool unfused = false;                      // slab plan whose step launches do not exchange: an explicit exchange follows every launch
  PointSetStorage srcK{};                    // PropagatorKernel = 0 on slabs: sources within one row of my rows (c-gradient correction)
  bool ll = false;                           // packed halo rows instead of fence + flag (ADSEIS_AC_LL=0 switches back)
  int ll_prev_kind = 0; i64 ll_prev_s = 0;   // the last fused launch: 1 forward / 2 adjoint step s (0: none, or an explicit exchange since)
  unsigned long long exp_lo = 0, exp_hi = 0; // signals the neighbours have sent me so far (sum over their fused launches)
  bool connected = false;
  // stats
  i64 last_launches = 0, last_segments = 0, last_recomputed = 0;
  // per-phase device timing of the last forward()/gradient(): CUDA events on the ctx stream around each run of
  // identical kernels (phase 0 = forward sweep, 1 = forward recomputation, 2 = adjoint sweep)
  std::vector<cudaEvent_t> ev_pool;
  typedef SpanRec Span;
  std::vector<Span> spans;
  size_t ev_used = 0;
};

static int span_begin(adseis_acoustic_plan* P, int phase, i64 launches) {
  while (P->ev_pool.size() < P->ev_used + 2) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    P->ev_pool.push_back(e);
  }
  adseis_acoustic_plan::Span sp{phase, launches, P->ev_pool[P->ev_used], P->ev_pool[P->ev_used + 1]};
  P->ev_used += 2;
  CUDA_TRY(cudaEventRecordWithFlags(sp.a, P->ctx->stream, P->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  P->spans.push_back(sp);
  return ADSEIS_OK;
}
static int span_end(adseis_acoustic_plan* P) {
  CUDA_TRY(cudaEventRecordWithFlags(P->spans.back().b, P->ctx->stream, P->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  return ADSEIS_OK;
}

static inline double* win_slot(adseis_acoustic_plan* P, i64 base, i64 s) { return P->hist + (s - base) * P->g.plane; }

#define LAUNCH_CHECK(P)                   \
  do {                                    \
    (P)->ctx->launches++;                 \
    (P)->last_launches++;                 \
    CUDA_TRY(cudaGetLastError());         \
  } while (0)

// Development aid (-DADSEIS_TIMELINE): a timed event after every step launch on the stream it went to; the dump lists,
// per launch, when it finished.  ADSEIS_TIMELINE=<marks to keep>, ADSEIS_TIMELINE_SKIP=<launches to skip first>.
#ifdef ADSEIS_TIMELINE
struct TlMark { cudaEvent_t ev; int tag; long long s; unsigned long long* dev; };
static std::vector<TlMark> g_tl;
static long long g_tl_seen = 0, g_tl_cap = -1, g_tl_skip = 0;
static unsigned long long* g_tl_dev = nullptr;   // 5 device time stamps per mark
static unsigned long long* g_tl_next = nullptr;  // slot handed to the launch that tl_mark() follows
static void tl_init() {
  if (g_tl_cap >= 0) return;
  const char* e = getenv("ADSEIS_TIMELINE"); g_tl_cap = e ? atoll(e) : 0;
  const char* k = getenv("ADSEIS_TIMELINE_SKIP"); g_tl_skip = k ? atoll(k) : 0;
  if (g_tl_cap > 0 && cudaMalloc(&g_tl_dev, (size_t)g_tl_cap * 5 * 8) == cudaSuccess) {
    std::vector<unsigned long long> init((size_t)g_tl_cap * 5, 0ULL);
    for (long long i = 0; i < g_tl_cap; i++) init[(size_t)i * 5] = ~0ULL;
    cudaMemcpy(g_tl_dev, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
  }
}
// device slot for the NEXT marked launch (null when that launch is not kept)
static unsigned long long* tl_slot() {
  tl_init();
  g_tl_next = nullptr;
  if (g_tl_seen < g_tl_skip || (long long)g_tl.size() >= g_tl_cap || !g_tl_dev) return nullptr;
  return g_tl_next = g_tl_dev + g_tl.size() * 5;
}
static void tl_mark(cudaStream_t st, int tag, long long s) {
  tl_init();
  if (g_tl_seen++ < g_tl_skip || (long long)g_tl.size() >= g_tl_cap) return;
  static int with_events = -1;
  if (with_events < 0) { const char* e = getenv("ADSEIS_TIMELINE_EVENTS"); with_events = (e && e[0] == '0') ? 0 : 1; }
  cudaEvent_t ev = nullptr;
  if (with_events) {   // (an event record between two launches costs ~6 us of stream time: off for undisturbed device stamps)
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, st);
  }
  g_tl.push_back({ev, tag, s, g_tl_next});
  g_tl_next = nullptr;
}
extern "C" __attribute__((visibility("default"))) int adseis_debug_timeline_dump(const char* path) {
  cudaDeviceSynchronize();
  FILE* f = fopen(path, "w");
  if (!f) return -1;
  unsigned long long t00 = 0;
  for (size_t k = 0; k < g_tl.size(); k++) {
    float ms = 0.f;
    if (g_tl[0].ev && g_tl[k].ev) cudaEventElapsedTime(&ms, g_tl[0].ev, g_tl[k].ev);
    unsigned long long d[5] = {0, 0, 0, 0, 0};
    if (g_tl[k].dev) cudaMemcpy(d, g_tl[k].dev, sizeof(d), cudaMemcpyDeviceToHost);
    if (!t00 && g_tl[k].dev) t00 = d[0];
    double r[5];
    for (int q = 0; q < 5; q++) r[q] = (g_tl[k].dev && d[q] && d[q] != ~0ULL) ? (double)(long long)(d[q] - t00) * 1e-3 : -1.0;
    fprintf(f, "%zu %d %lld %.3f %.3f %.3f %.3f %.3f %.3f\n", k, g_tl[k].tag, g_tl[k].s, ms * 1e3, r[0], r[1], r[2], r[3], r[4]);
  }
  fclose(f);
  return (int)g_tl.size();
}
#define TL_MARK(st, tag, s) tl_mark(st, tag, s)
#else
#define TL_MARK(st, tag, s) do {} while (0)
#endif

enum AcArr { AR_HIST, AR_PHI, AR_PSI, AR_UB, AR_PHIB, AR_PSIB };  // arrays with halo rows (slab plans)
static int halo_exchange(adseis_acoustic_plan* P, int narr, const int* arr, const i64* idx, const int* depth = nullptr);
static int halo_check(adseis_acoustic_plan* P);
struct AcDesc;
static long long desc_off(const AcDesc& d, int arr, i64 idx);

static int plan_segments(adseis_acoustic_plan* P, size_t budget_bytes, bool count_checkpoints) {
  // The wavefield history window holds `win` snapshots.  If NSTEP+1 snapshots fit, the whole history is kept;
  // otherwise the reverse sweep proceeds segment by segment (win-2 steps each) from checkpoints of 4 planes.
  // count_checkpoints: the checkpoints must also fit into budget_bytes (auto mode); else only the window does.
  const i64 NSTEP = P->p.NSTEP;
  const size_t plane_bytes = (size_t)P->g.plane * sizeof(double);
  i64 slots = (i64)(budget_bytes / plane_bytes);
  if (slots >= NSTEP + 1) {
    P->win = NSTEP + 1;
  } else {
    i64 best = -1;
    for (i64 W = slots; W >= 4; W--) {
      i64 nseg = (NSTEP - 1 + (W - 2) - 1) / (W - 2);
      if (!count_checkpoints || W + 4 * (nseg - 1) <= slots) { best = W; break; }
    }
    if (best < 0) {
      adseis_set_error("acoustic plan: a history budget of %zu bytes (%lld snapshots of %zu bytes) is too small",
                       budget_bytes, (long long)slots, plane_bytes);
      return ADSEIS_ENOMEM;
    }
    P->win = best;
  }
  // The LAST segment is never replayed (its snapshots are still in the window when the reverse sweep starts), so
  // it is the one that gets the full window; the remainder goes to the first segment.
  P->seg_b.clear(); P->seg_e.clear();
  const i64 per = P->win - 2;                                   // new steps per full segment
  const i64 nseg = std::max<i64>(1, (NSTEP - 1 + per - 1) / per);
  i64 b = 0, e = std::min(NSTEP, 1 + (NSTEP - 1) - (nseg - 1) * per);
  while (true) {
    P->seg_b.push_back(b); P->seg_e.push_back(e);
    if (e >= NSTEP) break;
    b = e - 1;
    e = std::min(b + P->win - 1, NSTEP);
  }
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_destroy(adseis_acoustic_plan* P) {
  if (!P) return ADSEIS_OK;
  cudaSetDevice(P->ctx->device);
  cudaStreamSynchronize(P->ctx->stream);
  cudaFree(P->c2); cudaFree(P->cvel); cudaFree(P->sigx); cudaFree(P->tauy);
  cudaFree(P->perm);
  for (int k = 0; k < 2; k++) {
    cudaFree(P->fk[k].perm); free_point_set(&P->fk[k].src); free_point_set(&P->fk[k].rcv);
    free_point_set(&P->fk[k].srcN); free_point_set(&P->fk[k].rcvN); free_point_set(&P->fk[k].srcR); free_point_set(&P->fk[k].rcvR);
  }
  if (P->sb) cudaStreamDestroy(P->sb);
  for (int k = 0; k < 2; k++) { if (P->evA[k]) cudaEventDestroy(P->evA[k]); if (P->evB[k]) cudaEventDestroy(P->evB[k]); }
  if (P->arena) {
    for (int k = 0; k < 2; k++) if (P->peer[k]) cudaIpcCloseMemHandle(P->peer[k]);
    cudaFree(P->arena);
  } else {
    for (int k = 0; k < 2; k++) { cudaFree(P->phi[k]); cudaFree(P->psi[k]); cudaFree(P->phib[k]); cudaFree(P->psib[k]); }
    for (int k = 0; k < 4; k++) cudaFree(P->ub[k]);
    cudaFree(P->hist);
  }
  for (double* c : P->ckpt) cudaFree(c);
  free_point_set(&P->src); free_point_set(&P->rcv); free_point_set(&P->srcK);
  free_point_set(&P->srcP); free_point_set(&P->rcvP); cudaFree(P->pbar);
  free_point_set(&P->srcM); free_point_set(&P->rcvM);
  free_point_set(&P->srcH); free_point_set(&P->rcvH);
  cudaFree(P->rcv_owned);
  cudaFree(P->srcv); cudaFree(P->rcvv); cudaFree(P->obs); cudaFree(P->res); cudaFree(P->loss);
  cudaFree(P->G); cudaFree(P->gradc); cudaFree(P->gradsrcv); cudaFree(P->ut[0]); cudaFree(P->ut[1]);
  for (int k = 0; k < 2; k++) if (P->gslot[k].exec) cudaGraphExecDestroy(P->gslot[k].exec);
  for (cudaEvent_t e : P->ev_pool) cudaEventDestroy(e);
  adseis_ctx* ctx = P->ctx;
  delete P;
  adseis_ctx_release_plan(ctx);
  return ADSEIS_OK;
}

static int validate_params(const adseis_acoustic_params* p) {
  REQUIRE(p, "acoustic: null params");
  REQUIRE(p->NX >= 3 && p->NY >= 3 && p->NSTEP >= 2, "acoustic: need NX,NY >= 3 and NSTEP >= 2 (got %lld,%lld,%lld)",
          (long long)p->NX, (long long)p->NY, (long long)p->NSTEP);
  REQUIRE((p->NX + 2) * (p->NY + 18) < 2147483647LL, "acoustic: grid too large for 32-bit cell offsets");
  REQUIRE(p->DELTAX > 0 && p->DELTAY > 0 && p->DELTAT > 0, "acoustic: DELTAX/DELTAY/DELTAT must be > 0");
  REQUIRE(p->NPOINTS_PML >= 1 && p->Rcoef > 0 && p->vp_ref > 0, "acoustic: bad PML parameters");
  REQUIRE(p->PropagatorKernel >= 0 && p->PropagatorKernel <= 2,
          "acoustic: PropagatorKernel=%d not supported (0 = TF-op scheme, Core.jl:528-549; 1 = custom-op scheme; "
          "2 is numerically identical to 1)", p->PropagatorKernel);
  return ADSEIS_OK;
}

// Launch order of a tiling (logical CTA id = perm[blockIdx.x]).  Slab plans: the CTAs that own cells of my first / last
// owned row next to a neighbour go first, so their halo pushes leave early and the rest of the step hides the NVLink
// latency.  Then the frame CTAs (latency-bound: one DRAM round trip + fp64 divides, no bandwidth) are spread evenly
// among the marching CTAs (bandwidth-bound) instead of forming the tail of the launch.
static void build_launch_order(adseis_acoustic_plan* P, const AcTiling& t, int nblocks, std::vector<int>* order_out,
                               int* n_edge_lo, int* n_edge_hi) {
  const bool halo_lo = P->slab.rank > 0, halo_hi = P->slab.rank < P->slab.nranks - 1;
  const int first = P->own0, last = P->own1 - 1;
  const int fcells = t.fthr * t.fcpt;
  std::vector<int> edge, march, frame;
  *n_edge_lo = *n_edge_hi = 0;
  for (int b = 0; b < nblocks; b++) {
    int rlo, rhi;
    if (b < t.nmarch) {
      int a0, a1;
      ac_row_tile(t, b / t.nct, &a0, &a1);
      rlo = a0; rhi = a1 - 1;
    } else {
      int fb = b - t.nmarch, r = 0;
      for (int k = 1; k < t.nrect; k++) if (fb >= t.rblk[k]) r = k;
      const int w = t.rc1[r] - t.rc0[r];
      const i64 ncell = (i64)(t.rr1[r] - t.rr0[r]) * w, i0 = (i64)(fb - t.rblk[r]) * fcells;
      rlo = t.rr0[r] + (int)(i0 / w);
      rhi = t.rr0[r] + (int)((std::min<i64>(ncell, i0 + fcells) - 1) / w);
    }
    const bool tl = halo_lo && rlo <= first && first <= rhi, th = halo_hi && rlo <= last && last <= rhi;
    if (tl) (*n_edge_lo)++;
    if (th) (*n_edge_hi)++;
    if (tl || th) edge.push_back(b);
    else (b < t.nmarch ? march : frame).push_back(b);
  }
  std::vector<int>& order = *order_out;
  order = edge;
  size_t im = 0, ifr = 0;
  const size_t nm = march.size(), nf = frame.size();
  while (im < nm || ifr < nf) {  // Bresenham merge: frame CTA k goes after ~k*nm/nf marching CTAs
    if (ifr < nf && (im >= nm || ifr * nm <= im * nf)) order.push_back(frame[ifr++]);
    else order.push_back(march[im++]);
  }
}

// Temporal blocking: tilings of the two-step forward path (see ac_fwd2_kernel).  Single-GPU plans only; the marched box
// is the PML-free box shrunk by TWO cells so that the one-cell rim of every tile is made of plain interior cells.
static void build_tb_tilings(adseis_acoustic_plan* P) {
  P->tb = false;
  const char* e = getenv("ADSEIS_AC_TB");   // "0": never, "1": whenever the box is large enough to tile, unset: auto
  if (e && e[0] == '0') return;
  const AcGeom& g = P->g;
  const bool slab = P->slab.nranks > 1;
  const int has_lo = P->slab.rank > 0 ? 1 : 0, has_hi = P->slab.rank < P->slab.nranks - 1 ? 1 : 0;
  if (slab) {
    // Slab plans: the two rows next to a neighbour are advanced by the frame-only launches (which carry the fused halo
    // exchange of the one-step kernels, one halo row, unchanged); the box pair kernel then reads owned rows only and
    // never touches a halo.  EVERY rank must take the same decision (the launch sequences must match): it depends on
    // global quantities only.
    const char* es = getenv("ADSEIS_AC_TB_SLAB");
    if (es && es[0] == '0') return;
    if (P->p.PropagatorKernel == 0) return;
    const i64 rows_min = (i64)g.H / P->slab.nranks - (P->p.NPOINTS_PML + 4) - 5;
    const i64 cells = rows_min * (i64)(g.W - 2 * (P->p.NPOINTS_PML + 4));
    // measured on B200s with packed halo rows and 128-thread frame-only CTAs (profiles/r02_slab_latency.md), C4 (4096^2),
    // Gcell-upd/s one step per launch -> pairs: 2 GPUs (8.3 M box cells per slab) 103.5 -> 126+, 4 GPUs (4.1 M) 197 -> 227,
    // 8 GPUs (2.0 M) 406 -> 357: an interior slab waits on two neighbours in every frame-only launch, and its box-pair kernel
    // has 14 rows per CTA (two of them rim rows) -- below ~3 M cells one step per launch wins
    const i64 min_cells = getenv("ADSEIS_AC_TB_SLAB_MIN") ? atoll(getenv("ADSEIS_AC_TB_SLAB_MIN")) : (3LL << 20);
    if (rows_min < 16 || (!(e && e[0] == '1') && cells < min_cells)) return;
  }
  const int fi0 = P->box_i0 + 2, fi1 = P->box_i1 - 2, fj0 = P->box_j0 + 2, fj1 = P->box_j1 - 2;
  const int mr0 = std::max(P->own0 + 2 * has_lo, fi0 - g.goff), mr1 = std::min(P->own1 - 2 * has_hi, fi1 + 1 - g.goff);
  const int mc0 = round_up(std::max(fj0, 1), 16);
  const int mc_end = mc0 + 2 * ((fj1 + 1 - mc0) / 2);
  if (mr1 - mr0 < 8 || mc_end - mc0 < 32) return;
  // Auto: a pair costs three launches (frame, box pair, frame) instead of two.  That pays once the box kernels are long
  // next to the two frame-only launches (~5-10 us each): measured on B200, 2000 x 1000 (C3) is 12 % SLOWER with pairs
  // (16.2 vs 14.5 us per forward step), 4096^2 is 27 % faster (profiles/r02_temporal_blocking.md).
  if (!slab && !(e && e[0] == '1') && (i64)(mr1 - mr0) * (mc_end - mc0) < (6LL << 20)) return;
  AcTiling& t = P->t2;
  memset(&t, 0, sizeof(t));
  t.mr0 = mr0; t.mr1 = mr1; t.mc0 = mc0; t.mc_end = mc_end;
  t.fcpt = AC_FRAME_CPT; t.fthr = AC_THREADS;
  const int nwarpcols = (mc_end - mc0 + AC_WCOLS - 1) / AC_WCOLS;
  t.nct = (nwarpcols + AC_WARPS - 1) / AC_WARPS;
  const int rows = mr1 - mr0;
  {
    // whole waves of 2 CTAs per SM, up to 3 waves while a CTA keeps >= 16 rows (two of its rows are rim overhead)
    const int slots = 2 * P->ctx->sm_count;
    const int per_wave = std::max(1, slots / t.nct);
    const int waves = std::min(3, std::max(1, rows / (16 * per_wave)));
    const int want_tr = per_wave * waves;
    int rb = (rows + want_tr - 1) / want_tr;
    rb = std::min(256, std::max(4, rb));
    if (getenv("ADSEIS_AC_RB2")) rb = std::max(2, atoi(getenv("ADSEIS_AC_RB2")));
    t.rb = rb;
  }
  t.ntr = (rows + t.rb - 1) / t.rb;
  t.nmarch = t.nct * t.ntr;
  P->nblocks2 = t.nmarch;
  // frame-only tilings: [0] everything outside [mr0, mr1) x [mc0, mc_end); [1] the same plus the rim ring of that box
  for (int kind = 0; kind < 2; kind++) {
    adseis_acoustic_plan::FrameKind& F = P->fk[kind];
    AcTiling& f = F.t;
    memset(&f, 0, sizeof(f));
    f.mr0 = f.mr1 = P->own0;  // no marched rows
    f.rb = 1;
    f.fcpt = getenv("ADSEIS_AC_FCPT2") ? std::max(1, atoi(getenv("ADSEIS_AC_FCPT2"))) : 1;  // alone in its launch: all parallelism
    f.fthr = AC_FO_THREADS;
    const int fcells2 = f.fthr * f.fcpt;
    auto add_rect = [&](int r0, int r1, int c0, int c1) {
      if (r1 <= r0 || c1 <= c0) return;
      const int k = f.nrect++;
      f.rr0[k] = r0; f.rr1[k] = r1; f.rc0[k] = c0; f.rc1[k] = c1;
      const i64 cells = (i64)(r1 - r0) * (c1 - c0);
      f.rblk[k + 1] = f.rblk[k] + (int)((cells + fcells2 - 1) / fcells2);
    };
    f.rblk[0] = 0;
    const int w = kind;   // rim width taken from the box
    add_rect(P->own0, mr0 + w, 0, g.W);
    add_rect(mr1 - w, P->own1, 0, g.W);
    add_rect(mr0 + w, mr1 - w, 0, mc0 + w);
    add_rect(mr0 + w, mr1 - w, mc_end - w, g.W);
    for (int k = f.nrect; k < 4; k++) { f.rblk[k + 1] = f.rblk[f.nrect]; f.rr0[k] = f.rr1[k] = f.rc0[k] = 0; f.rc1[k] = 1; }
    if (kind == 1) { f.bx_r0 = mr0; f.bx_r1 = mr1; f.bx_c0 = mc0; f.bx_c1 = mc_end; }
    F.nblocks = f.rblk[f.nrect];
    std::vector<int> order;
    build_launch_order(P, f, F.nblocks, &order, &F.n_edge_lo, &F.n_edge_hi);
    cudaFree(F.perm); F.perm = nullptr;
    if (slab && dev_upload(&F.perm, order, P->ctx->stream) != ADSEIS_OK) return;
  }
  P->tb = true;
  {
    const char* eo = getenv("ADSEIS_AC_TB_OVERLAP");
    P->tb_overlap = !(eo && eo[0] == '0');
  }
  const char* ea = getenv("ADSEIS_AC_TB_ADJ");
  P->tb_adj = P->p.PropagatorKernel != 0 && !(ea && ea[0] == '0');
  P->nub = P->tb_adj ? 4 : 3;
}

// Sources / receivers of a plan: keep the points whose padded row is owned (MPIAcoustic.jl:71-78, 98-104), group them
// by owner CTA (per-CTA CSR lists, see PointSet in common.cuh) and size the per-point buffers.  Called by plan_create
// and by adseis_acoustic_plan_set_points (one plan per GPU serves all shots of a multi-shot gradient).
// Whole-sweep (persistent) path: every owned cell is a general cell of ONE rectangle, split evenly over CTAs that are all
// resident at once (cooperative launch).  Built for small single-GPU grids; used when the tape is resident (one segment).
static void build_persist_tiling(adseis_acoustic_plan* P) {
  P->persist = false;
  const char* e = getenv("ADSEIS_AC_PERSIST");
  if (P->slab.nranks > 1 || (e && e[0] == '0')) return;
  const AcGeom& g = P->g;
  const i64 cells = (i64)(P->own1 - P->own0) * g.W;
  // measured on B200 (profiles/r02_small_grids.md): one launch per step costs ~8 / 10 us (forward / adjoint) whatever the
  // grid; the resident kernel costs a grid barrier plus one general cell per thread and step
  // 401 x 133: 16.9 -> 8.0 us per step pair (scheme 1), 25.6 -> 10.7 (scheme 0); break-even near 250 k cells (scheme 1) and
  // 150 k (scheme 0, whose frame cells re-evaluate their neighbours)
  if (!(e && e[0] == '1') && cells > ((P->p.PropagatorKernel == 0) ? (1LL << 17) : (1LL << 18))) return;
  int dev_coop = 0;
  if (cudaDeviceGetAttribute(&dev_coop, cudaDevAttrCooperativeLaunch, P->ctx->device) != cudaSuccess || !dev_coop) return;
  int per_sm = 0, per_sm2 = 0;
  const bool k0 = P->p.PropagatorKernel == 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k0 ? ac_fwd_persist_kernel<0> : ac_fwd_persist_kernel<1>, AC_PS_THREADS, 0) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k0 ? ac_adj_persist_kernel<0> : ac_adj_persist_kernel<1>, AC_PS_THREADS, 0) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  const i64 max_ctas = (i64)std::min(per_sm, per_sm2) * P->ctx->sm_count;
  if (max_ctas < 1) return;
  AcTiling& t = P->tp;
  memset(&t, 0, sizeof(t));
  t.mr0 = t.mr1 = P->own0;
  t.rb = 1;
  t.fthr = AC_PS_THREADS;
  t.fcpt = (int)std::max<i64>(1, (cells + AC_PS_THREADS * max_ctas - 1) / (AC_PS_THREADS * max_ctas));
  if (getenv("ADSEIS_AC_PERSIST_CPT")) t.fcpt = std::max(t.fcpt, atoi(getenv("ADSEIS_AC_PERSIST_CPT")));
  const i64 per = (i64)t.fthr * t.fcpt;
  t.nrect = 1;
  t.rr0[0] = P->own0; t.rr1[0] = P->own1; t.rc0[0] = 0; t.rc1[0] = g.W;
  t.rblk[0] = 0; t.rblk[1] = (int)((cells + per - 1) / per);
  for (int k = 1; k < 4; k++) { t.rblk[k + 1] = t.rblk[1]; t.rr0[k] = t.rr1[k] = t.rc0[k] = 0; t.rc1[k] = 1; }
  P->nblocksP = t.rblk[1];
  P->persist = P->nblocksP >= 1 && P->nblocksP <= max_ctas;
}

static int plan_build_points(adseis_acoustic_plan* P, int64_t nsrc, const int64_t* srci, const int64_t* srcj,
                             int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj) {
  const adseis_acoustic_params* p = &P->p;
  const AcGeom& g = P->g;
  const adseis_slab& sl = P->slab;
  cudaStream_t st = P->ctx->stream;
  const int H = g.H, W = g.W;
  std::vector<double> sx(H), ty(W);
  TRY(adseis_acoustic_pml_profiles(p, sx.data(), ty.data()));
  P->srcp = AcPoints{}; P->rcvp = AcPoints{};
  // per-point buffers are kept (same device addresses) when the counts do not change: the usual multi-shot case
  const bool same_counts = (P->rcvv != nullptr) && P->nsrc == nsrc && P->nrcv == nrcv;
  if (!same_counts) {
    cudaFree(P->rcv_owned); P->rcv_owned = nullptr;
    cudaFree(P->rcvv); P->rcvv = nullptr;
    cudaFree(P->obs); P->obs = nullptr;
    cudaFree(P->res); P->res = nullptr;
    cudaFree(P->srcv); P->srcv = nullptr; P->srcv_rows = 0;
    cudaFree(P->gradsrcv); P->gradsrcv = nullptr;
    P->points_gen++;
  }
  P->have_srcv = P->have_obs = P->have_fwd = P->have_grad = false;
  P->k0_corr = false;
  P->nsrc = nsrc; P->nrcv = nrcv;
  const int ioff = p->mpi_convention ? 0 : -1;  // 1-based padded -> 0-based padded ; 1-based unpadded -> padded
  auto owner_cta = [&](int li, int j) -> int {
    const AcTiling& t = P->t;
    if (li >= t.mr0 && li < t.mr1 && j >= t.mc0 && j < t.mc_end)
      return ac_row_tile_of(t, li) * t.nct + (j - t.mc0) / AC_TILE_COLS;
    for (int k = 0; k < t.nrect; k++)
      if (li >= t.rr0[k] && li < t.rr1[k] && j >= t.rc0[k] && j < t.rc1[k])
        return t.nmarch + t.rblk[k] + (int)(((i64)(li - t.rr0[k]) * (t.rc1[k] - t.rc0[k]) + (j - t.rc0[k])) / AC_FRAME_CELLS);
    return -1;
  };
  auto build = [&](i64 n, const int64_t* pi, const int64_t* pj, PointSetStorage* dst, std::vector<unsigned char>* owned,
                   const char* what) -> int {
    std::vector<int> own, cells, gid, none;
    if (owned) owned->assign((size_t)n, 0);
    for (i64 k = 0; k < n; k++) {
      i64 gi = pi[k] + ioff, gj = pj[k] + ioff;
      REQUIRE(gi >= 0 && gi < H && gj >= 0 && gj < W, "acoustic plan: %s %lld at (%lld,%lld) is outside the grid",
              what, (long long)k, (long long)pi[k], (long long)pj[k]);
      if (gi >= sl.row0 && gi < sl.row1) {
        const int li = (int)(gi - g.goff);
        const int o = owner_cta(li, (int)gj);
        REQUIRE(o >= 0, "acoustic plan: internal error: cell (%d,%d) has no owner CTA", li, (int)gj);
        own.push_back(o); cells.push_back(li * g.ld + (int)gj); gid.push_back((int)k);
        if (owned) (*owned)[k] = 1;
      }
    }
    PointSetHost h;
    build_point_set(own, cells, gid, none, P->nblocks, &h);
    return upload_point_set(h, dst, st);
  };
  if (p->PropagatorKernel == 0)
    for (i64 k = 0; k < nsrc && !P->k0_corr; k++) {
      const i64 gi = srci[k] + ioff, gj = srcj[k] + ioff;
      auto coef = [&](i64 i, i64 j) { return i >= 1 && i <= H - 2 && j >= 1 && j <= W - 2 && sx[i] != ty[j]; };
      P->k0_corr = coef(gi - 1, gj) || coef(gi + 1, gj) || coef(gi, gj - 1) || coef(gi, gj + 1);
    }
  std::vector<unsigned char> owned;
  TRY(build(nsrc, srci, srcj, &P->src, nullptr, "source"));
  if (P->unfused) {   // sources on my rows and on the neighbours' rows next to them, grouped by cell (no owner CTA needed)
    std::vector<int> own, cells, gid, none;
    for (i64 k = 0; k < nsrc; k++) {
      const i64 gi = srci[k] + ioff, gj = srcj[k] + ioff;
      if (gi >= sl.row0 - 1 && gi < sl.row1 + 1) {
        own.push_back(0); cells.push_back((int)(gi - g.goff) * g.ld + (int)gj); gid.push_back((int)k);
      }
    }
    PointSetHost h;
    build_point_set(own, cells, gid, none, 1, &h);
    TRY(upload_point_set(h, &P->srcK, st));
  }
  TRY(build(nrcv, rcvi, rcvj, &P->rcv, &owned, "receiver"));
  if (P->persist) {   // the same points grouped by the CTAs of the whole-sweep tiling
    auto buildp = [&](i64 n, const int64_t* pi, const int64_t* pj, PointSetStorage* dst) -> int {
      std::vector<int> own, cells, gid, none;
      const i64 per = (i64)P->tp.fthr * P->tp.fcpt;
      for (i64 k = 0; k < n; k++) {
        const i64 gi = pi[k] + ioff, gj = pj[k] + ioff;
        const int li = (int)(gi - g.goff);
        own.push_back((int)(((i64)(li - P->own0) * g.W + gj) / per));
        cells.push_back(li * g.ld + (int)gj); gid.push_back((int)k);
      }
      PointSetHost h;
      build_point_set(own, cells, gid, none, P->nblocksP, &h);
      return upload_point_set(h, dst, st);
    };
    TRY(buildp(nsrc, srci, srcj, &P->srcP));
    TRY(buildp(nrcv, rcvi, rcvj, &P->rcvP));
  }
  if (P->src.nu > 0) P->srcp = AcPoints{P->src.blk, P->src.cell, P->src.start, P->src.perm};
  if (P->rcv.nu > 0) P->rcvp = AcPoints{P->rcv.blk, P->rcv.cell, P->rcv.start, P->rcv.perm};
  if (!P->rcv_owned) TRY(dev_alloc(&P->rcv_owned, owned.size()));
  if (!owned.empty()) CUDA_TRY(cudaMemcpyAsync(P->rcv_owned, owned.data(), owned.size(), cudaMemcpyHostToDevice, st));
  if (!P->rcvv) TRY(dev_alloc(&P->rcvv, (size_t)((p->NSTEP + 1) * nrcv)));
  CUDA_TRY(cudaMemsetAsync(P->rcvv, 0, std::max<size_t>((size_t)((p->NSTEP + 1) * nrcv), 1) * 8, st));
  if (P->G && !P->gradsrcv) TRY(dev_alloc_zero(&P->gradsrcv, (size_t)(p->NSTEP * nsrc), st));  // adjoint state already exists
  // two-step path: the same points grouped by owner under the frame-only tiling (frame points), under the box tiling
  // (box points), and -- sources only -- under every box tile whose one-cell rim contains them (injection into the
  // tile's private copy of the intermediate time level)
  P->srcMp = P->rcvMp = P->srcHp = P->rcvHp = AcPoints{};
  for (int k = 0; k < 2; k++) P->fk[k].srcp = P->fk[k].rcvp = P->fk[k].srcNp = P->fk[k].rcvNp = P->fk[k].srcRp = P->fk[k].rcvRp = AcPoints{};
  if (P->tb) {
    const AcTiling& t2 = P->t2;
    auto in_box = [&](int li, int j, int w) {   // strictly inside the box shrunk by w
      return li >= t2.mr0 + w && li < t2.mr1 - w && j >= t2.mc0 + w && j < t2.mc_end - w;
    };
    auto owner_f = [&](const AcTiling& tf, int li, int j) -> int {
      for (int k = 0; k < tf.nrect; k++)
        if (li >= tf.rr0[k] && li < tf.rr1[k] && j >= tf.rc0[k] && j < tf.rc1[k])
          return tf.rblk[k] + (int)(((i64)(li - tf.rr0[k]) * (tf.rc1[k] - tf.rc0[k]) + (j - tf.rc0[k])) / (tf.fthr * tf.fcpt));
      return -1;
    };
    auto build2 = [&](i64 n, const int64_t* pi, const int64_t* pj, int mode, int nblk, PointSetStorage* dst) -> int {
      std::vector<int> own, cells, gid, none;
      for (i64 k = 0; k < n; k++) {
        const i64 gi = pi[k] + ioff, gj = pj[k] + ioff;
        const int li = (int)(gi - g.goff), j = (int)gj;
        if (mode == 0 || mode >= 3) {   // frame; 3: wide frame = frame + rim ring of the box, 4: its frame part, 5: its rim part
          const int w = mode >= 3 ? 1 : 0;
          if (in_box(li, j, w)) continue;
          if (mode == 4 && in_box(li, j, 0)) continue;
          if (mode == 5 && !in_box(li, j, 0)) continue;
          const int o = owner_f(P->fk[w].t, li, j);
          REQUIRE(o >= 0, "acoustic plan: internal error: frame cell (%d,%d) has no owner CTA", li, j);
          own.push_back(o); cells.push_back(li * g.ld + j); gid.push_back((int)k);
        } else if (mode == 1) {   // box
          if (!in_box(li, j, 0)) continue;
          own.push_back(ac_row_tile_of(t2, li) * t2.nct + (j - t2.mc0) / AC_TILE_COLS);
          cells.push_back(li * g.ld + j); gid.push_back((int)k);
        } else {                  // every box tile whose (tile + rim) holds the cell
          if (li < t2.mr0 - 1 || li > t2.mr1 || j < t2.mc0 - 1 || j > t2.mc_end) continue;
          for (int tr = 0; tr < t2.ntr; tr++) {
            int r0, r1;
            ac_row_tile(t2, tr, &r0, &r1);
            if (li < r0 - 1 || li > r1) continue;
            for (int ct = 0; ct < t2.nct; ct++) {
              const int c0 = t2.mc0 + ct * AC_TILE_COLS, c1 = std::min(c0 + AC_TILE_COLS, t2.mc_end);
              if (j < c0 - 1 || j > c1) continue;
              own.push_back(tr * t2.nct + ct); cells.push_back(li * g.ld + j); gid.push_back((int)k);
            }
          }
        }
      }
      PointSetHost h;
      build_point_set(own, cells, gid, none, nblk, &h);
      return upload_point_set(h, dst, st);
    };
    for (int k = 0; k < 2; k++) {
      TRY(build2(nsrc, srci, srcj, k ? 3 : 0, P->fk[k].nblocks, &P->fk[k].src));
      TRY(build2(nrcv, rcvi, rcvj, k ? 3 : 0, P->fk[k].nblocks, &P->fk[k].rcv));
    }
    TRY(build2(nsrc, srci, srcj, 4, P->fk[1].nblocks, &P->fk[1].srcN));
    TRY(build2(nrcv, rcvi, rcvj, 4, P->fk[1].nblocks, &P->fk[1].rcvN));
    TRY(build2(nsrc, srci, srcj, 5, P->fk[1].nblocks, &P->fk[1].srcR));
    TRY(build2(nrcv, rcvi, rcvj, 5, P->fk[1].nblocks, &P->fk[1].rcvR));
    TRY(build2(nsrc, srci, srcj, 1, P->nblocks2, &P->srcM));
    TRY(build2(nrcv, rcvi, rcvj, 1, P->nblocks2, &P->rcvM));
    TRY(build2(nsrc, srci, srcj, 2, P->nblocks2, &P->srcH));
    TRY(build2(nrcv, rcvi, rcvj, 2, P->nblocks2, &P->rcvH));
    auto view = [](const PointSetStorage& q) { return q.nu > 0 ? AcPoints{q.blk, q.cell, q.start, q.perm} : AcPoints{}; };
    for (int k = 0; k < 2; k++) { P->fk[k].srcp = view(P->fk[k].src); P->fk[k].rcvp = view(P->fk[k].rcv); }
    P->fk[1].srcNp = view(P->fk[1].srcN); P->fk[1].rcvNp = view(P->fk[1].rcvN);
    P->fk[1].srcRp = view(P->fk[1].srcR); P->fk[1].rcvRp = view(P->fk[1].rcvR);
    P->srcMp = view(P->srcM); P->rcvMp = view(P->rcvM);
    P->srcHp = view(P->srcH); P->rcvHp = view(P->rcvH);
  }
  return ADSEIS_OK;
}

// Replace the sources and receivers of a plan (a new shot on the same grid): the device state -- history window,
// checkpoints, adjoint planes, model -- is kept, so a multi-shot gradient (compute_loss_and_grads_GPU,
// src/Utils.jl:300-332) runs on ONE plan per GPU instead of one 100-GB history allocation per shot.
// set_srcv / set_obs must be called again afterwards.
ADSEIS_API int adseis_acoustic_plan_set_points(adseis_acoustic_plan* P, int64_t nsrc, const int64_t* srci,
                                               const int64_t* srcj, int64_t nrcv, const int64_t* rcvi,
                                               const int64_t* rcvj) {
  REQUIRE(P, "acoustic_plan_set_points: null plan");
  REQUIRE(nsrc >= 0 && nrcv >= 0 && (nsrc == 0 || (srci && srcj)) && (nrcv == 0 || (rcvi && rcvj)),
          "acoustic_plan_set_points: bad source/receiver arrays");
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  CUDA_TRY(cudaStreamSynchronize(P->ctx->stream));
  return plan_build_points(P, nsrc, srci, srcj, nrcv, rcvi, rcvj);
}

ADSEIS_API int adseis_acoustic_plan_create(adseis_ctx* ctx, const adseis_acoustic_params* p, const adseis_slab* slab,
                                           int64_t nsrc, const int64_t* srci, const int64_t* srcj, int64_t nrcv,
                                           const int64_t* rcvi, const int64_t* rcvj, size_t hist_bytes_budget,
                                           adseis_acoustic_plan** out) {
  REQUIRE(ctx && out, "acoustic_plan_create: null ctx/out");
  *out = nullptr;
  TRY(validate_params(p));
  REQUIRE(nsrc >= 0 && nrcv >= 0 && (nsrc == 0 || (srci && srcj)) && (nrcv == 0 || (rcvi && rcvj)),
          "acoustic_plan_create: bad source/receiver arrays");
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (AC_FWD_SMEM > 0) {
    CUDA_TRY(cudaFuncSetAttribute(ac_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AC_FWD_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(ac_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AC_FWD_SMEM));
  }
  CUDA_TRY(cudaFuncSetAttribute(ac_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AC_FWD2_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(ac_adj2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AC_ADJ2_SMEM));
  if (AC_ADJ_SMEM > 0) {
    CUDA_TRY(cudaFuncSetAttribute(ac_adj_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AC_ADJ_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(ac_adj_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AC_ADJ_SMEM));
  }
  cudaStream_t st = ctx->stream;
  adseis_acoustic_plan* P = new adseis_acoustic_plan();
  P->ctx = ctx;
  ctx->plans++;
  P->p = *p;
  if (P->p.PropagatorKernel == 2) P->p.PropagatorKernel = 1;  // Core.jl:504-525 == the custom op without the op
  p = &P->p;
  const int H = (int)p->NX + 2, W = (int)p->NY + 2;
  if (slab) P->slab = *slab; else { P->slab.rank = 0; P->slab.nranks = 1; P->slab.row0 = 0; P->slab.row1 = H; }
  const adseis_slab& sl = P->slab;
  if (!(sl.nranks >= 1 && sl.rank >= 0 && sl.rank < sl.nranks && sl.row0 >= 0 && sl.row1 <= H && sl.row0 < sl.row1 &&
        (sl.rank > 0 || sl.row0 == 0) && (sl.rank < sl.nranks - 1 || sl.row1 == H))) {
    delete P;
    ctx->plans--;
    adseis_set_error("acoustic_plan_create: inconsistent slab {rank %d/%d rows [%lld,%lld)}", sl.rank, sl.nranks,
                     (long long)sl.row0, (long long)sl.row1);
    return ADSEIS_EINVAL;
  }
  // PropagatorKernel = 0 (phi', psi' from the NEW wavefield, MPIAcoustic.jl:212-246): an edge-row frame cell re-evaluates
  // u' of the neighbour's edge row, which reads one row further -- two halo rows, exchanged by an explicit kernel after
  // every step launch (the reference exchanges u' inside its one_step the same way, with mpi_halo_exchange)
  P->hd = (p->PropagatorKernel == 0 && sl.nranks > 1) ? 2 : 1;
  P->unfused = P->hd == 2;
  if (sl.nranks > 1 && sl.row1 - sl.row0 < P->hd) {
    const int need = P->hd;
    delete P;
    ctx->plans--;
    adseis_set_error("acoustic_plan_create: a slab needs at least %d rows", need);
    return ADSEIS_EINVAL;
  }
  const int halo_lo = sl.rank > 0 ? P->hd : 0, halo_hi = sl.rank < sl.nranks - 1 ? P->hd : 0;
  AcGeom& g = P->g;
  g.H = H; g.W = W;
  g.goff = (int)sl.row0 - halo_lo;
  g.Hl = (int)(sl.row1 - sl.row0) + halo_lo + halo_hi;
  g.ld = round_up(W, 16);
  g.plane = (i64)g.Hl * g.ld;
  P->own0 = halo_lo;
  P->own1 = halo_lo + (int)(sl.row1 - sl.row0);
  const double dt = p->DELTAT, hx = p->DELTAX, hy = p->DELTAY;
  g.dt = dt; g.hx = hx; g.hy = hy;
  g.kx2 = 2 * dt * dt / hx / hx; g.ky2 = 2 * dt * dt / hy / hy;
  g.rx = dt / hx; g.ry = dt / hy;
  g.px = dt * dt / (2.0 * hx); g.py = dt * dt / (2.0 * hy);
  g.dt2 = dt * dt;
  g.rhx = 1.0 / hx; g.rhy = 1.0 / hy;
  P->model_elems = p->mpi_convention ? p->NX * p->NY : (i64)H * W;

#define PTRY(expr)                                   \
  do {                                               \
    int _r = (expr);                                 \
    if (_r != ADSEIS_OK) {                           \
      adseis_acoustic_plan_destroy(P);               \
      return _r;                                     \
    }                                                \
  } while (0)
#define PCUDA(expr)                                                                              \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      adseis_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));   \
      adseis_acoustic_plan_destroy(P);                                                          \
      return (_e == cudaErrorMemoryAllocation) ? ADSEIS_ENOMEM : ADSEIS_ECUDA;                  \
    }                                                                                           \
  } while (0)
  // PML profiles and the PML-free box
  std::vector<double> sx(H), ty(W);
  int rc = adseis_acoustic_pml_profiles(p, sx.data(), ty.data());
  if (rc) { adseis_acoustic_plan_destroy(P); return rc; }
  auto box = [](const std::vector<double>& v, int n, int* a, int* b) {  // maximal zero run inside 1..n
    int lo = 1;
    while (lo <= n && v[lo] != 0.0) lo++;
    int hi = n;
    while (hi >= 1 && v[hi] != 0.0) hi--;
    for (int k = lo; k <= hi; k++)
      if (v[k] != 0.0) { lo = 1; hi = 0; break; }  // not a single box: no fast region
    *a = lo; *b = hi;
  };
  int ia, ib, ja, jb;
  box(sx, (int)p->NX, &ia, &ib);
  box(ty, (int)p->NY, &ja, &jb);
  P->box_i0 = ia; P->box_i1 = ib; P->box_j0 = ja; P->box_j1 = jb;
  {
    // fast region = box shrunk by one cell: every stencil neighbour is PML-free, interior, and has phi=psi=0
    // (PropagatorKernel=0: by two cells -- the adjoint of a neighbour's u' collects phibar/psibar of ITS neighbours)
    const int shrink = p->PropagatorKernel == 0 ? 2 : 1;
    const int fi0 = ia + shrink, fi1 = ib - shrink, fj0 = ja + shrink, fj1 = jb - shrink;
    AcTiling& t = P->t;
    memset(&t, 0, sizeof(t));
    t.fcpt = AC_FRAME_CPT; t.fthr = AC_THREADS;
    // marched local rows: owned rows whose global index lies in [fi0, fi1]
    int mr0 = std::max(P->own0, fi0 - g.goff), mr1 = std::min(P->own1, fi1 + 1 - g.goff);
    int mc0 = round_up(std::max(fj0, 1), 16);
    int mc_end = mc0 + 2 * ((fj1 + 1 - mc0) / 2);
    if (fj1 + 1 - mc0 < 2 || mr1 - mr0 < 1) { mr0 = mr1 = P->own0; mc0 = mc_end = 0; }
    t.mr0 = mr0; t.mr1 = mr1; t.mc0 = mc0; t.mc_end = mc_end;
    P->fast_rows = mr1 - mr0;
    const int nwarpcols = (mc_end - mc0 + AC_WCOLS - 1) / AC_WCOLS;
    t.nct = (nwarpcols + AC_WARPS - 1) / AC_WARPS;
    // thin edge tiles (slab plans): one row next to each neighbour, when that row is a marched row
    if (halo_lo && mr0 == P->own0 && mr1 - mr0 >= 8) t.elo = 1;
    if (halo_hi && mr1 == P->own1 && mr1 - mr0 >= 8) t.ehi = 1;
    // rows per marching CTA: the TMA-ring kernels run 2 CTAs per SM, and all CTAs of a launch should be resident at
    // once in whole waves (a CTA that has to wait for a slot ends the launch late): marching + frame CTAs fill
    // 1..3 waves of 2 * SMs, up to 3 waves for load balance as long as a CTA keeps >= 12 rows to amortise its
    // pipeline prologue (small slabs of a domain decomposition get exactly one wave)
    int nframe_est = 0;
    {
      const i64 fc[4] = {(i64)(mr0 - P->own0) * g.W, (i64)(P->own1 - mr1) * g.W, (i64)(mr1 - mr0) * mc0,
                         (i64)(mr1 - mr0) * (g.W - mc_end)};
      for (int k = 0; k < 4; k++) nframe_est += (int)((fc[k] + AC_FRAME_CELLS - 1) / AC_FRAME_CELLS);
    }
    int rb = 32;
    const int nedge = (t.elo ? 1 : 0) + (t.ehi ? 1 : 0);
    const int mid = mr1 - mr0 - t.elo - t.ehi;
    if (t.nct > 0 && mid > 0) {
      const int slots = 2 * ctx->sm_count;
      const int per_wave = std::max(1, (slots + t.nct - 1) / t.nct);  // row tiles of one wave
      const int waves = std::min(3, std::max(1, mid / (12 * per_wave)));
      int want_tr = per_wave * waves;
      if (waves == 1) {  // one wave: every CTA of the launch (marching, thin edge and frame CTAs) gets a slot at once
        want_tr = (slots - nframe_est) / t.nct - nedge;
        if (want_tr < 1) want_tr = std::max(1, per_wave - nedge);
      }
      rb = (mid + want_tr - 1) / want_tr;
      rb = std::min(128, std::max(2, rb));
      if (getenv("ADSEIS_AC_RB")) rb = std::max(2, atoi(getenv("ADSEIS_AC_RB")));  // tuning experiments
    }
    t.rb = rb;
    t.ntr = (mr1 > mr0) ? nedge + (mid + rb - 1) / rb : 0;
    t.nmarch = t.nct * t.ntr;
    // frame rectangles: rows above / below the marched rows (all columns), columns left / right of the marched ones
    auto add_rect = [&](int r0, int r1, int c0, int c1) {
      if (r1 <= r0 || c1 <= c0) return;
      const int k = t.nrect++;
      t.rr0[k] = r0; t.rr1[k] = r1; t.rc0[k] = c0; t.rc1[k] = c1;
      const i64 cells = (i64)(r1 - r0) * (c1 - c0);
      t.rblk[k + 1] = t.rblk[k] + (int)((cells + AC_FRAME_CELLS - 1) / AC_FRAME_CELLS);
    };
    t.rblk[0] = 0;
    // (pitch-padding columns j >= W are never read by a cell that is stored, so nobody needs to write them)
    add_rect(P->own0, mr0, 0, g.W);
    add_rect(mr1, P->own1, 0, g.W);
    add_rect(mr0, mr1, 0, mc0);
    add_rect(mr0, mr1, mc_end, g.W);
    for (int k = t.nrect; k < 4; k++) { t.rblk[k + 1] = t.rblk[t.nrect]; t.rr0[k] = t.rr1[k] = t.rc0[k] = 0; t.rc1[k] = 1; }
    P->nblocks = t.nmarch + t.rblk[t.nrect];
  }
  {
    std::vector<int> order;
    build_launch_order(P, P->t, P->nblocks, &order, &P->n_edge_lo, &P->n_edge_hi);
    PTRY(dev_upload(&P->perm, order, st));
  }

  build_tb_tilings(P);
  build_persist_tiling(P);
  PTRY(dev_upload(&P->sigx, sx, st));
  PTRY(dev_upload(&P->tauy, ty, st));
  PTRY(dev_alloc_zero(&P->c2, (size_t)g.plane, st));
  PTRY(dev_alloc_zero(&P->cvel, (size_t)g.plane, st));
  if (sl.nranks == 1)
    for (int k = 0; k < 2; k++) {
      PTRY(dev_alloc_zero(&P->phi[k], (size_t)g.plane, st));
      PTRY(dev_alloc_zero(&P->psi[k], (size_t)g.plane, st));
    }

  PTRY(plan_build_points(P, nsrc, srci, srcj, nrcv, rcvi, rcvj));
  PTRY(dev_alloc_zero(&P->loss, 1 + RL_BLOCKS, st));

  // adjoint state is allocated lazily (first gradient); history sizing now
  size_t free_b = 0, total_b = 0;
  PCUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t plane_bytes = (size_t)g.plane * sizeof(double);
  // reserve: adjoint state (3 ubar + 2 phib + 2 psib + G + gradc) + slack
  const size_t reserve = 10 * plane_bytes + (size_t)(2 * (p->NSTEP + 1) * nrcv + 2 * p->NSTEP * nsrc) * 8 + (512u << 20);
  size_t budget = hist_bytes_budget;
  if (budget == 0) budget = free_b > reserve ? free_b - reserve : 0;
  if (budget > free_b) budget = free_b;
  PTRY(plan_segments(P, budget, hist_bytes_budget == 0));
  if (sl.nranks == 1) {
    cudaError_t e = cudaMalloc((void**)&P->hist, (size_t)P->win * plane_bytes);
    if (e != cudaSuccess) {
      adseis_set_error("acoustic_plan_create: cannot allocate %lld history snapshots of %zu bytes: %s",
                       (long long)P->win, plane_bytes, cudaGetErrorString(e));
      adseis_acoustic_plan_destroy(P);
      return ADSEIS_ENOMEM;
    }
  } else {
    // one arena: [descriptor | flags | history window | phi,psi x2 | ubar x3 | phibar,psibar x2]
    if (P->win < 6) {
      adseis_set_error("acoustic_plan_create: slab plans need a history window of >= 6 snapshots (got %lld)",
                       (long long)P->win);
      adseis_acoustic_plan_destroy(P);
      return ADSEIS_ENOMEM;
    }
    AcDesc& d = P->desc;
    d.magic = AC_DESC_MAGIC; d.Hl = g.Hl; d.ld = g.ld; d.plane = g.plane; d.win = P->win; d.own0 = P->own0; d.own1 = P->own1;
    d.n_edge_lo = P->n_edge_lo; d.n_edge_hi = P->n_edge_hi;
    d.n_edge_lo_f = P->tb ? P->fk[0].n_edge_lo : 0; d.n_edge_hi_f = P->tb ? P->fk[0].n_edge_hi : 0; d.nub = P->nub;
    d.n_edge_lo_w = P->tb ? P->fk[1].n_edge_lo : 0; d.n_edge_hi_w = P->tb ? P->fk[1].n_edge_hi : 0;
    long long off = 512;  // bytes; descriptor lives in [0,512)
    d.off_flags = off; off += 512;
    auto take = [&](long long nplanes) { long long o = off; off += nplanes * (long long)plane_bytes; return o; };
    d.off_hist = take(P->win);
    for (int k = 0; k < 2; k++) d.off_phi[k] = take(1);
    for (int k = 0; k < 2; k++) d.off_psi[k] = take(1);
    for (int k = 0; k < 4; k++) d.off_ub[k] = (k < P->nub) ? take(1) : 0;
    for (int k = 0; k < 2; k++) d.off_phib[k] = take(1);
    for (int k = 0; k < 2; k++) d.off_psib[k] = take(1);
    {
      const char* ell = getenv("ADSEIS_AC_LL");
      P->ll = !(ell && ell[0] == '0') && !P->unfused;
      d.off_ll = 0;
      if (P->ll) { d.off_ll = off; off += 8LL * g.ld * 16; off = (off + 511) / 512 * 512; }
    }
    P->arena_bytes = (size_t)off;
    cudaError_t e = cudaMalloc((void**)&P->arena, P->arena_bytes);
    if (e != cudaSuccess) {
      adseis_set_error("acoustic_plan_create: cannot allocate the %zu-byte slab arena: %s", P->arena_bytes,
                       cudaGetErrorString(e));
      adseis_acoustic_plan_destroy(P);
      return ADSEIS_ENOMEM;
    }
    PCUDA(cudaMemsetAsync(P->arena, 0, P->arena_bytes, st));
    PCUDA(cudaMemcpyAsync(P->arena, &d, sizeof(d), cudaMemcpyHostToDevice, st));
    char* base = (char*)P->arena;
    P->hist = (double*)(base + d.off_hist);
    for (int k = 0; k < 2; k++) { P->phi[k] = (double*)(base + d.off_phi[k]); P->psi[k] = (double*)(base + d.off_psi[k]); }
    for (int k = 0; k < P->nub; k++) P->ub[k] = (double*)(base + d.off_ub[k]);
    for (int k = 0; k < 2; k++) { P->phib[k] = (double*)(base + d.off_phib[k]); P->psib[k] = (double*)(base + d.off_psib[k]); }
    PCUDA(cudaStreamSynchronize(st));
  }
  for (size_t k = 1; k < P->seg_b.size(); k++) {
    double* c = nullptr;
    PTRY(dev_alloc(&c, (size_t)4 * g.plane));
    P->ckpt.push_back(c);
  }
  *out = P;
  return ADSEIS_OK;
}

// copy the caller's dense model into a pitched local array (rows of this slab, halo rows included)
static int load_model_plane(adseis_acoustic_plan* P, const double* src, double* dst) {
  const AcGeom& g = P->g;
  cudaStream_t st = P->ctx->stream;
  CUDA_TRY(cudaMemsetAsync(dst, 0, (size_t)g.plane * 8, st));
  if (!P->p.mpi_convention) {
    // local row li <-> global padded row goff+li ; all W columns
    int l0 = std::max(0, -g.goff), l1 = std::min(g.Hl, g.H - g.goff);
    CUDA_TRY(cudaMemcpy2DAsync(dst + (i64)l0 * g.ld, (size_t)g.ld * 8, src + (i64)(g.goff + l0) * g.W, (size_t)g.W * 8,
                               (size_t)g.W * 8, (size_t)(l1 - l0), cudaMemcpyDefault, st));
  } else {
    const int NY = g.W - 2;
    int l0 = std::max(0, 1 - g.goff), l1 = std::min(g.Hl, g.H - 1 - g.goff);  // global rows 1..NX
    if (l1 > l0)
      CUDA_TRY(cudaMemcpy2DAsync(dst + (i64)l0 * g.ld + 1, (size_t)g.ld * 8, src + (i64)(g.goff + l0 - 1) * NY,
                                 (size_t)NY * 8, (size_t)NY * 8, (size_t)(l1 - l0), cudaMemcpyDefault, st));
  }
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_set_model(adseis_acoustic_plan* P, const double* c, int on_device) {
  (void)on_device;
  REQUIRE(P && c, "acoustic_plan_set_model: null");
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  if (!P->p.mpi_convention) {
    TRY(load_model_plane(P, c, P->cvel));
    const i64 n = P->g.plane;
    k_square<<<(unsigned)((n + 255) / 256), 256, 0, P->ctx->stream>>>(P->c2, P->cvel, n);  // Core.jl:564
    P->ctx->launches++;
    CUDA_TRY(cudaGetLastError());
  } else {
    TRY(load_model_plane(P, c, P->c2));  // MPIAcoustic.jl:336: no squaring
  }
  P->have_model = true;
  P->have_fwd = P->have_grad = false;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_set_srcv(adseis_acoustic_plan* P, const double* srcv, int64_t rows, int on_device) {
  (void)on_device;
  REQUIRE(P && (srcv || P->nsrc == 0), "acoustic_plan_set_srcv: null");
  REQUIRE(rows >= P->p.NSTEP, "acoustic_plan_set_srcv: srcv has %lld rows, need >= NSTEP=%lld", (long long)rows,
          (long long)P->p.NSTEP);
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  if (P->srcv_rows != P->p.NSTEP || !P->srcv) {
    cudaFree(P->srcv);
    P->srcv = nullptr;
    TRY(dev_alloc(&P->srcv, (size_t)(P->p.NSTEP * P->nsrc)));
    P->srcv_rows = P->p.NSTEP;
  }
  if (P->nsrc > 0)
    CUDA_TRY(cudaMemcpyAsync(P->srcv, srcv, (size_t)(P->p.NSTEP * P->nsrc) * 8, cudaMemcpyDefault, P->ctx->stream));
  P->have_srcv = true;
  P->have_fwd = P->have_grad = false;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_set_obs(adseis_acoustic_plan* P, const double* obs, int on_device) {
  (void)on_device;
  REQUIRE(P && (obs || P->nrcv == 0), "acoustic_plan_set_obs: null");
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  const size_t n = (size_t)((P->p.NSTEP + 1) * P->nrcv);
  if (!P->obs) {
    TRY(dev_alloc(&P->obs, n));
    TRY(dev_alloc(&P->res, n));
  }
  if (n > 0) CUDA_TRY(cudaMemcpyAsync(P->obs, obs, n * 8, cudaMemcpyDefault, P->ctx->stream));
  P->have_obs = true;
  P->have_grad = false;
  return ADSEIS_OK;
}

// ---- forward steps s_first..s_last of segment k into the window (slot s at index s - base) -----------------
// Peer pointers and flag expectations of the next step launch (slab plans); arr_u/arr_p: the arrays it produces.
// `frame_only`: the launch uses the frame-only tiling of the two-step path (its own launch order and edge-CTA counts).
// The flag a rank waits on counts the signals of ALL fused launches its neighbour has issued so far; both ranks issue
// the same sequence of launches, so the expectation is accumulated on the host, launch by launch.
static AcFuse make_fuse(adseis_acoustic_plan* P, int arr_u, i64 idx_u, int arr_p, i64 idx_p, int fkind = 0,
                        bool recv = false, double* in_u = nullptr, double* in_p = nullptr) {
  AcFuse f;
  memset(&f, 0, sizeof(f));
  f.perm = fkind ? P->fk[fkind - 1].perm : P->perm;
  if (!P->arena || P->unfused) return f;
  f.own0 = P->own0; f.own_last = P->own1 - 1;
  f.has_lo = P->peer[0] != nullptr; f.has_hi = P->peer[1] != nullptr;
  P->sepoch++;
  if (P->ll) {
    // LL rows: region (from, array, parity); I am rank-1's "hi" and rank+1's "lo" neighbour
    auto row = [&](char* base, const AcDesc& d, int from, int arr, unsigned long long ep) {
      return (ulonglong2*)(base + d.off_ll + (long long)((from * 2 + arr) * 2 + (int)(ep & 1ULL)) * d.ld * 16);
    };
    const unsigned long long es = P->sepoch, er = P->sepoch - 1;
    f.ll = 1;
    f.ep_send = (unsigned)(es % 0xFFFFFFFEULL) + 1u;
    f.ep_recv = recv ? (unsigned)(er % 0xFFFFFFFEULL) + 1u : 0u;
    f.in_u = in_u; f.in_p = in_p;
    f.my_flags = (unsigned long long*)((char*)P->arena + P->desc.off_flags);
    if (f.has_lo) {
      f.tx_lo_u = row(P->peer[0], P->dpeer[0], 1, 0, es); f.tx_lo_p = row(P->peer[0], P->dpeer[0], 1, 1, es);
      f.rx_lo_u = row((char*)P->arena, P->desc, 0, 0, er); f.rx_lo_p = row((char*)P->arena, P->desc, 0, 1, er);
    }
    if (f.has_hi) {
      f.tx_hi_u = row(P->peer[1], P->dpeer[1], 0, 0, es); f.tx_hi_p = row(P->peer[1], P->dpeer[1], 0, 1, es);
      f.rx_hi_u = row((char*)P->arena, P->desc, 1, 0, er); f.rx_hi_p = row((char*)P->arena, P->desc, 1, 1, er);
    }
    return f;
  }
  if (f.has_lo) {  // my first owned row -> rank-1's upper halo row; it bumps my flags[3], I bump its flags[4]
    const AcDesc& d = P->dpeer[0];
    f.lo_u = (double*)(P->peer[0] + desc_off(d, arr_u, idx_u)) + (d.Hl - 1) * d.ld;
    f.lo_p = (double*)(P->peer[0] + desc_off(d, arr_p, idx_p)) + (d.Hl - 1) * d.ld;
    f.sig_lo = (unsigned long long*)(P->peer[0] + d.off_flags) + 4;
    f.expect_lo = P->exp_lo;
    P->exp_lo += (unsigned long long)(fkind == 2 ? d.n_edge_hi_w : fkind == 1 ? d.n_edge_hi_f : d.n_edge_hi);
  }
  if (f.has_hi) {  // my last owned row -> rank+1's lower halo row (its local row 0)
    const AcDesc& d = P->dpeer[1];
    f.hi_u = (double*)(P->peer[1] + desc_off(d, arr_u, idx_u));
    f.hi_p = (double*)(P->peer[1] + desc_off(d, arr_p, idx_p));
    f.sig_hi = (unsigned long long*)(P->peer[1] + d.off_flags) + 3;
    f.expect_hi = P->exp_hi;
    P->exp_hi += (unsigned long long)(fkind == 2 ? d.n_edge_lo_w : fkind == 1 ? d.n_edge_lo_f : d.n_edge_lo);
  }
  f.my_flags = (unsigned long long*)((char*)P->arena + P->desc.off_flags);
  return f;
}

// Two-stream pipeline of the two-step path: box-pair launches stay on the context's stream (A), frame launches go to a
// second stream (B).  Within a pair they are independent -- the box kernel reads the two previous time levels only and
// the FIRST frame launch of a pair is a wide one (it also recomputes the rim ring of the box, so the second frame launch
// depends on it alone).  Across pairs the streams leapfrog: A(p) waits for B(p-1), B(p) waits for A(p-1).
struct TbPipe {
  adseis_acoustic_plan* P;
  bool on = false, box_first = false;
  int n = 0;   // index of the current pair
  int begin() {
    if (!P->tb_overlap) return ADSEIS_OK;
    if (!P->sb) {
      CUDA_TRY(cudaStreamCreateWithFlags(&P->sb, cudaStreamNonBlocking));
      for (int k = 0; k < 2; k++) {
        CUDA_TRY(cudaEventCreateWithFlags(&P->evA[k], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&P->evB[k], cudaEventDisableTiming));
      }
    }
    on = true; n = 0;
    // slab plans: the box-pair kernel is the critical path of a pair (two launches that become eligible together are
    // started ~3 us apart), so it is enqueued first; on one GPU the frames go first (the measured configuration)
    const char* e = getenv("ADSEIS_AC_TB_BOXFIRST");
    box_first = e ? e[0] == '1' : P->arena != nullptr;
    // everything enqueued on A so far (memsets, copies, earlier launches) precedes the first frames
    CUDA_TRY(cudaEventRecord(P->evA[1], P->ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(P->sb, P->evA[1], 0));
    return ADSEIS_OK;
  }
  cudaStream_t frames() const { return on ? P->sb : P->ctx->stream; }
  int pre_frames() {   // frames of pair n follow the box launch of pair n-1
    if (on && n > 0) CUDA_TRY(cudaStreamWaitEvent(P->sb, P->evA[(n - 1) & 1], 0));
    return ADSEIS_OK;
  }
  int post_frames() {
    if (on) CUDA_TRY(cudaEventRecord(P->evB[n & 1], P->sb));
    return ADSEIS_OK;
  }
  int pre_box() {      // the box launch of pair n follows the frames of pair n-1
    if (on && n > 0) CUDA_TRY(cudaStreamWaitEvent(P->ctx->stream, P->evB[(n - 1) & 1], 0));
    return ADSEIS_OK;
  }
  int post_box() {
    if (on) CUDA_TRY(cudaEventRecord(P->evA[n & 1], P->ctx->stream));
    return ADSEIS_OK;
  }
  void next() { n++; }
  int end() {          // join: later work on A sees every frame launch
    if (on && n > 0) CUDA_TRY(cudaStreamWaitEvent(P->ctx->stream, P->evB[(n - 1) & 1], 0));
    on = false;
    return ADSEIS_OK;
  }
};

// one forward step s (slot s from slots s-1, s-2) with the one-step kernel.  fkind 0: the full tiling (marching + frame
// CTAs); 1: only the cells outside the two-step box -- the box cells of that slot are written by ac_fwd2_kernel; 2: the
// same plus the rim ring of the box
static int launch_forward_step(adseis_acoustic_plan* P, i64 base, i64 s, bool sample, int fkind, cudaStream_t st) {
  const AcGeom& g = P->g;
  AcPoints none{};
  AcFuse fuse;
  if (fkind && !P->arena) memset(&fuse, 0, sizeof(fuse));
  else {
    fuse = make_fuse(P, AR_HIST, s - base, AR_PHI, s & 1, fkind, P->ll_prev_kind == 1 && P->ll_prev_s == s - 1,
                     win_slot(P, base, s - 1), P->phi[(s - 1) & 1]);
    P->ll_prev_kind = 1; P->ll_prev_s = s;
  }
  const adseis_acoustic_plan::FrameKind* F = fkind ? &P->fk[fkind - 1] : nullptr;
  const AcPoints srcp = fkind == 2 ? F->srcNp : (F ? F->srcp : P->srcp), rcvp = F ? F->rcvp : P->rcvp;
  if (fkind == 2) fuse.rim = F->srcRp;
#ifdef ADSEIS_TIMELINE
  fuse.tl = tl_slot();
#endif
  CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_slab() || (fkind && adseis_pdl_tb_slab()),
                       F ? (P->p.PropagatorKernel == 0 ? ac_fwd_kernel<0, 1> : ac_fwd_kernel<1, 1>)
                         : (P->p.PropagatorKernel == 0 ? ac_fwd_kernel<0, 0> : ac_fwd_kernel<1, 0>),
                       F ? F->nblocks : P->nblocks, F ? AC_FO_THREADS : AC_FWD_THREADS, F ? 0 : AC_FWD_SMEM, st,
      g, F ? F->t : P->t, win_slot(P, base, s - 1), win_slot(P, base, s - 2), P->c2, P->phi[(s - 1) & 1],
      P->psi[(s - 1) & 1], P->sigx, P->tauy, win_slot(P, base, s), P->phi[s & 1], P->psi[s & 1], srcp,
      P->nsrc > 0 ? P->srcv + (s - 1) * P->nsrc : nullptr, sample ? rcvp : none,
      (sample && P->nrcv > 0) ? P->rcvv + s * P->nrcv : nullptr, fuse));
  LAUNCH_CHECK(P);
  TL_MARK(st, fkind, s);
  if (P->arena && P->unfused) {
    const int arr[3] = {AR_HIST, AR_PHI, AR_PSI}, dep[3] = {2, 2, 1};
    const i64 idx[3] = {s - base, s & 1, s & 1};
    TRY(halo_exchange(P, 3, arr, idx, dep));
  }
  return ADSEIS_OK;
}

static int run_forward_steps(adseis_acoustic_plan* P, i64 base, i64 s_first, i64 s_last, bool sample) {
  const AcGeom& g = P->g;
  cudaStream_t st = P->ctx->stream;
  AcPoints none{};
  if (P->arena && !P->connected) {
    adseis_set_error("acoustic slab plan: adseis_acoustic_plan_ipc_connect has not been called");
    return ADSEIS_ESTATE;
  }
  i64 s = s_first;
  if (P->tb) {
    // pairs of steps (s, s+1): frame of s, frame of s+1, box of s and s+1 in one launch.  The box launch reads time
    // levels s-1 and s-2 only.  One stream: narrow frame, box, narrow frame (the second frame launch reads the box cells
    // of slot s next to the frame).  Two streams: wide frame + narrow frame on B, box on A, concurrently.
    TbPipe pipe{P};
    TRY(pipe.begin());
    for (; s + 1 <= s_last; s += 2) {
      auto frames = [&]() -> int {
        TRY(pipe.pre_frames());
        TRY(launch_forward_step(P, base, s, sample, 2, pipe.frames()));
        TRY(launch_forward_step(P, base, s + 1, sample, 1, pipe.frames()));
        TRY(pipe.post_frames());
        return ADSEIS_OK;
      };
      auto box = [&]() -> int {
        TRY(pipe.pre_box());
        CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_tb_slab() == 2, ac_fwd2_kernel, P->nblocks2, AC_FWD_THREADS, AC_FWD2_SMEM, st,
            g, P->t2, win_slot(P, base, s - 1), win_slot(P, base, s - 2), P->c2, win_slot(P, base, s), win_slot(P, base, s + 1),
            P->srcHp, P->nsrc > 0 ? P->srcv + (s - 1) * P->nsrc : nullptr, P->srcMp,
            P->nsrc > 0 ? P->srcv + s * P->nsrc : nullptr, sample ? P->rcvMp : none,
            (sample && P->nrcv > 0) ? P->rcvv + s * P->nrcv : nullptr,
            (sample && P->nrcv > 0) ? P->rcvv + (s + 1) * P->nrcv : nullptr
#ifdef ADSEIS_TIMELINE
            , tl_slot()
#endif
            ));
        LAUNCH_CHECK(P);
        TL_MARK(st, 3, s);
        TRY(pipe.post_box());
        return ADSEIS_OK;
      };
      if (pipe.on) {
        if (pipe.box_first) { TRY(box()); TRY(frames()); }
        else { TRY(frames()); TRY(box()); }
        pipe.next();
      } else {   // one stream: narrow frame, box, narrow frame (the second frame launch reads the box cells of slot s)
        TRY(launch_forward_step(P, base, s, sample, 1, st));
        TRY(box());
        TRY(launch_forward_step(P, base, s + 1, sample, 1, st));
      }
    }
    TRY(pipe.end());
  }
  for (; s <= s_last; s++) TRY(launch_forward_step(P, base, s, sample, 0, st));
  return ADSEIS_OK;
}

typedef int (*segment_cb)(adseis_acoustic_plan* P, size_t k, void* user);

// ---- whole sweep in one cooperative launch (small grids) --------------------------------------------------------------
static bool use_persist(adseis_acoustic_plan* P) {
  return P->persist && !P->arena && P->seg_b.size() == 1 && !(P->p.PropagatorKernel == 0 && P->k0_corr);
}
static AcPersist persist_args(adseis_acoustic_plan* P) {
  AcPersist a;
  memset(&a, 0, sizeof(a));
  a.hist = P->hist; a.plane = P->g.plane;
  for (int k = 0; k < 2; k++) { a.phi[k] = P->phi[k]; a.psi[k] = P->psi[k]; a.phib[k] = P->phib[k]; a.psib[k] = P->psib[k]; a.ut[k] = P->ut[k]; }
  a.c2 = P->c2; a.sigx = P->sigx; a.tauy = P->tauy;
  auto view = [](const PointSetStorage& q) { return q.nu > 0 ? AcPoints{q.blk, q.cell, q.start, q.perm} : AcPoints{}; };
  a.src = view(P->srcP); a.rcv = view(P->rcvP);
  a.srcv = P->srcv; a.nsrc = (int)P->nsrc; a.nrcv = (int)P->nrcv;
  for (int k = 0; k < 4; k++) a.ub[k] = P->ub[k];
  a.nub = P->nub;
  a.G = P->G; a.res = P->res; a.gradsrcv = P->gradsrcv;
  a.bar = P->pbar;
  return a;
}
template <class K>
static int launch_persist(adseis_acoustic_plan* P, K kernel, AcPersist a) {
  cudaStream_t st = P->ctx->stream;
  const size_t nb = (size_t)P->nblocksP + 1;
  if (!P->pbar) TRY(dev_alloc_zero(&P->pbar, nb, st));
  a.bar = P->pbar;
  CUDA_TRY(cudaMemsetAsync(P->pbar, 0, nb * 8, st));
  {
    // grid barrier (default) or per-CTA progress flags of the neighbours only (ADSEIS_AC_PERSIST_SYNC=1; a step reads one
    // row / column around a cell with scheme 1, two with scheme 0).  Measured on B200, C1: the flags are SLOWER -- 12.0 vs
    // 8.1 us per step pair (a release fence + up to nine acquire polls per CTA and step against one atomic + one poll)
    const char* e = getenv("ADSEIS_AC_PERSIST_SYNC");
    const i64 per = (i64)P->tp.fthr * P->tp.fcpt, reach = (P->p.PropagatorKernel == 0 ? 2 : 1) * (i64)P->g.W + 2;
    const i64 dep = (reach + per - 1) / per;
    if ((e && e[0] == '1') && 2 * dep + 1 <= AC_PS_THREADS) {
      a.prog = P->pbar + 1; a.dep_lo = (int)dep; a.dep_hi = (int)dep;
    }
  }
  AcGeom g = P->g;
  AcTiling t = P->tp;
  void* args[3] = {&g, &t, &a};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kernel, dim3(P->nblocksP), dim3(AC_PS_THREADS), args, 0, st));
  LAUNCH_CHECK(P);
  return ADSEIS_OK;
}

// Full forward sweep, segment by segment.  After segment k finishes (slots seg_b[k]..seg_e[k] in the window)
// `cb` is invoked (may be null).  Saves the start state of every later segment when save_ckpt.
static int forward_sweep(adseis_acoustic_plan* P, bool save_ckpt, segment_cb cb, void* user) {
  const AcGeom& g = P->g;
  cudaStream_t st = P->ctx->stream;
  const size_t pb = (size_t)g.plane * 8;
  for (int k = 0; k < 2; k++) {
    CUDA_TRY(cudaMemsetAsync(P->phi[k], 0, pb, st));
    CUDA_TRY(cudaMemsetAsync(P->psi[k], 0, pb, st));
  }
  CUDA_TRY(cudaMemsetAsync(P->hist, 0, 2 * pb, st));  // slots 0,1 = 0 (Core.jl:607-612)
  if (P->nrcv > 0) CUDA_TRY(cudaMemsetAsync(P->rcvv, 0, (size_t)(2 * P->nrcv) * 8, st));
  TRY(halo_exchange(P, 0, nullptr, nullptr));  // slab plans: nobody pushes before everybody's memsets are done
  const size_t nseg = P->seg_b.size();
  for (size_t k = 0; k < nseg; k++) {
    const i64 b = P->seg_b[k], e = P->seg_e[k];
    if (k > 0) {
      // slab plans: the neighbours' last pushes must have landed before halo rows are copied.  Packed halo rows are
      // unpacked by the NEXT step launch, so the halo rows of the last slot and of its phi are exchanged explicitly
      // here (they go into the checkpoint and into the rebased window)
      const i64 pbse = P->seg_b[k - 1];
      if (P->ll) {
        const int arr[2] = {AR_HIST, AR_PHI};
        const i64 idx[2] = {b + 1 - pbse, (b + 1) & 1};
        TRY(halo_exchange(P, 2, arr, idx));
      } else {
        TRY(halo_exchange(P, 0, nullptr, nullptr));
      }
      // window index 0,1 <- slots b, b+1 (the last two slots of the previous segment)
      double* s0 = win_slot(P, pbse, b);
      double* s1 = win_slot(P, pbse, b + 1);
      if (save_ckpt) {
        double* c = P->ckpt[k - 1];
        CUDA_TRY(cudaMemcpyAsync(c, s0, pb, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c + g.plane, s1, pb, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c + 2 * g.plane, P->phi[(b + 1) & 1], pb, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c + 3 * g.plane, P->psi[(b + 1) & 1], pb, cudaMemcpyDeviceToDevice, st));
      }
      CUDA_TRY(cudaMemcpyAsync(P->hist, s0, pb, cudaMemcpyDeviceToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(P->hist + g.plane, s1, pb, cudaMemcpyDeviceToDevice, st));
    }
    TRY(span_begin(P, 0, e - (b + 2) + 1));
    if (use_persist(P) && e >= b + 2) {
      AcPersist a = persist_args(P);
      a.s_first = b + 2; a.s_last = e; a.rcvv = P->rcvv;
      if (P->p.PropagatorKernel == 0) TRY(launch_persist(P, ac_fwd_persist_kernel<0>, a));
      else TRY(launch_persist(P, ac_fwd_persist_kernel<1>, a));
    } else {
      TRY(run_forward_steps(P, b, b + 2, e, true));
    }
    TRY(span_end(P));
    P->win_base = b; P->win_last = e;
    if (cb) TRY(cb(P, k, user));
  }
  return ADSEIS_OK;
}

static int check_ready(adseis_acoustic_plan* P, bool need_obs) {
  if (!P->have_model || (!P->have_srcv && P->nsrc > 0) || (need_obs && !P->have_obs)) {
    adseis_set_error("acoustic plan: set_model / set_srcv%s must be called first", need_obs ? " / set_obs" : "");
    return ADSEIS_ESTATE;
  }
  return ADSEIS_OK;
}

// ---- CUDA-graph replay of a whole sweep ---------------------------------------------------------------------
static bool graphs_enabled(adseis_acoustic_plan* P) {
  if (P->graphs < 0) {
    const char* e = getenv("ADSEIS_GRAPHS");
    // worth it when a step kernel is short next to its launch cost; large grids keep direct launches (a graph of
    // tens of thousands of nodes costs more to build than it saves)
    const bool small = (i64)P->g.H * P->g.W <= (6LL << 20);
    P->graphs = (P->arena == nullptr) && !use_persist(P) && (e ? e[0] != '0' : small) ? 1 : 0;   // (a cooperative launch is not captured)
  }
  return P->graphs == 1;
}

static unsigned long long graph_key(adseis_acoustic_plan* P, int kind) {
  unsigned long long h = 1469598103934665603ULL;
  auto mix = [&](const void* p) { h ^= (unsigned long long)(uintptr_t)p; h *= 1099511628211ULL; };
  mix(P->hist); mix(P->c2); mix(P->phi[0]); mix(P->psi[0]); mix(P->srcv); mix(P->rcvv); mix(P->perm);
  const PointSetStorage* ps[] = {&P->src, &P->rcv, &P->fk[0].src, &P->fk[0].rcv, &P->fk[1].src, &P->fk[1].rcv, &P->srcM, &P->rcvM, &P->srcH, &P->rcvH};
  for (const PointSetStorage* q : ps) { mix(q->blk); mix(q->cell); mix(q->start); mix(q->perm); mix((void*)(uintptr_t)(q->nu + 1)); }
  mix((void*)(uintptr_t)(P->nsrc * 131 + P->nrcv + 7)); mix((void*)(uintptr_t)(P->k0_corr ? 3 : 5));
  if (kind == 1) {
    mix(P->obs); mix(P->res); mix(P->rcv_owned); mix(P->loss); mix(P->G); mix(P->gradc); mix(P->gradsrcv); mix(P->cvel);
    for (int k = 0; k < 4; k++) mix(P->ub[k]);
    mix(P->phib[0]); mix(P->psib[0]); mix(P->ut[0]);
    for (double* c : P->ckpt) mix(c);
  }
  return h | 1ULL;
}

// Run `body` (which only enqueues work on the context's stream) directly, or -- when graphs are on -- capture it once
// into a CUDA graph and replay that.  Any failure of the capture path switches graphs off for this plan and falls back.
template <class Body>
static int run_captured(adseis_acoustic_plan* P, int kind, Body body) {
  cudaStream_t st = P->ctx->stream;
  P->last_launches = 0; P->last_recomputed = 0;
  P->spans.clear(); P->ev_used = 0;
  if (!graphs_enabled(P)) return body();
  adseis_acoustic_plan::GraphSlot& gs = P->gslot[kind];
  const unsigned long long key = graph_key(P, kind);
  if (gs.exec && gs.key != key) { cudaGraphExecDestroy(gs.exec); gs.exec = nullptr; }
  if (!gs.exec) {
    while (P->ev_pool.size() < 6 * P->seg_b.size() + 8) {   // no event creation inside the capture
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreate(&e));
      P->ev_pool.push_back(e);
    }
    const i64 l0 = P->ctx->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); P->graphs = 0; return body(); }
    P->capturing = true;
    const int rc = body();
    P->capturing = false;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    P->ctx->launches = l0;   // nothing has run yet
    if (rc != ADSEIS_OK || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      P->graphs = 0;
      if (rc != ADSEIS_OK) return rc;
      P->last_launches = 0; P->last_recomputed = 0; P->spans.clear(); P->ev_used = 0;
      return body();
    }
    const cudaError_t ie = cudaGraphInstantiate(&gs.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      cudaGetLastError();
      gs.exec = nullptr; P->graphs = 0;
      P->last_launches = 0; P->last_recomputed = 0; P->spans.clear(); P->ev_used = 0;
      return body();
    }
    gs.key = key; gs.launches = P->last_launches; gs.recomputed = P->last_recomputed; gs.ev_used = P->ev_used;
    gs.spans_copy = P->spans;
  } else {
    P->last_launches = gs.launches; P->last_recomputed = gs.recomputed; P->ev_used = gs.ev_used;
    P->spans = gs.spans_copy;
  }
  CUDA_TRY(cudaGraphLaunch(gs.exec, st));
  P->ctx->launches += gs.launches;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_forward(adseis_acoustic_plan* P) {
  REQUIRE(P, "acoustic_plan_forward: null");
  TRY(check_ready(P, false));
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  TRY(run_captured(P, 0, [&]() { return forward_sweep(P, false, nullptr, nullptr); }));
  P->last_segments = (i64)P->seg_b.size();
  P->win_base = P->seg_b.back(); P->win_last = P->seg_e.back();
  P->have_fwd = true;
  return ADSEIS_OK;
}

static int ensure_adjoint_state(adseis_acoustic_plan* P) {
  if (P->G) {
    if (!P->gradsrcv) TRY(dev_alloc_zero(&P->gradsrcv, (size_t)(P->p.NSTEP * P->nsrc), P->ctx->stream));
    return ADSEIS_OK;
  }
  const size_t n = (size_t)P->g.plane;
  cudaStream_t st = P->ctx->stream;
  if (!P->arena) {
    for (int k = 0; k < P->nub; k++) TRY(dev_alloc_zero(&P->ub[k], n, st));
    for (int k = 0; k < 2; k++) { TRY(dev_alloc_zero(&P->phib[k], n, st)); TRY(dev_alloc_zero(&P->psib[k], n, st)); }
  }
  if (P->p.PropagatorKernel == 0)
    for (int k = 0; k < 2; k++) TRY(dev_alloc_zero(&P->ut[k], n, st));
  TRY(dev_alloc_zero(&P->G, n, st));
  TRY(dev_alloc_zero(&P->gradc, (size_t)P->model_elems, st));
  TRY(dev_alloc_zero(&P->gradsrcv, (size_t)(P->p.NSTEP * P->nsrc), st));
  return ADSEIS_OK;
}

static int gradient_body(adseis_acoustic_plan* P);

ADSEIS_API int adseis_acoustic_plan_gradient(adseis_acoustic_plan* P) {
  REQUIRE(P, "acoustic_plan_gradient: null");
  TRY(check_ready(P, true));
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  TRY(ensure_adjoint_state(P));
  TRY(run_captured(P, 1, [&]() { return gradient_body(P); }));
  P->last_segments = (i64)P->seg_b.size();
  P->win_base = P->seg_b.front(); P->win_last = P->seg_e.front();   // the reverse sweep ends in the first segment
  P->have_fwd = true;
  P->have_grad = true;
  return ADSEIS_OK;
}

static int gradient_body(adseis_acoustic_plan* P) {
  const AcGeom& g = P->g;
  cudaStream_t st = P->ctx->stream;
  const i64 NSTEP = P->p.NSTEP;
  const size_t pb = (size_t)g.plane * 8;
  const i64 NUB = P->nub;
  // ---- forward, keeping checkpoints; the last segment stays in the window
  TRY(forward_sweep(P, true, nullptr, nullptr));
  const size_t nseg = P->seg_b.size();
  P->last_segments = (i64)nseg;
  // ---- misfit and adjoint sources
  const i64 nr = (NSTEP + 1) * P->nrcv;
  k_residual_partial<<<RL_BLOCKS, 256, 0, st>>>(P->rcvv, P->obs, P->rcv_owned, (int)std::max<i64>(P->nrcv, 1), 1, nr, P->res, P->loss);
  LAUNCH_CHECK(P);
  k_residual_final<<<1, RL_BLOCKS, 0, st>>>(P->loss);
  LAUNCH_CHECK(P);
  // ---- reverse sweep
  for (int k = 0; k < P->nub; k++) CUDA_TRY(cudaMemsetAsync(P->ub[k], 0, pb, st));
  for (int k = 0; k < 2; k++) { CUDA_TRY(cudaMemsetAsync(P->phib[k], 0, pb, st)); CUDA_TRY(cudaMemsetAsync(P->psib[k], 0, pb, st)); }
  CUDA_TRY(cudaMemsetAsync(P->G, 0, pb, st));
  for (int k = 0; k < 2; k++) if (P->ut[k]) CUDA_TRY(cudaMemsetAsync(P->ut[k], 0, pb, st));
  if (P->nsrc > 0) CUDA_TRY(cudaMemsetAsync(P->gradsrcv, 0, (size_t)(NSTEP * P->nsrc) * 8, st));
  TRY(halo_exchange(P, 0, nullptr, nullptr));
  // ubar[NSTEP] = receiver term only; grad_srcv row NSTEP-1
  if (P->rcv.nu > 0) {
    k_points_inject<<<(P->rcv.nu + 127) / 128, 128, 0, st>>>(P->ub[NSTEP % NUB], P->rcvp, P->rcv.nu,
                                                             P->res + NSTEP * P->nrcv, 1.0);
    LAUNCH_CHECK(P);
  }
  if (P->src.nu > 0 && NSTEP - 1 >= 1) {
    k_points_sample<<<(P->src.nu + 127) / 128, 128, 0, st>>>(P->ub[NSTEP % NUB], P->srcp, P->src.nu,
                                                             P->gradsrcv + (NSTEP - 1) * P->nsrc, g.dt2);
    LAUNCH_CHECK(P);
  }
  if (P->arena) {
    const int arr[1] = {AR_UB};
    const i64 idx[1] = {NSTEP % NUB};
    TRY(halo_exchange(P, 1, arr, idx));
  }
  AcPoints none{};
  for (i64 k = (i64)nseg - 1; k >= 0; k--) {
    const i64 b = P->seg_b[k], e = P->seg_e[k];
    if (k != (i64)nseg - 1) {
      // restore the start state of segment k and recompute its forward steps (bit-identical replay)
      // (packed halo rows: the first replay launch receives nothing, so nothing proves that the neighbour has consumed
      // the words its own last adjoint launch still needs before this rank overwrites them two launches later -- a
      // handshake closes that window)
      if (P->arena && P->ll) TRY(halo_exchange(P, 0, nullptr, nullptr));
      if (k == 0) {
        CUDA_TRY(cudaMemsetAsync(P->hist, 0, 2 * pb, st));
        CUDA_TRY(cudaMemsetAsync(P->phi[1], 0, pb, st));
        CUDA_TRY(cudaMemsetAsync(P->psi[1], 0, pb, st));
      } else {
        double* c = P->ckpt[k - 1];
        CUDA_TRY(cudaMemcpyAsync(P->hist, c, pb, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(P->hist + g.plane, c + g.plane, pb, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(P->phi[(b + 1) & 1], c + 2 * g.plane, pb, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(P->psi[(b + 1) & 1], c + 3 * g.plane, pb, cudaMemcpyDeviceToDevice, st));
      }
      TRY(span_begin(P, 1, e - (b + 2) + 1));
      TRY(run_forward_steps(P, b, b + 2, e, false));
      TRY(span_end(P));
      P->last_recomputed += e - (b + 2) + 1;
      P->win_base = b; P->win_last = e;
      if (P->arena && P->ll) {
        // packed halo rows: the adjoint rows sent by the last adjoint launch before the replay were never unpacked
        const int arr[2] = {AR_UB, AR_PHIB};
        const i64 idx[2] = {e % NUB, e & 1};
        TRY(halo_exchange(P, 2, arr, idx));
      }
    }
    TRY(span_begin(P, 2, e - (b + 2) + 1));
    // one adjoint step s: ubar[s-1] from ubar[s], ubar[s+1], u[s-1]; fkind as in launch_forward_step
    auto adj_step = [&](i64 s, int fkind, cudaStream_t stl) -> int {
      AcFuse fuse;
      if (fkind && !P->arena) memset(&fuse, 0, sizeof(fuse));
      else {
        fuse = make_fuse(P, AR_UB, (s - 1 + NUB) % NUB, AR_PHIB, (s - 1) & 1, fkind, P->ll_prev_kind == 2 && P->ll_prev_s == s + 1,
                         P->ub[s % NUB], P->phib[s & 1]);
        P->ll_prev_kind = 2; P->ll_prev_s = s;
      }
      AcK0 k0{};
      if (P->p.PropagatorKernel == 0) {
        k0.wnew = win_slot(P, b, s); k0.ut_in = P->ut[(s + 1) & 1]; k0.ut_out = P->ut[s & 1]; k0.ub2 = P->ub[(s + 1) % NUB];
        const PointSetStorage& K = P->unfused ? P->srcK : P->src;   // slabs: incl. the neighbours' sources next to my rows
        if (P->k0_corr && K.nu > 0) {
          k_ac_k0_src_corr<<<(K.nu + 127) / 128, 128, 0, stl>>>(g, K.cell, K.start, K.perm, K.nu,
                                                                P->srcv + (s - 1) * P->nsrc, P->phib[s & 1],
                                                                P->psib[s & 1], P->sigx, P->tauy, P->G, P->own0, P->own1);
          LAUNCH_CHECK(P);
        }
      }
      const adseis_acoustic_plan::FrameKind* F = fkind ? &P->fk[fkind - 1] : nullptr;
      const AcPoints rcvp = fkind == 2 ? F->rcvNp : (F ? F->rcvp : P->rcvp), srcp = F ? F->srcp : P->srcp;
      if (fkind == 2) fuse.rim = F->rcvRp;
#ifdef ADSEIS_TIMELINE
      fuse.tl = tl_slot();
#endif
      CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_slab() || (fkind && adseis_pdl_tb_slab()),
                           F ? (P->p.PropagatorKernel == 0 ? ac_adj_kernel<0, 1> : ac_adj_kernel<1, 1>)
                             : (P->p.PropagatorKernel == 0 ? ac_adj_kernel<0, 0> : ac_adj_kernel<1, 0>),
                           F ? F->nblocks : P->nblocks, F ? AC_FO_THREADS : AC_ADJ_THREADS, F ? 0 : AC_ADJ_SMEM, stl,
          g, F ? F->t : P->t, P->ub[s % NUB], P->ub[(s + 1) % NUB], win_slot(P, b, s - 1), P->c2, P->phib[s & 1],
          P->psib[s & 1], P->sigx, P->tauy, P->ub[(s - 1 + NUB) % NUB], P->phib[(s - 1) & 1], P->psib[(s - 1) & 1], P->G,
          rcvp, P->nrcv > 0 ? P->res + (s - 1) * P->nrcv : nullptr, (s - 2 >= 1) ? srcp : none,
          (s - 2 >= 1 && P->nsrc > 0) ? P->gradsrcv + (s - 2) * P->nsrc : nullptr, fuse, k0));
      LAUNCH_CHECK(P);
      TL_MARK(stl, 10 + fkind, s);
      if (P->arena && P->unfused) {
        const int arr[3] = {AR_UB, AR_PHIB, AR_PSIB}, dep[3] = {1, 2, 1};
        const i64 idx[3] = {(s - 1 + NUB) % NUB, (s - 1) & 1, (s - 1) & 1};
        TRY(halo_exchange(P, 3, arr, idx, dep));
      }
      return ADSEIS_OK;
    };
    i64 s = e;
    if (use_persist(P) && e >= b + 2) {
      AcPersist a = persist_args(P);
      a.s_first = b + 2; a.s_last = e;
      if (P->p.PropagatorKernel == 0) TRY(launch_persist(P, ac_adj_persist_kernel<0>, a));
      else TRY(launch_persist(P, ac_adj_persist_kernel<1>, a));
      s = b + 1;
    }
    if (P->tb_adj) {
      // pairs (s, s-1): frame of step s, frame of step s-1, box of steps s and s-1 in one launch; streams as in the
      // forward sweep (one stream: narrow frame, box, narrow frame)
      TbPipe pipe{P};
      TRY(pipe.begin());
      for (; s - 1 >= b + 2; s -= 2) {
        auto frames = [&]() -> int {
          TRY(pipe.pre_frames());
          TRY(adj_step(s, 2, pipe.frames()));
          TRY(adj_step(s - 1, 1, pipe.frames()));
          TRY(pipe.post_frames());
          return ADSEIS_OK;
        };
        auto box = [&]() -> int {
          TRY(pipe.pre_box());
          CUDA_TRY(launch_step(P->arena == nullptr || adseis_pdl_tb_slab() == 2, ac_adj2_kernel, P->nblocks2, AC2_THREADS, AC_ADJ2_SMEM, st,
              g, P->t2, P->ub[s % NUB], P->ub[(s + 1) % NUB], win_slot(P, b, s - 1), win_slot(P, b, s - 2), P->c2,
              P->ub[(s - 1 + NUB) % NUB], P->ub[(s - 2 + NUB) % NUB], P->G, P->rcvHp,
              P->nrcv > 0 ? P->res + (s - 1) * P->nrcv : nullptr, P->rcvMp, P->nrcv > 0 ? P->res + (s - 2) * P->nrcv : nullptr,
              P->srcMp, (s - 2 >= 1 && P->nsrc > 0) ? P->gradsrcv + (s - 2) * P->nsrc : nullptr,
              (s - 3 >= 1 && P->nsrc > 0) ? P->gradsrcv + (s - 3) * P->nsrc : nullptr
#ifdef ADSEIS_TIMELINE
              , tl_slot()
#endif
              ));
          LAUNCH_CHECK(P);
          TL_MARK(st, 13, s);
          TRY(pipe.post_box());
          return ADSEIS_OK;
        };
        if (pipe.on) {
          if (pipe.box_first) { TRY(box()); TRY(frames()); }
          else { TRY(frames()); TRY(box()); }
          pipe.next();
        } else {
          TRY(adj_step(s, 1, st));
          TRY(box());
          TRY(adj_step(s - 1, 1, st));
        }
      }
      TRY(pipe.end());
    }
    for (; s >= b + 2; s--) TRY(adj_step(s, 0, st));
    TRY(span_end(P));
  }
  CUDA_TRY(cudaMemsetAsync(P->gradc, 0, (size_t)P->model_elems * 8, st));
  {
    dim3 gg((unsigned)((g.W + 127) / 128), (unsigned)(P->own1 - P->own0));
    k_grad_finalize<<<gg, 128, 0, st>>>(P->G, P->cvel, g.ld, g.goff, P->own0, P->own1, g.H, g.W,
                                        P->p.mpi_convention, P->gradc);
    LAUNCH_CHECK(P);
  }
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_get(adseis_acoustic_plan* P, int what, double* dst, int to_device) {
  (void)to_device;
  REQUIRE(P && dst, "acoustic_plan_get: null");
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  const double* src = nullptr;
  size_t n = 0;
  switch (what) {
    case ADSEIS_GET_RCVV:
      if (!P->have_fwd) { adseis_set_error("acoustic_plan_get: no forward results yet"); return ADSEIS_ESTATE; }
      src = P->rcvv; n = (size_t)((P->p.NSTEP + 1) * P->nrcv); break;
    case ADSEIS_GET_LOSS: src = P->loss; n = 1; break;
    case ADSEIS_GET_GRAD_C: src = P->gradc; n = (size_t)P->model_elems; break;
    case ADSEIS_GET_GRAD_SRCV: src = P->gradsrcv; n = (size_t)(P->p.NSTEP * P->nsrc); break;
    case ADSEIS_GET_GRAD_C_OWNED: {
      const i64 H = P->g.H, W = P->g.W;
      if (P->p.mpi_convention) {
        const i64 r0 = std::max<i64>(P->slab.row0, 1), r1 = std::min<i64>(P->slab.row1, H - 1);
        src = P->gradc + (r0 - 1) * (W - 2); n = (size_t)(std::max<i64>(r1 - r0, 0) * (W - 2));
      } else {
        src = P->gradc + P->slab.row0 * W; n = (size_t)((P->slab.row1 - P->slab.row0) * W);
      }
    } break;
    default: adseis_set_error("acoustic_plan_get: unknown item %d", what); return ADSEIS_EINVAL;
  }
  if (what != ADSEIS_GET_RCVV && !P->have_grad) {
    adseis_set_error("acoustic_plan_get: no gradient results yet");
    return ADSEIS_ESTATE;
  }
  if (n > 0) CUDA_TRY(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDefault, P->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(P->ctx->stream));
  TRY(halo_check(P));
  return ADSEIS_OK;
}

static int copy_snapshot_out(adseis_acoustic_plan* P, const double* plane, double* dst) {
  const AcGeom& g = P->g;
  cudaStream_t st = P->ctx->stream;
  if (!P->p.mpi_convention) {
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)g.W * 8, plane, (size_t)g.ld * 8, (size_t)g.W * 8, (size_t)g.H,
                               cudaMemcpyDefault, st));
  } else {
    const int NY = g.W - 2;
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)NY * 8, plane + g.ld + 1, (size_t)g.ld * 8, (size_t)NY * 8,
                               (size_t)(g.H - 2), cudaMemcpyDefault, st));
  }
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_get_snapshot(adseis_acoustic_plan* P, int64_t slot, double* dst, int to_device) {
  (void)to_device;
  REQUIRE(P && dst, "acoustic_plan_get_snapshot: null");
  REQUIRE(P->slab.nranks == 1, "acoustic_plan_get_snapshot: single-GPU plans only");
  if (!P->have_fwd || slot < P->win_base || slot > P->win_last) {
    adseis_set_error("acoustic_plan_get_snapshot: slot %lld is not resident (window holds [%lld,%lld])",
                     (long long)slot, (long long)P->win_base, (long long)P->win_last);
    return ADSEIS_ESTATE;
  }
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  TRY(copy_snapshot_out(P, win_slot(P, P->win_base, slot), dst));
  CUDA_TRY(cudaStreamSynchronize(P->ctx->stream));
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_info(adseis_acoustic_plan* P, int64_t info[8]) {
  REQUIRE(P && info, "acoustic_plan_info: null");
  info[0] = P->win; info[1] = P->last_segments; info[2] = P->last_launches; info[3] = P->g.Hl; info[4] = P->g.ld;
  info[5] = P->last_recomputed; info[6] = (i64)P->seg_b.size(); info[7] = P->fast_rows;
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_timings(adseis_acoustic_plan* P, double out[6]) {
  REQUIRE(P && out, "acoustic_plan_timings: null");
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  for (int k = 0; k < 6; k++) out[k] = 0.0;
  for (const auto& sp : P->spans) {
    CUDA_TRY(cudaEventSynchronize(sp.b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, sp.a, sp.b));
    out[sp.phase] += ms;
    out[3 + sp.phase] += (double)sp.launches;
  }
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_ipc_export(adseis_acoustic_plan* P, void* handle_out) {
  REQUIRE(P && handle_out, "acoustic_plan_ipc_export: null");
  if (!P->arena) {
    adseis_set_error("acoustic_plan_ipc_export: not a slab plan (nranks == 1)");
    return ADSEIS_ESTATE;
  }
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, P->arena));
  static_assert(sizeof(h) == ADSEIS_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_plan_ipc_connect(adseis_acoustic_plan* P, const void* lo, const void* hi) {
  REQUIRE(P, "acoustic_plan_ipc_connect: null");
  if (!P->arena) {
    adseis_set_error("acoustic_plan_ipc_connect: not a slab plan (nranks == 1)");
    return ADSEIS_ESTATE;
  }
  CUDA_TRY(cudaSetDevice(P->ctx->device));
  const void* hs[2] = {lo, hi};
  const bool need[2] = {P->slab.rank > 0, P->slab.rank < P->slab.nranks - 1};
  for (int k = 0; k < 2; k++) {
    if (!need[k]) continue;
    REQUIRE(hs[k], "acoustic_plan_ipc_connect: missing handle of rank %d", P->slab.rank + (k ? 1 : -1));
    cudaIpcMemHandle_t h;
    memcpy(&h, hs[k], sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      adseis_set_error("acoustic_plan_ipc_connect: cudaIpcOpenMemHandle(rank %d) -> %s", P->slab.rank + (k ? 1 : -1),
                       cudaGetErrorString(e));
      return ADSEIS_ECOMM;
    }
    P->peer[k] = (char*)ptr;
    CUDA_TRY(cudaMemcpy(&P->dpeer[k], ptr, sizeof(AcDesc), cudaMemcpyDeviceToHost));
    const AcDesc& d = P->dpeer[k];
    if (d.magic != AC_DESC_MAGIC || d.ld != P->desc.ld || d.win != P->desc.win || d.nub != P->desc.nub ||
        (d.n_edge_lo_f + d.n_edge_hi_f > 0) != (P->desc.n_edge_lo_f + P->desc.n_edge_hi_f > 0)) {
      adseis_set_error("acoustic_plan_ipc_connect: neighbour %d has an incompatible layout (magic %llx ld %lld win %lld; "
                       "mine ld %lld win %lld)", P->slab.rank + (k ? 1 : -1), d.magic, d.ld, d.win, P->desc.ld, P->desc.win);
      return ADSEIS_ECOMM;
    }
  }
  P->connected = true;
  return ADSEIS_OK;
}

static long long desc_off(const AcDesc& d, int arr, i64 idx) {
  switch (arr) {
    case AR_HIST: return d.off_hist + idx * d.plane * 8;
    case AR_PHI: return d.off_phi[idx];
    case AR_PSI: return d.off_psi[idx];
    case AR_UB: return d.off_ub[idx];
    case AR_PHIB: return d.off_phib[idx];
    default: return d.off_psib[idx];
  }
}

// Push the first / last owned row of up to two arrays to the neighbours and wait for theirs (nothing to do on a
// single GPU).  narr == 0 is a pure barrier.
static int halo_exchange(adseis_acoustic_plan* P, int narr, const int* arr, const i64* idx, const int* depth) {
  if (!P->arena) return ADSEIS_OK;
  if (!P->connected) {
    adseis_set_error("acoustic slab plan: adseis_acoustic_plan_ipc_connect has not been called");
    return ADSEIS_ESTATE;
  }
  const AcGeom& g = P->g;
  AcHaloArgs a;
  memset(&a, 0, sizeof(a));
  a.ld = g.ld;
  a.has_lo = P->peer[0] != nullptr;
  a.has_hi = P->peer[1] != nullptr;
  char* base = (char*)P->arena;
  for (int k = 0; k < narr; k++) {
    const double* mine = (const double*)(base + desc_off(P->desc, arr[k], idx[k]));
    const int dep = depth ? depth[k] : 1;
    for (int r = 0; r < dep; r++) {
      if (a.has_lo) {  // my first owned rows -> rank-1's upper halo rows (its local rows own1, own1+1, ...)
        const AcDesc& d = P->dpeer[0];
        a.src[a.nrow] = mine + (i64)(P->own0 + r) * g.ld;
        a.dst[a.nrow] = (double*)(P->peer[0] + desc_off(d, arr[k], idx[k])) + (d.own1 + r) * d.ld;
        a.nrow++;
      }
      if (a.has_hi) {  // my last owned rows -> rank+1's lower halo rows (its local rows own0-dep .. own0-1)
        const AcDesc& d = P->dpeer[1];
        a.src[a.nrow] = mine + (i64)(P->own1 - dep + r) * g.ld;
        a.dst[a.nrow] = (double*)(P->peer[1] + desc_off(d, arr[k], idx[k])) + (d.own0 - dep + r) * d.ld;
        a.nrow++;
      }
    }
  }
  // rank-1 sees me as its "hi" neighbour: I add to its flag[1]; rank+1 sees me as "lo": its flag[0]
  if (a.has_lo) a.sig_lo = (unsigned long long*)(P->peer[0] + P->dpeer[0].off_flags) + 1;
  if (a.has_hi) a.sig_hi = (unsigned long long*)(P->peer[1] + P->dpeer[1].off_flags) + 0;
  a.my_flags = (unsigned long long*)(base + P->desc.off_flags);
  P->epoch++;
  P->ll_prev_kind = 0;   // the next fused launch finds its halo rows in place
  a.expect = (unsigned long long)AC_HX_BLOCKS * P->epoch;
  k_halo_exchange<<<AC_HX_BLOCKS, 256, 0, P->ctx->stream>>>(a);
  LAUNCH_CHECK(P);
  return ADSEIS_OK;
}

static int halo_check(adseis_acoustic_plan* P) {
  if (!P->arena) return ADSEIS_OK;
  unsigned long long f[3] = {0, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(f, (char*)P->arena + P->desc.off_flags, sizeof(f), cudaMemcpyDeviceToHost, P->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(P->ctx->stream));
  if (f[2] != 0) {
    adseis_set_error("acoustic slab plan (rank %d): timed out waiting for a neighbour's halo rows", P->slab.rank);
    return ADSEIS_ECOMM;
  }
  return ADSEIS_OK;
}

// ------------------------------------------------------------------------------------------------------------
// one-call host-buffer entry points
// ------------------------------------------------------------------------------------------------------------
struct hist_copy_ctx { double* out; };
static int hist_copy_cb(adseis_acoustic_plan* P, size_t k, void* user) {
  hist_copy_ctx* h = (hist_copy_ctx*)user;
  const i64 b = P->seg_b[k], e = P->seg_e[k];
  for (i64 s = (k == 0 ? b : b + 2); s <= e; s++)
    TRY(copy_snapshot_out(P, win_slot(P, b, s), h->out + s * P->model_elems));
  return ADSEIS_OK;
}

ADSEIS_API int adseis_acoustic_forward(adseis_ctx* ctx, const adseis_acoustic_params* p, const double* c, int64_t nsrc,
                                       const int64_t* srci, const int64_t* srcj, const double* srcv,
                                       int64_t srcv_rows, int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj,
                                       double* rcvv_out, double* u_hist_out) {
  adseis_acoustic_plan* P = nullptr;
  TRY(adseis_acoustic_plan_create(ctx, p, nullptr, nsrc, srci, srcj, nrcv, rcvi, rcvj, 0, &P));
  int rc = adseis_acoustic_plan_set_model(P, c, 0);
  if (!rc) rc = adseis_acoustic_plan_set_srcv(P, srcv, srcv_rows, 0);
  if (!rc) {
    P->last_launches = 0;
    P->spans.clear(); P->ev_used = 0;
    hist_copy_ctx h{u_hist_out};
    rc = forward_sweep(P, false, u_hist_out ? hist_copy_cb : nullptr, &h);
    P->have_fwd = (rc == 0);
  }
  if (!rc && rcvv_out && nrcv > 0) rc = adseis_acoustic_plan_get(P, ADSEIS_GET_RCVV, rcvv_out, 0);
  if (!rc) rc = adseis_ctx_sync(ctx);
  adseis_acoustic_plan_destroy(P);
  return rc;
}

ADSEIS_API int adseis_acoustic_misfit_grad(adseis_ctx* ctx, const adseis_acoustic_params* p, const double* c,
                                           int64_t nsrc, const int64_t* srci, const int64_t* srcj, const double* srcv,
                                           int64_t srcv_rows, int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj,
                                           const double* obs, double* loss_out, double* rcvv_out, double* grad_c_out,
                                           double* grad_srcv_out) {
  adseis_acoustic_plan* P = nullptr;
  TRY(adseis_acoustic_plan_create(ctx, p, nullptr, nsrc, srci, srcj, nrcv, rcvi, rcvj, 0, &P));
  int rc = adseis_acoustic_plan_set_model(P, c, 0);
  if (!rc) rc = adseis_acoustic_plan_set_srcv(P, srcv, srcv_rows, 0);
  if (!rc) rc = adseis_acoustic_plan_set_obs(P, obs, 0);
  if (!rc) rc = adseis_acoustic_plan_gradient(P);
  if (!rc && loss_out) rc = adseis_acoustic_plan_get(P, ADSEIS_GET_LOSS, loss_out, 0);
  if (!rc && rcvv_out && nrcv > 0) rc = adseis_acoustic_plan_get(P, ADSEIS_GET_RCVV, rcvv_out, 0);
  if (!rc && grad_c_out) rc = adseis_acoustic_plan_get(P, ADSEIS_GET_GRAD_C, grad_c_out, 0);
  if (!rc && grad_srcv_out && nsrc > 0) rc = adseis_acoustic_plan_get(P, ADSEIS_GET_GRAD_SRCV, grad_srcv_out, 0);
  if (!rc) rc = adseis_ctx_sync(ctx);
  adseis_acoustic_plan_destroy(P);
  return rc;
}

#ifdef AC_RING_DEBUG
// debug builds only: copy out (and reset) the ring mismatch records of ac_fwd_kernel; returns the count
extern "C" __attribute__((visibility("default"))) int adseis_debug_ring(double* out /* 64*12 */) {
  unsigned n = 0, z = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_ring_dbg_n, sizeof(n));
  cudaMemcpyFromSymbol(out, g_ring_dbg, sizeof(double) * 64 * 12);
  cudaMemcpyToSymbol(g_ring_dbg_n, &z, sizeof(z));
  return (int)n;
}
#endif
