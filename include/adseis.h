/* adseis.h -- C ABI of libadseis_b200.so: B200-native (sm_100a) 2-D acoustic / elastic staggered-grid FDTD with
 * PML/CPML and exact reverse-mode adjoints.  Drop-in for the custom-op library `libADSeismic` of
 * kailaix/ADSeismic.jl (deps/CustomOps/CMakeLists.txt:59-76) and for the time loops its Julia graph builders
 * wrap around those ops (src/Core.jl, src/MPIAcoustic.jl, src/MPIElastic.jl).  Citations are file:line in the
 * reference tree.  No torch / TensorFlow types appear here: plain pointers, sizes and POD structs only.
 *
 * Conventions (all the reference's own):
 *   - fp64 reals, int64 indices.
 *   - single-process ("S") grids are PADDED (NX+2)x(NY+2), row-major, flat index i*(NY+2)+j (Core.jl:673);
 *     srci/srcj/rcvi/rcvj are 1-BASED into that padded grid (Core.jl:600,727; AddSource.cpp:61).
 *   - "MPI convention" (mpi_convention=1): global UNPADDED NX x NY arrays, 1-based indices into the unpadded grid
 *     (MPIAcoustic.jl:59-111, 378, 427; MPIElastic.jl:66-134); acoustic `c` is already c^2 (MPIAcoustic.jl:336).
 *   - acoustic receivers: rcvv[(NSTEP+1)][nrcv] (Core.jl:728); elastic receivers: rcvv[nrcv][(NSTEP+1)]
 *     (GetReceive.cpp:19).  srcv is [rows>=NSTEP][nsrc] row-major.
 *   - every entry point returns 0 on success or a negative ADSEIS_E* code; adseis_last_error() gives the text.
 *   - a ctx is bound to ONE GPU and is single-caller (not re-entrant); the library never touches host threads.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns ADSEIS_ECUDA.
 */
#ifndef ADSEIS_H_
#define ADSEIS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADSEIS_OK 0
#define ADSEIS_EINVAL (-1)  /* bad argument */
#define ADSEIS_ECUDA (-2)   /* CUDA runtime error (text in adseis_last_error) */
#define ADSEIS_ENOMEM (-3)  /* device memory exhausted */
#define ADSEIS_ESTATE (-4)  /* call order (e.g. gradient before set_obs) */
#define ADSEIS_ECOMM (-5)   /* peer / IPC error */

typedef struct adseis_ctx adseis_ctx;

/* ---------------------------------------------------------------------------------------------
 * context
 * --------------------------------------------------------------------------------------------- */
int adseis_version(void);
const char* adseis_last_error(void);
int adseis_device_count(int* n);
/* device < 0: use the current device.  Creates one compute stream (non-blocking). */
int adseis_ctx_create(int device, adseis_ctx** out);
/* A context outlives its plans: destroying it while plans are alive (finalisers of a garbage-collected host language
 * run in any order) only marks it, and the last plan's destructor frees it. */
int adseis_ctx_destroy(adseis_ctx* ctx);
int adseis_ctx_sync(adseis_ctx* ctx);
/* cudaStream_t of the ctx (as void*), so a host framework can order its own work against ours. */
int adseis_ctx_stream(adseis_ctx* ctx, void** stream);
/* number of kernels launched by this ctx since creation (bench.py's gpu_launches). */
int adseis_ctx_launch_count(adseis_ctx* ctx, int64_t* n);
/* elapsed device time helpers: records events on the ctx stream. */
int adseis_ctx_timer_start(adseis_ctx* ctx);
int adseis_ctx_timer_stop_ms(adseis_ctx* ctx, double* ms);
/* memory info of the bound device */
int adseis_ctx_mem_info(adseis_ctx* ctx, size_t* free_bytes, size_t* total_bytes);

/* ---------------------------------------------------------------------------------------------
 * acoustic   (replaces: AcousticPropagatorParams src/Struct.jl:82-121, compute_PML_Params! src/Core.jl:622-655,
 *             AcousticPropagatorSolver src/Core.jl:562-620, SimulatedObservation! src/Core.jl:726-730,
 *             the AcousticOneStep and AcousticOneStepCpu ops under deps/CustomOps, and the loss/gradient of
 *             src/Utils.jl:300-332; MPI variants src/MPIAcoustic.jl:154-431)
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t NX, NY, NSTEP;
  double DELTAX, DELTAY, DELTAT;
  int32_t USE_PML_XMIN, USE_PML_XMAX, USE_PML_YMIN, USE_PML_YMAX;
  int64_t NPOINTS_PML;
  double Rcoef, vp_ref;
  int32_t mpi_convention;   /* 0: AcousticPropagatorSolver inputs; 1: MPIAcousticPropagatorSolver inputs */
  int32_t PropagatorKernel; /* 0: TF-op scheme, phi/psi from the NEW wavefield (Core.jl:528-549, MPIAcoustic.jl:212-246; slab plans
                                  keep two halo rows and exchange them after every step);
                               1: custom-op scheme, phi/psi from the OLD wavefield (AcousticOneStepCpu.h); 2 == 1 */
} adseis_acoustic_params;

/* Slab of a 1-D domain decomposition along i (rows).  Global padded rows [0, NX+2) are split into `nranks`
 * contiguous slabs; this rank owns padded rows [row0, row1).  A plain single-GPU plan uses {0,1,0,NX+2}. */
typedef struct {
  int32_t rank, nranks;
  int64_t row0, row1;
} adseis_slab;

typedef struct adseis_acoustic_plan adseis_acoustic_plan;

/* Compute the separable PML profiles the reference stores as Sigma_x/Sigma_y (Core.jl:622-655):
 * sigx[NX+2] (function of i), tauy[NY+2] (function of j).  Host-only helper (no GPU needed). */
int adseis_acoustic_pml_profiles(const adseis_acoustic_params* p, double* sigx, double* tauy);

/* Suggested slab bounds for `nranks` (balanced on interior rows). Host-only helper. */
int adseis_slab_partition(int64_t NX, int32_t nranks, int32_t rank, adseis_slab* out);

/* A plan owns all device state for one (params, source set, receiver set): model, PML profiles, wavefield
 * history / checkpoints, adjoint state, receiver traces, gradients.
 * hist_bytes_budget: device bytes the plan may use for wavefield history (0 = auto: most of the free memory).
 * When the full history (NSTEP+1 snapshots) does not fit, gradients use segment checkpointing with one
 * bit-identical forward recomputation per segment.
 * slab == NULL: single GPU.  Sources / receivers are given with GLOBAL indices; a slab plan keeps the ones it
 * owns (MPIAcoustic.jl:71-78, 98-104) and ignores the rest, so every rank may pass the full lists.  With
 * PropagatorKernel = 0 a slab plan must at least be given, besides its own sources, the sources on the row just
 * outside either end of its slab (their injected part is removed from c-gradient terms of its own cells);
 * srcv / grad_srcv keep one column per source passed, columns of sources owned elsewhere stay zero. */
int adseis_acoustic_plan_create(adseis_ctx* ctx, const adseis_acoustic_params* p, const adseis_slab* slab,
                                int64_t nsrc, const int64_t* srci, const int64_t* srcj, int64_t nrcv,
                                const int64_t* rcvi, const int64_t* rcvj, size_t hist_bytes_budget,
                                adseis_acoustic_plan** out);
int adseis_acoustic_plan_destroy(adseis_acoustic_plan* plan);
/* Replace the plan's sources and receivers (the next shot on the same grid and model): the device state is kept, so a
 * multi-shot gradient -- compute_loss_and_grads_GPU / compute_forward_GPU, src/Utils.jl:300-332, 574-600 -- runs on
 * one plan per GPU.  Invalidates srcv / obs / results: call set_srcv (and set_obs) again. */
int adseis_acoustic_plan_set_points(adseis_acoustic_plan* plan, int64_t nsrc, const int64_t* srci,
                                    const int64_t* srcj, int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj);

/* c: velocity, (NX+2)*(NY+2) [mpi_convention=0] or c^2, NX*NY [mpi_convention=1]; ALWAYS the global array
 * (a slab plan copies its rows).  on_device!=0: `c` is a device pointer on the plan's GPU. */
int adseis_acoustic_plan_set_model(adseis_acoustic_plan* plan, const double* c, int on_device);
/* srcv: [rows][nsrc] with rows >= NSTEP, global source order. */
int adseis_acoustic_plan_set_srcv(adseis_acoustic_plan* plan, const double* srcv, int64_t rows, int on_device);
/* obs: [(NSTEP+1)][nrcv], global receiver order. */
int adseis_acoustic_plan_set_obs(adseis_acoustic_plan* plan, const double* obs, int on_device);

/* Forward sweep (asynchronous on the ctx stream): fills the receiver traces. */
int adseis_acoustic_plan_forward(adseis_acoustic_plan* plan);
/* Forward + reverse sweep: loss = sum (rcvv-obs)^2, d loss/d c, d loss/d srcv (asynchronous). */
int adseis_acoustic_plan_gradient(adseis_acoustic_plan* plan);

#define ADSEIS_GET_RCVV 1      /* (NSTEP+1)*nrcv ; slab plans: entries of receivers owned elsewhere are 0 */
#define ADSEIS_GET_LOSS 2      /* 1 */
#define ADSEIS_GET_GRAD_C 3    /* same shape as the model passed to set_model; slab plans fill their rows, rest 0 */
#define ADSEIS_GET_GRAD_SRCV 4 /* NSTEP*nsrc ; slab plans: columns of sources owned elsewhere are 0 */
#define ADSEIS_GET_GRAD_C_OWNED 8 /* slab plans: only the model rows this slab owns, contiguous (the rows
                                   * [max(row0,1)-1, min(row1,NX+1)-1) of an NX x NY model under mpi_convention, rows
                                   * [row0,row1) of the padded model otherwise): what a sharded optimiser needs back */
/* Synchronises the ctx stream, then copies result `what` to dst (host, or device when to_device!=0). */
int adseis_acoustic_plan_get(adseis_acoustic_plan* plan, int what, double* dst, int to_device);
/* Copy wavefield snapshot `slot` (0..NSTEP) in the caller's layout ((NX+2)*(NY+2), or NX*NY under
 * mpi_convention); only valid after a forward()/gradient() whose history held that slot (returns ADSEIS_ESTATE
 * otherwise).  Single-GPU plans only. */
int adseis_acoustic_plan_get_snapshot(adseis_acoustic_plan* plan, int64_t slot, double* dst, int to_device);
/* plan facts: [0]=history slots resident, [1]=segments used by the last gradient(), [2]=kernel launches of the
 * last forward()/gradient(), [3]=local padded rows, [4]=pitch (doubles), [5]=recomputed forward steps */
int adseis_acoustic_plan_info(adseis_acoustic_plan* plan, int64_t info[8]);
/* Device time of the last forward()/gradient(), measured with CUDA events on the ctx stream around each run of
 * identical time-step kernels: out[0..2] = milliseconds in {forward sweep, forward recomputation, adjoint sweep},
 * out[3..5] = kernel launches in each.  Synchronises on the recorded events. */
int adseis_acoustic_plan_timings(adseis_acoustic_plan* plan, double out[6]);

/* ---- multi-GPU halo exchange over peer memory (NVLink): one process per GPU --------------------------------
 * Each slab plan exports a CUDA IPC handle of its device arena; the host framework (torch.distributed, MPI, ...)
 * all-gathers the 64-byte handles plus the 8-byte arena offsets and hands the neighbours' ones back.  After
 * that the time-step kernels store their edge rows straight into the neighbour's halo rows and signal with
 * per-step flags in peer memory -- no host round trip and no collective call per step. */
#define ADSEIS_IPC_HANDLE_BYTES 64
int adseis_acoustic_plan_ipc_export(adseis_acoustic_plan* plan, void* handle_out /*64 B*/);
/* lo/hi: handles of rank-1 / rank+1 (NULL at the physical edges). */
int adseis_acoustic_plan_ipc_connect(adseis_acoustic_plan* plan, const void* handle_lo, const void* handle_hi);

/* One-call host-buffer entry points (the Julia `AcousticPropagatorSolver` + `SimulatedObservation!` shims call
 * these): create plan, H2D, run, D2H, destroy.  u_hist_out may be NULL; else (NSTEP+1) snapshots in the caller's
 * layout. */
int adseis_acoustic_forward(adseis_ctx* ctx, const adseis_acoustic_params* p, const double* c, int64_t nsrc,
                            const int64_t* srci, const int64_t* srcj, const double* srcv, int64_t srcv_rows,
                            int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj, double* rcvv_out,
                            double* u_hist_out);
int adseis_acoustic_misfit_grad(adseis_ctx* ctx, const adseis_acoustic_params* p, const double* c, int64_t nsrc,
                                const int64_t* srci, const int64_t* srcj, const double* srcv, int64_t srcv_rows,
                                int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj, const double* obs,
                                double* loss_out, double* rcvv_out /*nullable*/, double* grad_c_out /*nullable*/,
                                double* grad_srcv_out /*nullable*/);

/* ---- op-level entry points: same argument lists as the reference op bodies, DEVICE pointers, dense
 * (NX+2)*(NY+2) arrays, asynchronous on `stream` (a cudaStream_t, may be NULL = ctx stream).
 * Replace AcousticOneStepForward / AcousticOneStepBackward (AcousticOneStep.h:9-48; CPU bodies
 * AcousticOneStepCpu.h:1-48, 51-125).  Backward overwrites its five outputs (the reference zero-fills then
 * accumulates, AcousticOneStepCpu.cpp:363-367). */
int adseis_op_acoustic_step_fwd(adseis_ctx* ctx, const double* w, const double* wold, const double* phi,
                                const double* psi, const double* sigma, const double* tau, const double* c,
                                double dt, double hx, double hy, int64_t NX, int64_t NY, double* u,
                                double* phiout, double* psiout, void* stream);
int adseis_op_acoustic_step_bwd(adseis_ctx* ctx, double* grad_w, double* grad_wold, double* grad_phi,
                                double* grad_psi, double* grad_c, const double* grad_u, const double* grad_phiout,
                                const double* grad_psiout, const double* w, const double* sigma, const double* tau,
                                const double* c, double dt, double hx, double hy, int64_t NX, int64_t NY,
                                void* stream);

/* ---------------------------------------------------------------------------------------------
 * elastic   (replaces: ElasticPropagatorParams src/Struct.jl:4-32, compute_PML_Params src/Core.jl:231-407,
 *            ElasticPropagatorSolver src/Core.jl:31-228, SimulatedObservation! src/Core.jl:701-712, the
 *            AddSource/GetReceive/Gather/ScatterAdd/ScatterNd ops, and tf.gradients through them;
 *            variant 1 = MPIElasticPropagatorSolver src/MPIElastic.jl:374-682 on the global grid)
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t NX, NY, NSTEP;
  double DELTAX, DELTAY, DELTAT;
  double f0, vp_ref;
  int32_t USE_PML_XMIN, USE_PML_XMAX, USE_PML_YMIN, USE_PML_YMAX;
  int64_t NPOINTS_PML;
  double NPOWER, K_MAX_PML, ALPHA_MAX_PML, Rcoef;
  int32_t variant; /* 0 = "S": src/Core.jl (padded (NX+2)x(NY+2), averaged materials);
                      1 = "M": src/MPIElastic.jl (global NX x NY, no averaging, every cell updated) */
  int32_t reserved;
} adseis_elastic_params;

typedef struct adseis_elastic_plan adseis_elastic_plan;

/* CPML coefficient rows (Core.jl:231-407): a[2*n], b[2*n] (row 0 integer grid, row 1 half grid) for one axis;
 * axis 0 = x (n=NX, h=DELTAX), 1 = y.  Host-only helper. K is ignored exactly as the reference does
 * (adbroadcast idx 3/4 returns its first argument, Core.jl:686-693); K_MAX_PML must be 1. */
int adseis_elastic_cpml_profiles(const adseis_elastic_params* p, int axis, double* a, double* b);

/* Suggested slab bounds over the rows of the INTERNAL elastic array (variant 0: the padded (NX+2) rows; variant 1:
 * NX rows + 2 ghost rows per side, MPIElastic.jl:175-189), balanced on the NX interior rows.  Host-only helper.
 * A slab plan keeps 2 halo rows per interior side (4th-order staggered stencils; the reference's
 * mpi_halo_exchange2, MPIElastic.jl:475-480); its boundaries must stay 2 rows clear of the x-CPML strips. */
int adseis_elastic_slab_partition(const adseis_elastic_params* p, int32_t nranks, int32_t rank, adseis_slab* out);

/* slab == NULL (or nranks == 1): single GPU.  Sources / receivers are given with GLOBAL indices; a slab plan keeps
 * the ones whose row it owns (MPIElastic.jl:71-86, 116-131). */
int adseis_elastic_plan_create(adseis_ctx* ctx, const adseis_elastic_params* p, const adseis_slab* slab,
                               int64_t nsrc, const int64_t* srci, const int64_t* srcj, const int64_t* srctype,
                               int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj, const int64_t* rcvtype,
                               size_t hist_bytes_budget, adseis_elastic_plan** out);
int adseis_elastic_plan_destroy(adseis_elastic_plan* plan);
/* rho, lambda, mu: (NX+2)*(NY+2) padded [variant 0] or NX*NY [variant 1]; global arrays. */
int adseis_elastic_plan_set_model(adseis_elastic_plan* plan, const double* rho, const double* lambda,
                                  const double* mu, int on_device);
int adseis_elastic_plan_set_srcv(adseis_elastic_plan* plan, const double* srcv, int64_t rows, int on_device);
/* obs: [nrcv][(NSTEP+1)] */
int adseis_elastic_plan_set_obs(adseis_elastic_plan* plan, const double* obs, int on_device);
int adseis_elastic_plan_forward(adseis_elastic_plan* plan);
/* want_material_grads==0: only d loss/d srcv (needs no forward history at all). */
int adseis_elastic_plan_gradient(adseis_elastic_plan* plan, int want_material_grads);
#define ADSEIS_GET_GRAD_RHO 5
#define ADSEIS_GET_GRAD_LAMBDA 6
#define ADSEIS_GET_GRAD_MU 7
/* what: ADSEIS_GET_RCVV (nrcv*(NSTEP+1)), _LOSS, _GRAD_SRCV (NSTEP*nsrc), _GRAD_RHO/_LAMBDA/_MU (model shape) */
int adseis_elastic_plan_get(adseis_elastic_plan* plan, int what, double* dst, int to_device);
/* field: 0 vx, 1 vy, 2 sxx, 3 syy, 4 sxy (post-injection values of `slot`), caller's layout. */
int adseis_elastic_plan_get_snapshot(adseis_elastic_plan* plan, int field, int64_t slot, double* dst,
                                     int to_device);
int adseis_elastic_plan_info(adseis_elastic_plan* plan, int64_t info[8]);
/* Slab plans (one process per GPU): same protocol as the acoustic one -- export the arena's IPC handle, all-gather,
 * connect to rank-1 / rank+1.  Afterwards every time-step kernel stores its two edge rows of the fields it produced
 * (sigma_xx, sigma_xy after the stress pass; v_x, v_y after the velocity pass; their adjoints in the reverse sweep)
 * straight into the neighbour's halo rows over NVLink.  Results: GET_RCVV / GET_GRAD_SRCV hold zeros for points
 * owned elsewhere, GET_GRAD_* hold this slab's rows (rest zero), GET_LOSS the partial sum -- reduce with a SUM. */
int adseis_elastic_plan_ipc_export(adseis_elastic_plan* plan, void* handle_out);
int adseis_elastic_plan_ipc_connect(adseis_elastic_plan* plan, const void* handle_lo, const void* handle_hi);

int adseis_elastic_forward(adseis_ctx* ctx, const adseis_elastic_params* p, const double* rho,
                           const double* lambda, const double* mu, int64_t nsrc, const int64_t* srci,
                           const int64_t* srcj, const int64_t* srctype, const double* srcv, int64_t srcv_rows,
                           int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj, const int64_t* rcvtype,
                           double* rcvv_out, double* hist_out /* nullable: 5*(NSTEP+1) snapshots */);
int adseis_elastic_misfit_grad(adseis_ctx* ctx, const adseis_elastic_params* p, const double* rho,
                               const double* lambda, const double* mu, int64_t nsrc, const int64_t* srci,
                               const int64_t* srcj, const int64_t* srctype, const double* srcv, int64_t srcv_rows,
                               int64_t nrcv, const int64_t* rcvi, const int64_t* rcvj, const int64_t* rcvtype,
                               const double* obs, double* loss_out, double* rcvv_out, double* grad_rho_out,
                               double* grad_lambda_out, double* grad_mu_out, double* grad_srcv_out);

/* ---- op-level: AddSource (SourceOps/AddSource.cpp:33-87, bwd :91-159) and GetReceive
 * (ReceiveOps/GetReceive.cpp:10-46, bwd :48-97) on DEVICE pointers, dense (NX+2)*(NY+2) fields. */
int adseis_op_add_source_fwd(adseis_ctx* ctx, double* vx_, double* vy_, double* sxx_, double* syy_, double* sxy_,
                             const double* vx, const double* vy, const double* sxx, const double* syy,
                             const double* sxy, const int64_t* srci, const int64_t* srcj, const double* srcv,
                             const int64_t* srctype, int64_t nsrc, int64_t NX, int64_t NY, void* stream);
int adseis_op_get_receive_fwd(adseis_ctx* ctx, double* out, const double* vx, const double* vy, const double* sxx,
                              const double* syy, const double* sxy, int64_t nt, const int64_t* rcvi,
                              const int64_t* rcvj, const int64_t* rcvtype, int64_t nrcv, int64_t NX, int64_t NY,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADSEIS_H_ */
