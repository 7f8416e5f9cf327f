"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing (rendezvous, handle exchange,
all-reduce of losses / gradients); the per-step halo traffic itself never goes through torch or NCCL -- slab plans
store their edge rows straight into the neighbour's device memory over NVLink (CUDA IPC, see csrc/acoustic.cu).

Two modes, as in the reference:
  * shot parallelism  (src/Utils.jl:300-332, 574-600): shot k (1-based) runs on device k % n_gpu; per-device loss and
    gradient are summed -- here with one all-reduce.
  * domain decomposition (src/MPIAcoustic.jl): the reference splits into M x N square blocks over MPI ranks; here the
    padded grid is cut into 1-D slabs along i (rows are contiguous in memory), one slab per GPU.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .acoustic import AcousticPlan
from .elastic import ElasticPlan


# ------------------------------------------------------------------------------------------------------------
# partitioning helpers (pure host logic; exercised on CPU with the gloo backend in tests/test_parallel_cpu.py)
# ------------------------------------------------------------------------------------------------------------
def shot_assignment(nshots, world):
    """jobs of each rank, 0-based shot indices.  Reference rule (src/Utils.jl:326): device i (1-based) takes the
    1-based shots k with k % n_gpu == i-1."""
    return [[k - 1 for k in range(1, nshots + 1) if k % world == r] for r in range(world)]


def slab_partition(NX, world, rank):
    """(row0, row1): padded rows [row0, row1) owned by `rank` (adseis_slab_partition)."""
    s = _lib.SlabC()
    _lib.check(_lib.load().adseis_slab_partition(int(NX), int(world), int(rank), C.byref(s)))
    return int(s.row0), int(s.row1)


def slab_geometry(NX, NY, world, rank, halo=1):
    """Local array geometry of a slab, identical to the C side: dict(row0,row1,goff,Hl,own0,own1,ld)."""
    row0, row1 = slab_partition(NX, world, rank)
    lo = halo if rank > 0 else 0
    hi = halo if rank < world - 1 else 0
    ld = (NY + 2 + 15) // 16 * 16
    return dict(row0=row0, row1=row1, goff=row0 - lo, Hl=(row1 - row0) + lo + hi, own0=lo, own1=lo + (row1 - row0),
                ld=ld)


def owned_points(pi, row0, row1, mpi_convention):
    """Mask of the sources/receivers whose padded row lies in [row0,row1) (MPIAcoustic.jl:71-78, 98-104)."""
    gi = np.asarray(pi, dtype=np.int64) + (0 if mpi_convention else -1)
    return (gi >= row0) & (gi < row1)


def plan_window(NSTEP, slots):
    """History window W (snapshots) for a budget of `slots` snapshots incl. 4-plane checkpoints (mirrors
    plan_segments in csrc/acoustic.cu)."""
    if slots >= NSTEP + 1:
        return NSTEP + 1
    for W in range(int(slots), 5, -1):
        nseg = (NSTEP - 1 + (W - 2) - 1) // (W - 2)
        if W + 4 * (nseg - 1) <= slots:
            return W
    raise MemoryError("not enough device memory for a history window of 6 snapshots")


def _dist():
    import torch.distributed as dist
    return dist


def init_process_group(backend=None):
    """Rendezvous from the torchrun environment (RANK / WORLD_SIZE / MASTER_*); returns (rank, world, local_rank)."""
    import torch
    dist = _dist()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def all_reduce_sum_(t):
    """In-place sum over ranks of a torch tensor (CUDA under nccl, CPU under gloo)."""
    dist = _dist()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_scalar(x, op="sum", device=None):
    import torch
    dist = _dist()
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())


# ------------------------------------------------------------------------------------------------------------
# shot parallelism
# ------------------------------------------------------------------------------------------------------------
def _shot_context(n):
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    return rank, world, shot_assignment(n, world)[rank]


class ShotPlanCache:
    """One device-resident plan per GPU, re-pointed at every shot (adseis_acoustic_plan_set_points): the history
    window / checkpoints / adjoint planes are allocated once per inversion instead of once per shot.  Keep one
    instance across the L-BFGS iterations of an FWI run; close() it at the end."""

    def __init__(self, hist_bytes_budget=0):
        self.plan, self.key, self.budget = None, None, hist_bytes_budget

    def acoustic(self, param, src, rcv, ctx):
        key = ("ac", param.NX, param.NY, param.NSTEP, param.DELTAX, param.DELTAY, param.DELTAT, param.PropagatorKernel,
               param.mpi_convention, param.NPOINTS_PML, param.Rcoef, param.vp_ref)
        if self.plan is None or self.key != key:
            self.close()
            self.plan = AcousticPlan(param, src.srci, src.srcj, rcv.rcvi, rcv.rcvj, ctx=ctx,
                                     hist_bytes_budget=self.budget)
            self.key = key
        else:
            self.plan.set_points(src.srci, src.srcj, rcv.rcvi, rcv.rcvj)
        return self.plan

    def close(self):
        if self.plan is not None:
            self.plan.close()
        self.plan, self.key = None, None


def compute_loss_and_grads_GPU(param, srcs, rcvs, Rs, model, ctx=None, plan_cache=None, material_grads=True,
                               fetch_traces=True):
    """compute_loss_and_grads_GPU (src/Utils.jl:300-332): sum over shots of sum((rcvv-Rs)^2) and its gradient w.r.t. the
    model.  Acoustic (`srcs` of AcousticSource, `model` = velocity c) or elastic (`srcs` of ElasticSource, `model` =
    (rho, lambda, mu) -> gradient tuple in the same order).  Shots are dealt to the ranks of the current process group
    with the reference's round-robin rule (shot k, 1-based, on device k % n_gpu); the per-rank partial sums are
    all-reduced over NCCL, so every rank returns the full (loss, grad).  Works single-process too.
    `plan_cache`: a ShotPlanCache kept by the caller across calls (FWI iterations).  `fetch_traces=False` skips the
    device-to-host copy of every shot's simulated traces into rcvs[k].rcvv (an optimiser only needs loss and gradient)."""
    import torch
    from .structs import ElasticSource
    rank, world, jobs = _shot_context(len(srcs))
    ctx = ctx or _lib.default_context()
    on_gpu = torch.cuda.is_available()
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    elastic = len(srcs) > 0 and isinstance(srcs[0], ElasticSource)
    own_cache = plan_cache is None
    cache = plan_cache if plan_cache is not None else ShotPlanCache()
    shape = param.model_shape() if elastic else ((param.NX, param.NY) if param.mpi_convention else (param.NX + 2, param.NY + 2))
    ngrad = 3 if elastic else 1
    gsum = [torch.zeros(shape, dtype=torch.float64, device=dev) for _ in range(ngrad)]
    gtmp = torch.empty(shape, dtype=torch.float64, device=dev)
    loss = 0.0
    for k in jobs:
        src, rcv = srcs[k], rcvs[k]
        if elastic:
            plan = ElasticPlan(param, src.srci, src.srcj, src.srctype, rcv.rcvi, rcv.rcvj, rcv.rcvtype, ctx=ctx,
                               hist_bytes_budget=cache.budget)
            plan.set_model(*model)
        else:
            plan = cache.acoustic(param, src, rcv, ctx)
            plan.set_model(model)
        plan.set_srcv(src.srcv)
        plan.set_obs(Rs[k])
        if elastic:
            plan.gradient(material_grads)
            getters = (plan.grad_rho, plan.grad_lambda, plan.grad_mu) if material_grads else ()
        else:
            plan.gradient()
            getters = (plan.grad_c,)
        loss += plan.loss()
        for g, get in zip(gsum, getters):
            get(out=gtmp)          # device -> device on the context's stream, synchronised on return
            g += gtmp
            if on_gpu:             # torch's stream must be done with gtmp before it is overwritten
                torch.cuda.current_stream().synchronize()
        if fetch_traces:
            rcv.rcvv = plan.rcvv()
        if elastic:
            plan.close()
    if own_cache:
        cache.close()
    for g in gsum:
        all_reduce_sum_(g)         # NCCL over NVLink: the reference sums per-GPU gradients on the host
    loss = all_reduce_scalar(loss, "sum", device=dev)
    out = [g.cpu().numpy() for g in gsum]
    return loss, (tuple(out) if elastic else out[0])


def compute_forward_GPU(param, srcs, rcvs, model, ctx=None, plan_cache=None):
    """compute_forward_GPU (src/Utils.jl:574-600): simulated observations of every shot, shots dealt round-robin to the
    ranks; returns the list Rs (every rank gets all of it; rcvs[k].rcvv is filled too)."""
    from .structs import ElasticSource
    from .elastic import elastic_forward
    rank, world, jobs = _shot_context(len(srcs))
    ctx = ctx or _lib.default_context()
    elastic = len(srcs) > 0 and isinstance(srcs[0], ElasticSource)
    own_cache = plan_cache is None
    cache = plan_cache if plan_cache is not None else ShotPlanCache()
    mine = {}
    for k in jobs:
        src, rcv = srcs[k], rcvs[k]
        if elastic:
            mine[k] = elastic_forward(param, src, *model, rcv, ctx=ctx)[0]
        else:
            plan = cache.acoustic(param, src, rcv, ctx)
            plan.set_model(model)
            plan.set_srcv(src.srcv)
            plan.forward()
            mine[k] = plan.rcvv()
    if own_cache:
        cache.close()
    dist = _dist()
    parts = [mine]
    if dist.is_initialized() and world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    Rs = [None] * len(srcs)
    for part in parts:
        for k, v in part.items():
            Rs[k] = v
    for k, rcv in enumerate(rcvs):
        rcv.rcvv = Rs[k]
    return Rs


# ------------------------------------------------------------------------------------------------------------
# domain decomposition (acoustic)
# ------------------------------------------------------------------------------------------------------------
class DomainDecomposedAcoustic:
    """One slab of the acoustic solver per rank of the current process group.  The model is always the GLOBAL array
    (each rank uploads only its rows).  Sources and receivers are given globally; each rank keeps the ones whose row it
    owns (MPIAcoustic.jl:71-78, 98-104), so srcv / obs / traces travel between host and device only on the owning
    rank; results are reduced on request."""

    def __init__(self, param, srci, srcj, rcvi, rcvj, ctx=None, hist_slots=None):
        import torch
        dist = _dist()
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.param = param
        self.ctx = ctx or _lib.default_context()
        # PropagatorKernel = 0 (phi', psi' from the new wavefield, MPIAcoustic.jl:212-246) keeps two halo rows per neighbour
        k0 = param.PropagatorKernel == 0 and self.world > 1
        self.geo = slab_geometry(param.NX, param.NY, self.world, self.rank, halo=2 if k0 else 1)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        plane_bytes = self.geo["Hl"] * self.geo["ld"] * 8
        srci, srcj, rcvi, rcvj = (np.atleast_1d(np.asarray(x, dtype=np.int64)) for x in (srci, srcj, rcvi, rcvj))
        self.nsrc, self.nrcv = len(srci), len(rcvi)
        # the plan keeps the points it owns and ignores the rest; with PropagatorKernel = 0 it also needs the neighbours'
        # sources next to its rows (their injected part is removed from its c-gradient terms, csrc k_ac_k0_src_corr)
        self.smask = owned_points(srci, self.geo["row0"] - (1 if k0 else 0), self.geo["row1"] + (1 if k0 else 0),
                                  param.mpi_convention)
        self.rmask = owned_points(rcvi, self.geo["row0"], self.geo["row1"], param.mpi_convention)
        if self.world == 1:
            self.plan = AcousticPlan(param, srci, srcj, rcvi, rcvj, ctx=self.ctx)
            return
        nsrc, nrcv = int(self.smask.sum()), int(self.rmask.sum())
        if hist_slots is None:
            free_b, _ = self.ctx.mem_info()
            own_model_bytes = (self.geo["row1"] - self.geo["row0"]) * (param.NY + 2) * 8
            # adjoint accumulators, checkpoints of the C side, gradient buffers, torch / NCCL workspaces
            reserve = (16 * plane_bytes + 6 * own_model_bytes + (param.NX + 2) * (param.NY + 2) * 8 +
                       (4 * (param.NSTEP + 1) * nrcv + 2 * param.NSTEP * nsrc) * 8 + (6 << 30))
            slots = max(0, free_b - reserve) // plane_bytes
            W = plan_window(param.NSTEP, slots)
            hist_slots = int(all_reduce_scalar(W, "min", device=self.dev))   # every rank must use the same window
        self.hist_slots = hist_slots
        slab = (self.rank, self.world, self.geo["row0"], self.geo["row1"])
        self.plan = AcousticPlan(param, srci[self.smask], srcj[self.smask], rcvi[self.rmask], rcvj[self.rmask],
                                 ctx=self.ctx, slab=slab, hist_bytes_budget=hist_slots * plane_bytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.plan.ipc_export())
        lo = handles[self.rank - 1] if self.rank > 0 else None
        hi = handles[self.rank + 1] if self.rank < self.world - 1 else None
        self.plan.ipc_connect(lo, hi)
        dist.barrier()

    def _cols(self, a, mask):
        if self.world == 1:
            return a
        return np.ascontiguousarray(np.asarray(a)[:, mask])

    def set_model(self, c):
        self.plan.set_model(c)

    def set_srcv(self, srcv):
        self.plan.set_srcv(self._cols(srcv, self.smask))

    def set_obs(self, obs):
        self.plan.set_obs(self._cols(obs, self.rmask))

    def forward(self):
        self.plan.forward()

    def gradient(self):
        self.plan.gradient()

    def loss(self):
        return all_reduce_scalar(self.plan.loss(), "sum", device=self.dev)

    def _scatter_cols(self, a, mask, n):
        import torch
        if self.world == 1:
            return a
        full = np.zeros((a.shape[0], n))
        full[:, mask] = a
        return all_reduce_sum_(torch.from_numpy(full).to(self.dev)).cpu().numpy()   # every point has exactly one owner

    def rcvv(self):
        return self._scatter_cols(self.plan.rcvv(), self.rmask, self.nrcv)

    def grad_c(self, reduce=True):
        import torch
        shape = self.plan.model_shape
        t = torch.zeros(shape, dtype=torch.float64, device=self.dev)
        first, cnt = self.plan.owned_model_rows()
        if cnt > 0:
            self.plan.grad_c_owned(out=t[first:first + cnt])      # device -> device, own rows only
        if reduce:
            all_reduce_sum_(t)
        return t

    def grad_c_owned(self, out=None):
        """(first_row, rows): this rank's shard of the model gradient (no collective)."""
        return self.plan.grad_c_owned(out=out)

    def grad_srcv(self):
        return self._scatter_cols(self.plan.grad_srcv(), self.smask, self.nsrc)

    def close(self):
        self.plan.close()


# ------------------------------------------------------------------------------------------------------------
# domain decomposition (elastic; replaces src/MPIElastic.jl's M x N blocks + 18 mpi_halo_exchange2 per step)
# ------------------------------------------------------------------------------------------------------------
EL_HALO = 2


def elastic_slab_partition(param, world, rank):
    """(row0, row1): rows of the INTERNAL elastic array ((NX+2) padded rows for variant 0, NX + 2*2 ghost rows for
    variant 1) owned by `rank` (adseis_elastic_slab_partition)."""
    s = _lib.SlabC()
    pc = param.to_c()
    _lib.check(_lib.load().adseis_elastic_slab_partition(C.byref(pc), int(world), int(rank), C.byref(s)))
    return int(s.row0), int(s.row1)


def elastic_owned_points(param, pi, row0, row1):
    """Mask of sources/receivers whose internal row lies in [row0,row1): 1-based padded index (variant 0) or
    1-based global index (variant 1, MPIElastic.jl:85-86) -> 0-based internal row."""
    gi = np.asarray(pi, dtype=np.int64) + (-1 if param.variant == 0 else 1)
    return (gi >= row0) & (gi < row1)


class DomainDecomposedElastic:
    """One slab of the elastic solver per rank of the current process group.  Inputs are the GLOBAL arrays on every
    rank; results are reduced on request (receivers / sources are owned by exactly one slab)."""

    def __init__(self, param, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=None, hist_slots=None):
        import torch
        dist = _dist()
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.param = param
        self.ctx = ctx or _lib.default_context()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        if self.world == 1:
            self.plan = ElasticPlan(param, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=self.ctx)
            return
        self.row0, self.row1 = elastic_slab_partition(param, self.world, self.rank)
        # every rank must use the same history window: probe the slot size with a 1-slot plan-free estimate
        H, W = param.NX + (2 if param.variant == 0 else 4), param.NY + (2 if param.variant == 0 else 4)
        ld = (W + 15) // 16 * 16
        Hl = (self.row1 - self.row0) + (EL_HALO if self.rank > 0 else 0) + (EL_HALO if self.rank < self.world - 1 else 0)
        slot_bytes_max = int(all_reduce_scalar(6 * Hl * ld * 8, "max", device=self.dev))   # 5 planes + compact memories
        if hist_slots is None:
            free_b, _ = self.ctx.mem_info()
            reserve = 40 * Hl * ld * 8 + 6 * H * W * 8 + (6 << 30)
            slots = max(2, (free_b - reserve) // slot_bytes_max) if free_b > reserve else 2
            slots = min(slots, param.NSTEP + 1)
            hist_slots = int(all_reduce_scalar(slots, "min", device=self.dev))
        self.hist_slots = int(hist_slots)
        slab = (self.rank, self.world, self.row0, self.row1)
        # the C side sizes the window as budget // slot_bytes: ask for exactly hist_slots slots of MY slot size
        probe = ElasticPlan(param, [], [], [], [], [], [], ctx=self.ctx, slab=slab, hist_bytes_budget=1)
        my_slot_bytes = probe.info()["slot_doubles"] * 8
        probe.close()
        self.plan = ElasticPlan(param, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=self.ctx, slab=slab,
                                hist_bytes_budget=self.hist_slots * my_slot_bytes)
        assert self.plan.info()["hist_slots"] == self.hist_slots, (self.plan.info(), self.hist_slots)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.plan.ipc_export())
        lo = handles[self.rank - 1] if self.rank > 0 else None
        hi = handles[self.rank + 1] if self.rank < self.world - 1 else None
        self.plan.ipc_connect(lo, hi)
        dist.barrier()

    def set_model(self, rho, lam, mu):
        self.plan.set_model(rho, lam, mu)

    def set_srcv(self, srcv):
        self.plan.set_srcv(srcv)

    def set_obs(self, obs):
        self.plan.set_obs(obs)

    def forward(self):
        self.plan.forward()

    def gradient(self, material_grads=True):
        self.plan.gradient(material_grads)

    def _sum(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        return all_reduce_sum_(t).cpu().numpy()

    def loss(self):
        return all_reduce_scalar(self.plan.loss(), "sum", device=self.dev)

    def rcvv(self):
        return self._sum(self.plan.rcvv())          # every receiver is owned by exactly one slab, the others hold 0

    def grad_srcv(self):
        return self._sum(self.plan.grad_srcv())

    def grads(self, reduce=True):
        """(grad_rho, grad_lambda, grad_mu): own rows filled, the rest zero; summed over ranks when reduce."""
        out = (self.plan.grad_rho(), self.plan.grad_lambda(), self.plan.grad_mu())
        return tuple(self._sum(a) for a in out) if (reduce and self.world > 1) else out

    def close(self):
        self.plan.close()


def bench_domain_decomposed(A, w, args, rank, world, local_rank):
    """bench.py at N > 1: the C4 workload slab-partitioned over the ranks (strong scaling).  Returns the JSON dict
    (meaningful on rank 0) plus the context under "_ctx" for the extra workloads."""
    import time
    import torch
    import bench as B
    init_process_group("nccl")
    dist = _dist()
    ctx = A.Context(local_rank)
    p, sh = w["param"], w["shots"][0]
    dd = DomainDecomposedAcoustic(p, sh["srci"], sh["srcj"], sh["rcvi"], sh["rcvj"], ctx=ctx)
    nrcv = len(sh["rcvi"])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_c2, h_srcv = pin(w["model"]), pin(dd._cols(sh["srcv"], dd.smask))
    dd.set_model(w["model_obs"]); dd.set_srcv(sh["srcv"]); dd.forward()
    h_obs = pin(dd.plan.rcvv())                     # this rank's receivers only
    dd.plan.set_model(h_c2.numpy()); dd.plan.set_srcv(h_srcv.numpy()); dd.plan.set_obs(h_obs.numpy())
    for _ in range(args.warmup):
        dd.gradient()
    ctx.sync(); dist.barrier()
    clocks = B.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        dd.gradient()
    ms = ctx.timer_stop_ms()
    dist.barrier()
    ms = all_reduce_scalar(ms, "max", device=dd.dev)
    launches = ctx.launch_count() - l0
    tm, info = dd.plan.timings(), dd.plan.info()
    loss = dd.loss()
    sec = ms / 1e3 / args.steps
    cells = p.NX * p.NY * (p.NSTEP - 1)
    # e2e: host model rows / owned srcv / owned obs in, loss + this rank's gradient rows out, every step
    first, cnt = dd.plan.owned_model_rows()
    h_grad = torch.empty((cnt, dd.plan.model_shape[1]), dtype=torch.float64).pin_memory()
    ctx.sync(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dd.plan.set_model(h_c2.numpy()); dd.plan.set_srcv(h_srcv.numpy()); dd.plan.set_obs(h_obs.numpy())
        dd.gradient()
        loss_e2e = dd.loss()
        dd.plan.grad_c_owned(out=h_grad.numpy())
    ctx.sync(); dist.barrier()
    sec_e2e = all_reduce_scalar(time.perf_counter() - t0, "max", device=dd.dev) / args.steps
    clk = clocks.stop() if rank == 0 else None
    # gradient fingerprint (must not depend on N): sum and sum of squares over the shards
    g = h_grad.numpy().astype(np.longdouble)
    gsum = all_reduce_scalar(float(g.sum()), "sum", device=dd.dev)
    gsq = all_reduce_scalar(float((g * g).sum()), "sum", device=dd.dev)
    h2d_mine = (cnt * dd.plan.model_shape[1] + h_srcv.numel() + h_obs.numel()) * 8
    h2d = all_reduce_scalar(h2d_mine, "sum", device=dd.dev)
    d2h = all_reduce_scalar(h_grad.numel() * 8 + 8, "sum", device=dd.dev)
    peak, peak_src = B.measured_peaks()
    rows = dd.geo["row1"] - dd.geo["row0"]
    frac_rows = rows / float(p.NX + 2)
    ab = A.workloads.algorithmic_bytes(w)
    adj_us = tm["adjoint_ms"] * 1e3 / max(tm["adjoint_launches"], 1)
    fwd_us = (tm["forward_ms"] + tm["recompute_ms"]) * 1e3 / max(tm["forward_launches"] + tm["recompute_launches"], 1)
    roof = B.roof_entry("ac_adj_kernel (+ fused halo exchange, per GPU, rank 0)", ab["adjoint"] * frac_rows, adj_us, peak,
                        peak_src, share=tm["adjoint_ms"] / (ms / args.steps),
                        note="launch duration includes the in-kernel peer halo push / wait of every step")
    roof["other_kernels"] = dict(ac_fwd_kernel=B.roof_entry("ac_fwd_kernel (+ fused halo exchange)", ab["forward"] * frac_rows,
                                                            fwd_us, peak, peak_src))
    whole = (ab["forward"] + ab["adjoint"]) * (p.NSTEP - 1) / sec / 1e9 / world
    roof["whole_gradient"] = dict(achieved=whole, frac=whole / peak, unit="GB/s per GPU")
    cfg = dict(workload=w["name"], grid=[p.NX, p.NY], nstep=p.NSTEP, shots=1, dx=p.DELTAX, dt=p.DELTAT,
               npml=p.NPOINTS_PML, nrcv=nrcv, parallelism="slab domain decomposition x%d (NVLink peer halo rows)" % world,
               history_slots=info["hist_slots"], segments=info["segments"],
               recomputed_forward_steps=info["recomputed_steps"],
               l2_policy="working set far exceeds L2; no explicit flush")
    out = dict(metric=B.METRIC, value=cells / sec / 1e9, unit=B.UNIT, n_gpus=world, steps=args.steps,
               warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f64", data="synthetic", config=cfg, clocks=clk,
               e2e=dict(value=cells / sec_e2e / 1e9, unit=B.UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                        ms_per_step=sec_e2e * 1e3,
                        note="every rank moves only the model rows, source columns, trace columns and gradient rows it owns"),
               gpu_launches=launches * world, roofline=roof, loss=loss, loss_e2e=loss_e2e, grad_checksum=[gsum, gsq])
    dd.close()
    out["_ctx"] = ctx
    return out


def bench_elastic_domain_decomposed(A, w, steps, warmup, rank, world, ctx):
    """C5 at N > 1: elastic variant M, slab-decomposed; source-time-function gradient and material gradient."""
    import torch
    import bench as B
    dist = _dist()
    p, sh = w["param"], w["shots"][0]
    dd = DomainDecomposedElastic(p, sh["srci"], sh["srcj"], sh["srctype"], sh["rcvi"], sh["rcvj"], sh["rcvtype"], ctx=ctx)
    dd.set_model(*w["model_obs"]); dd.set_srcv(sh["srcv"]); dd.forward()
    obs = dd.rcvv()
    dd.set_model(*w["model"]); dd.set_obs(obs)

    def timed(fn):
        for _ in range(warmup):
            fn()
        ctx.sync(); dist.barrier()
        l0 = ctx.launch_count()
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop_ms() / steps
        dist.barrier()
        return all_reduce_scalar(ms, "max", device=dd.dev), (ctx.launch_count() - l0) // steps

    ms_f, _ = timed(dd.forward)
    ms_s, _ = timed(lambda: dd.gradient(False))
    ms_m, l_m = timed(lambda: dd.gradient(True))
    info = dd.plan.info()
    loss = dd.loss()
    gl = dd.plan.grad_lambda().astype(np.longdouble)
    gsum = all_reduce_scalar(float(gl.sum()), "sum", device=dd.dev)
    gsq = all_reduce_scalar(float((gl * gl).sum()), "sum", device=dd.dev)
    dd.close()
    peak, peak_src = B.measured_peaks()
    ab = A.workloads.algorithmic_bytes(w)
    n, replay = p.NSTEP, info["recomputed_steps"]
    cells = p.NX * p.NY * n
    fwd_us, adj_mat_us = ms_f * 1e3 / n, (ms_m - ms_f * (1 + replay / n)) * 1e3 / n
    roof = B.roof_entry("el_vel_adj<1> + el_sigma_adj<1> (+ fused halo pushes), per GPU", ab["adjoint"] / world, adj_mat_us,
                        peak, peak_src)
    roof["other_kernels"] = {"forward step": B.roof_entry("el_sigma_fwd + el_vel_fwd", ab["forward"] / world, fwd_us, peak, peak_src)}
    return dict(workload=w["name"], metric=B.METRIC, unit=B.UNIT, value=cells / (ms_m * 1e-3) / 1e9, ms_per_step=ms_m,
                value_forward_only=cells / (ms_f * 1e-3) / 1e9, value_source_gradient=cells / (ms_s * 1e-3) / 1e9,
                n_gpus=world, scaling="strong", gpu_launches=int(l_m) * world, roofline=roof, loss=loss,
                grad_checksum=[gsum, gsq],
                config=dict(grid=[p.NX, p.NY], nstep=n, variant="M", history_slots=info["hist_slots"],
                            segments=info["segments"], recomputed_forward_steps=replay,
                            parallelism="slab domain decomposition x%d" % world))
