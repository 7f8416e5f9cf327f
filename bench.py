#!/usr/bin/env python
"""bench.py -- benchmark of the FDTD forward+adjoint hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--extra c5,c2,c1,c3|none]

Metric: Gcell-updates/s of ONE full gradient evaluation (forward + exact adjoint)
        = NX*NY*nsteps*nshots / seconds / 1e9, nsteps = NSTEP-1 acoustic, NSTEP elastic            

HEADLINE line = workload C4 (BASELINE.json configs[3], the one the target is quoted on): 2-D acoustic 4096 x 4096,
NSTEP=5000, MPIAcoustic conventions.  A "step" is one gradient evaluation of that shot.  On one GPU the 672 GB
wavefield history does not fit in HBM, so the reverse sweep uses segment checkpointing with one bit-identical forward
recomputation per segment; the recomputed cell-updates are NOT counted in the metric (they are overhead) and are
reported in `config`.  At N > 1 the grid is slab-partitioned over the ranks (strong scaling).

`extra` (same JSON line) = the other BASELINE.json configurations as sub-records, each with its own value, roofline on
SURVEY 8(d)'s ALGORITHMIC bytes (acoustic 32 / 56 B, elastic 104 / 192 B per cell-update) and cpu_baseline:
  c5  elastic variant M 2000^2 nt=2000 (source-time-function AND material gradient; slab-decomposed at N > 1)
  c2  elastic variant S 500^2 nt=1000        c1  acoustic 401x133 nt=1000 (both PropagatorKernel schemes)
  c3  acoustic 2000x1000 nt=3000, 64 shots dealt round-robin to the GPUs + one NCCL all-reduce (weak per shot)

value    : inputs already resident in HBM, CUDA events on the launching stream, max over ranks.
e2e      : the same gradient through the host-buffer API (pinned host model/srcv/obs -> H2D -> gradient -> loss and
           gradient D2H), wall clock with a device synchronise on both sides.
roofline : dominant kernel, algorithmic bytes per launch / its mean launch duration measured with CUDA events inside
           the timed region, against MEASURED_PEAKS.json:hbm_gbs.
cpu_baseline / --impl reference : the reference's own C++ op bodies (oracle/_ref, built from /root/reference) on the
           host cores, on a bounded sample of the same workload; thread count = the cores this process may use
           (os.sched_getaffinity), never an inherited OMP_NUM_THREADS.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gcell-updates/s (fwd+adjoint)"
UNIT = "Gcell-updates/s"


def host_threads():
    """Cores this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers,
    which silently made round 1's reference arm single-threaded at N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        pmax = max(power)                       # "under load": samples drawing more than half of the maximum power
        loaded = [s for s, p in zip(sm, power) if p >= 0.5 * pmax] or sm
        return dict(sm_mhz=statistics.median(loaded), sm_max_mhz=max(smax), power_w_max=pmax, samples=len(sm),
                    reasons=sorted(reasons))


def measured_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full summary (profiles/), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))["kernels"]
        for k in d:
            if k == kernel or k.startswith(kernel + "<"):
                return d[k]["dram_bytes_per_launch"]
        return None
    except Exception:
        return None


def roof_entry(kernel, bytes_per_launch, us, peak, peak_src, share=None, traffic=None, **kw):
    ach = bytes_per_launch / (us * 1e-6) / 1e9 if us > 0 else 0.0
    d = dict(bound="hbm", kernel=kernel, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic,
             peak_source=peak_src, bytes_per_launch=bytes_per_launch, us_per_launch=us)
    if share is not None:
        d["share_of_step"] = share
    d.update(kw)
    return d


def checksum(a):
    """Order-independent fingerprint of a gradient array: (sum, sum of squares) accumulated in extended precision."""
    a = np.asarray(a, dtype=np.longdouble).ravel()
    return [float(a.sum()), float((a * a).sum())]


# ------------------------------------------------------------------------------------------------------------
# CPU reference arms (the only code here that touches oracle/)
# ------------------------------------------------------------------------------------------------------------
def _c4_cpu_inputs(w, nstep_s):
    from oracle import pyoracle as po
    p = w["param"]
    sig, tau = po.acoustic_pml(p.NX, p.NY, p.DELTAX, p.DELTAY, npml=p.NPOINTS_PML, Rcoef=p.Rcoef, vp_ref=p.vp_ref)
    return po, p, sig, tau, po.ricker(nstep_s, 100.0, 500.0).reshape(-1, 1) * 1e6


def cpu_acoustic_single_thread(w, nstep_s):
    """Single-process reference path on ONE thread (what TF executes for PropagatorKernel=1 on CPU, minus TF's own
    overhead): AcousticOneStepCpuForward/Backward + ScatterAddOps bodies (oracle/_ref); the plain-C port
    (oracle/liboracle.so) when oracle/_ref is absent."""
    po, p, sig, tau, srcv = _c4_cpu_inputs(w, nstep_s)
    sh = w["shots"][0]
    NX, NY = p.NX, p.NY
    which, kind = ("ref", "reference") if po.has_ref() else ("oracle", "port")
    off = 1 if p.mpi_convention else 0
    if p.mpi_convention:
        c = np.zeros((NX + 2, NY + 2)); c[1:-1, 1:-1] = np.sqrt(w["model"])
    else:
        c = w["model"]
    if srcv.shape[0] < nstep_s or w["name"].startswith("C1"):
        srcv = sh["srcv"][:nstep_s]
    a = (NX, NY, nstep_s, p.DELTAT, p.DELTAX, p.DELTAY, sig, tau, c)
    t0 = time.perf_counter()
    u, r = po.acoustic_forward(*a, sh["srci"] + off, sh["srcj"] + off, srcv, sh["rcvi"] + off, sh["rcvj"] + off, which=which)
    po.acoustic_misfit_grad(*a, sh["srci"] + off, sh["srcj"] + off, sh["rcvi"] + off, sh["rcvj"] + off, np.zeros_like(r), u,
                            which=which)
    dt = time.perf_counter() - t0
    return dict(value=NX * NY * (nstep_s - 1) / dt / 1e9, unit=UNIT, cores=1, kind=kind, seconds=dt,
                sample="%dx%d grid, %d of %d time steps, forward+adjoint, single-process custom-op path (1 thread)" %
                       (NX, NY, nstep_s - 1, p.NSTEP - 1))


def cpu_elastic_reference(w, n=384, nstep_s=4):
    """Elastic CPU arm: (a) the reference's op-graph path -- every x[idx] a GatherOps call, every update a
    ScatterAddOps / ScatterNdOps call, add_source, get_receive, differentiated with the ops' backward bodies
    (oracle/ref_graph.inc; what TF executes, minus TF's dispatch) -- on an n x n sub-grid of the workload for a few
    steps (per-cell cost does not depend on the grid size; the graph keeps every intermediate, ~2 KB per cell-step);
    (b) the fused plain-C port (oracle.c) on the same sample, quoted beside it as `port_value`."""
    from oracle import pyoracle as po
    p = w["param"]
    v = p.variant
    H, W = po.elastic_dims(v, n, n)
    kw = dict(npml=p.NPOINTS_PML, npower=p.NPOWER, kmax=p.K_MAX_PML, alpha_max=p.ALPHA_MAX_PML, Rcoef=p.Rcoef, vp_ref=p.vp_ref)
    ax, bx = po.elastic_cpml_1d(n, p.DELTAX, p.DELTAT, **kw)
    ay, by = po.elastic_cpml_1d(n, p.DELTAY, p.DELTAT, **kw)
    rho, lam, mu = (np.full((H, W), float(np.asarray(x).mean())) for x in w["model"])
    srcv = po.ricker(nstep_s, 2.0, 2.0, 1e6).reshape(-1, 1)
    nr = 64
    a = (v, n, n, nstep_s, p.DELTAT, p.DELTAX, p.DELTAY, ax, bx, ay, by, rho, lam, mu, [n // 5], [n // 2], [0], srcv,
         np.full(nr, n // 5 + 3), np.arange(n // 2 - nr // 2, n // 2 + nr // 2), np.zeros(nr, dtype=np.int64))
    obs = np.zeros((nr, nstep_s + 1))
    out = dict(unit=UNIT, cores=1)
    t0 = time.perf_counter(); po.elastic_misfit_grad(*a, obs); tp = time.perf_counter() - t0
    out["port_value"] = n * n * nstep_s / tp / 1e9
    if po.has_ref():
        t0 = time.perf_counter(); po.ref_elastic(*a, obs); tr = time.perf_counter() - t0
        out.update(value=n * n * nstep_s / tr / 1e9, kind="reference", seconds=tr)
    else:
        out.update(value=out["port_value"], kind="port", seconds=tp)
    out["sample"] = ("%dx%d sub-grid of the workload, %d time steps, forward + all gradients, variant %s: reference "
                     "gather/scatter op graph with the ops' own backward bodies, 1 thread (port_value: fused plain-C "
                     "restatement on the same sample)" % (n, n, nstep_s, "M" if v else "S"))
    return out


def reference_arm_c4(w, steps, warmup, nstep_s):
    """`--impl reference`: the reference's block-decomposed (MPI) path with ranks emulated by OpenMP threads on all host
    cores: MpiAcousticOneStepCpuForward/Backward bodies + halo copies (oracle/_ref)."""
    po, p, sig, tau, srcv = _c4_cpu_inputs(w, nstep_s)
    NX, NY, sh = p.NX, p.NY, w["shots"][0]
    if not po.has_ref():
        ts = []
        for k in range(warmup + steps):
            r = cpu_acoustic_single_thread(w, nstep_s)
            if k >= warmup:
                ts.append(r["seconds"])
        return sum(ts) / len(ts), dict(kind="port", cores=1, sample=r["sample"])
    threads = host_threads()
    kblk = 1
    while kblk * kblk < 2 * threads and NX % (2 * kblk) == 0 and NY % (2 * kblk) == 0 and NX // (2 * kblk) >= 64:
        kblk *= 2
    n = NX // kblk
    assert NX % n == 0 and NY % n == 0 and NX == NY, "reference arm needs a square grid divisible into blocks"
    obs = np.zeros((nstep_s + 1, len(sh["rcvi"])))
    ts = []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        u = po.ref_mpi_acoustic_forward(NX, NY, n, nstep_s, p.DELTAT, p.DELTAX, p.DELTAY, sig, tau, w["model"],
                                        sh["srci"], sh["srcj"], srcv, nthreads=threads)
        po.ref_mpi_acoustic_gradient(NX, NY, n, nstep_s, p.DELTAT, p.DELTAX, p.DELTAY, sig, tau, w["model"],
                                     sh["srci"], sh["srcj"], sh["rcvi"], sh["rcvj"], obs, u, nthreads=threads)
        if k >= warmup:
            ts.append(time.perf_counter() - t0)
        del u
    return sum(ts) / len(ts), dict(kind="reference", cores=threads,
                                   sample="%dx%d grid, %d of %d time steps per step, forward+adjoint, MPIAcoustic block "
                                          "decomposition %dx%d blocks of %d^2 emulated with %d OpenMP threads (no MPI "
                                          "runtime in the image)" % (NX, NY, nstep_s - 1, p.NSTEP - 1, kblk, kblk, n, threads))


# ------------------------------------------------------------------------------------------------------------
# GPU legs
# ------------------------------------------------------------------------------------------------------------
def _pin(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


def run_acoustic(A, ctx, w, steps, warmup, hist_slots=0, e2e=True, sampler=None):
    """One acoustic shot on one GPU: device-resident gradient timing, per-kernel CUDA-event spans, host-buffer e2e."""
    import torch
    p, sh = w["param"], w["shots"][0]
    pitch = (p.NY + 2 + 15) // 16 * 16
    plan = A.AcousticPlan(p, sh["srci"], sh["srcj"], sh["rcvi"], sh["rcvj"], ctx=ctx,
                          hist_bytes_budget=hist_slots * (p.NX + 2) * pitch * 8)
    nrcv = len(sh["rcvi"])
    h_c, h_srcv = _pin(w["model"]), _pin(sh["srcv"])
    plan.set_model(w["model_obs"]); plan.set_srcv(sh["srcv"]); plan.forward()      # observed data
    h_obs = torch.empty((p.NSTEP + 1, nrcv), dtype=torch.float64).pin_memory()
    plan.rcvv(out=h_obs.numpy())
    h_grad = torch.empty(plan.model_shape, dtype=torch.float64).pin_memory()
    plan.set_model(h_c.numpy()); plan.set_srcv(h_srcv.numpy()); plan.set_obs(h_obs.numpy())
    for _ in range(warmup):
        plan.gradient()
    ctx.sync()
    if sampler:
        sampler.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(steps):
        plan.gradient()
    ms = ctx.timer_stop_ms()
    launches = ctx.launch_count() - l0
    tm, info, loss = plan.timings(), plan.info(), plan.loss()
    sec = ms / 1e3 / steps
    cells = p.NX * p.NY * (p.NSTEP - 1)
    out = dict(value=cells / sec / 1e9, ms_per_step=sec * 1e3, launches=launches, loss=loss, info=info, tm=tm,
               step_ms=ms / steps)
    if e2e:
        h2d = (h_c.numel() + h_srcv.numel() + h_obs.numel()) * 8
        d2h = h_grad.numel() * 8 + 8
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            plan.set_model(h_c.numpy()); plan.set_srcv(h_srcv.numpy()); plan.set_obs(h_obs.numpy())
            plan.gradient()
            loss_e2e = plan.loss()                 # D2H (synchronises)
            plan.grad_c(out=h_grad.numpy())        # D2H
        ctx.sync()
        sec_e2e = (time.perf_counter() - t0) / steps
        out["e2e"] = dict(value=cells / sec_e2e / 1e9, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                          ms_per_step=sec_e2e * 1e3)
        out["loss_e2e"] = loss_e2e
        out["grad_checksum"] = checksum(h_grad.numpy())
    if sampler:
        out["clocks"] = sampler.stop()
    plan.close()
    return out


def acoustic_roofline(A, w, r):
    peak, peak_src = measured_peaks()
    ab = A.workloads.algorithmic_bytes(w)
    tm = r["tm"]
    adj_us = tm["adjoint_ms"] * 1e3 / max(tm["adjoint_launches"], 1)
    nf = tm["forward_launches"] + tm["recompute_launches"]
    fwd_us = (tm["forward_ms"] + tm["recompute_ms"]) * 1e3 / max(nf, 1)
    p = w["param"]
    # two steps per launch (temporal blocking) switch on automatically for boxes of >= 6 M cells (csrc/acoustic.cu)
    n = p.NPOINTS_PML + 3
    tb = os.environ.get("ADSEIS_AC_TB", "") != "0" and (os.environ.get("ADSEIS_AC_TB") == "1" or
                                                         (p.NX - 2 * n) * (p.NY - 2 * n) >= (6 << 20))
    half = lambda x: None if x is None else x / 2.0
    if tb:
        roof = roof_entry("ac_adj2_kernel (two adjoint steps per launch) + 2 frame-only ac_adj_kernel launches; figures "
                          "are PER TIME STEP", ab["adjoint"], adj_us, peak, peak_src, share=tm["adjoint_ms"] / r["step_ms"],
                          traffic=half(ncu_traffic("ac_adj2_kernel")),
                          note="achieved = SURVEY 8(d) algorithmic bytes of ONE step (56 B/cell) / time per step; the "
                               "pair kernel moves 36 B per cell-step (ncu traffic above, per step), so frac may exceed 1")
        roof["other_kernels"] = dict(ac_fwd2_kernel=roof_entry(
            "ac_fwd2_kernel (two forward steps per launch) + 2 frame-only ac_fwd_kernel launches, per time step",
            ab["forward"], fwd_us, peak, peak_src, share=(tm["forward_ms"] + tm["recompute_ms"]) / r["step_ms"],
            traffic=half(ncu_traffic("ac_fwd2_kernel"))))
    else:
        # small grids run a whole sweep in one cooperative launch (csrc: ac_*_persist_kernel): a handful of launches per gradient
        sweep = r["info"]["launches"] < p.NSTEP // 2
        ka, kf = ("ac_adj_persist_kernel (whole sweep in one launch; figures are PER TIME STEP)",
                  "ac_fwd_persist_kernel (whole sweep in one launch, per time step)") if sweep else ("ac_adj_kernel", "ac_fwd_kernel")
        roof = roof_entry(ka, ab["adjoint"], adj_us, peak, peak_src, share=tm["adjoint_ms"] / r["step_ms"],
                          traffic=None if sweep else ncu_traffic("ac_adj_kernel"))
        roof["other_kernels"] = dict(ac_fwd_kernel=roof_entry(kf, ab["forward"], fwd_us, peak, peak_src,
                                                              share=(tm["forward_ms"] + tm["recompute_ms"]) / r["step_ms"],
                                                              traffic=None if sweep else ncu_traffic("ac_fwd_kernel")))
    # whole gradient against the 8(d) roofline: one forward + one adjoint pass over the grid per counted step
    whole = (ab["forward"] + ab["adjoint"]) * (p.NSTEP - 1) / (r["step_ms"] * 1e-3) / 1e9
    roof["whole_gradient"] = dict(achieved=whole, frac=whole / peak, unit="GB/s",
                                  note="(32+56 B) x cells x counted steps / gradient time; replayed steps are overhead")
    return roof


def run_elastic(A, ctx, w, steps, warmup):
    """One elastic shot on one GPU.  Three timed legs: forward only, source-time-function gradient (no tape), full
    material gradient (tape + checkpoint replay); the adjoint kernels' times follow by difference."""
    import torch
    p, sh = w["param"], w["shots"][0]
    plan = A.ElasticPlan(p, sh["srci"], sh["srcj"], sh["srctype"], sh["rcvi"], sh["rcvj"], sh["rcvtype"], ctx=ctx)
    nrcv = len(sh["rcvi"])
    h_m = [_pin(x) for x in w["model"]]
    h_srcv = _pin(sh["srcv"])
    plan.set_model(*w["model_obs"]); plan.set_srcv(sh["srcv"]); plan.forward()
    h_obs = _pin(plan.rcvv())
    plan.set_model(*[x.numpy() for x in h_m]); plan.set_srcv(h_srcv.numpy()); plan.set_obs(h_obs.numpy())
    cells = p.NX * p.NY * p.NSTEP

    def timed(fn):
        for _ in range(warmup):
            fn()
        ctx.sync()
        l0 = ctx.launch_count()
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop_ms() / steps
        return ms, (ctx.launch_count() - l0) // steps

    ms_f, l_f = timed(plan.forward)
    ms_s, l_s = timed(lambda: plan.gradient(False))
    ms_m, l_m = timed(lambda: plan.gradient(True))
    info = plan.info()
    loss = plan.loss()
    g = [plan.grad_rho(), plan.grad_lambda(), plan.grad_mu()]
    # e2e of the material gradient: host model / srcv / obs in, loss + three gradient planes out
    h_g = [torch.empty(plan.model_shape, dtype=torch.float64).pin_memory() for _ in range(3)]
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        plan.set_model(*[x.numpy() for x in h_m]); plan.set_srcv(h_srcv.numpy()); plan.set_obs(h_obs.numpy())
        plan.gradient(True)
        plan.loss()
        plan.grad_rho(out=h_g[0].numpy()); plan.grad_lambda(out=h_g[1].numpy()); plan.grad_mu(out=h_g[2].numpy())
    ctx.sync()
    sec_e2e = (time.perf_counter() - t0) / steps
    plan.close()
    peak, peak_src = measured_peaks()
    ab = A.workloads.algorithmic_bytes(w)
    n = p.NSTEP
    replay = info["recomputed_steps"]
    fwd_us = ms_f * 1e3 / n
    adj_src_us = (ms_s - ms_f) * 1e3 / n
    adj_mat_us = (ms_m - ms_f * (1 + replay / n)) * 1e3 / n
    roof = roof_entry("el_vel_adj<1> + el_sigma_adj<1> (one adjoint step = 2 launches)", ab["adjoint"], adj_mat_us, peak,
                      peak_src, share=adj_mat_us * n / (ms_m * 1e3))
    roof["other_kernels"] = {
        "el_sigma_fwd + el_vel_fwd (one forward step = 2 launches)": roof_entry("forward step", ab["forward"], fwd_us, peak, peak_src),
        "el_vel_adj<0> + el_sigma_adj<0> (source-time-function adjoint step)": roof_entry(
            "adjoint step, no material gradient", ab["adjoint_source_only"], adj_src_us, peak, peak_src)}
    whole = (ab["forward"] + ab["adjoint"]) * n / (ms_m * 1e-3) / 1e9
    roof["whole_gradient"] = dict(achieved=whole, frac=whole / peak, unit="GB/s",
                                  note="(104+192 B) x cells x NSTEP / material-gradient time; replayed steps are overhead")
    return dict(workload=w["name"], metric=METRIC, unit=UNIT, value=cells / (ms_m * 1e-3) / 1e9, ms_per_step=ms_m,
                value_forward_only=cells / (ms_f * 1e-3) / 1e9,
                value_source_gradient=cells / (ms_s * 1e-3) / 1e9,
                e2e=dict(value=cells / sec_e2e / 1e9, unit=UNIT, ms_per_step=sec_e2e * 1e3,
                         h2d_bytes_per_step=(sum(x.numel() for x in h_m) + h_srcv.numel() + h_obs.numel()) * 8,
                         d2h_bytes_per_step=3 * h_g[0].numel() * 8 + 8),
                gpu_launches=l_m, roofline=roof, loss=loss, grad_checksum=[checksum(x) for x in g],
                config=dict(grid=[p.NX, p.NY], nstep=n, variant="M" if p.variant else "S", history_slots=info["hist_slots"],
                            segments=info["segments"], recomputed_forward_steps=replay, nrcv=nrcv))


def run_shots(A, ctx, w, steps, warmup, rank, world, dev):
    """C3: all shots of the workload, dealt round-robin to the ranks (src/Utils.jl:326), ONE plan per GPU re-pointed at
    every shot, one NCCL all-reduce of the gradient per evaluation.  Device-timed with the max over ranks."""
    import torch
    from adseis_b200 import parallel
    p = w["param"]
    srcs = [A.AcousticSource(s["srci"], s["srcj"], s["srcv"]) for s in w["shots"]]
    rcvs = [A.AcousticReceiver(s["rcvi"], s["rcvj"]) for s in w["shots"]]
    cache = parallel.ShotPlanCache()
    Rs = parallel.compute_forward_GPU(p, srcs, rcvs, w["model_obs"], ctx=ctx, plan_cache=cache)
    for _ in range(warmup):
        parallel.compute_loss_and_grads_GPU(p, srcs, rcvs, Rs, w["model"], ctx=ctx, plan_cache=cache, fetch_traces=False)
    ctx.sync()
    if world > 1:
        torch.distributed.barrier()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss, g = parallel.compute_loss_and_grads_GPU(p, srcs, rcvs, Rs, w["model"], ctx=ctx, plan_cache=cache, fetch_traces=False)
    ctx.sync()
    sec = (time.perf_counter() - t0) / steps
    sec = parallel.all_reduce_scalar(sec, "max", device=dev)
    launches = parallel.all_reduce_scalar(ctx.launch_count() - l0, "sum", device=dev) / steps
    tm, info = cache.plan.timings(), cache.plan.info()
    cache.close()
    cells = p.NX * p.NY * (p.NSTEP - 1) * len(srcs)
    peak, peak_src = measured_peaks()
    ab = A.workloads.algorithmic_bytes(w)
    adj_us = tm["adjoint_ms"] * 1e3 / max(tm["adjoint_launches"], 1)
    fwd_us = tm["forward_ms"] * 1e3 / max(tm["forward_launches"], 1)
    roof = roof_entry("ac_adj_kernel (last shot of rank 0)", ab["adjoint"], adj_us, peak, peak_src)
    roof["other_kernels"] = dict(ac_fwd_kernel=roof_entry("ac_fwd_kernel", ab["forward"], fwd_us, peak, peak_src))
    whole = (ab["forward"] + ab["adjoint"]) * (p.NSTEP - 1) * len(srcs) / sec / 1e9 / world
    roof["whole_gradient"] = dict(achieved=whole, frac=whole / peak, unit="GB/s per GPU")
    return dict(workload=w["name"], metric=METRIC, unit=UNIT, value=cells / sec / 1e9, ms_per_step=sec * 1e3,
                n_gpus=world, scaling="strong (fixed 64 shots; shots are independent: no data-path collective but the "
                                      "final gradient all-reduce)",
                e2e=dict(value=cells / sec / 1e9, unit=UNIT, ms_per_step=sec * 1e3,
                         note="compute_loss_and_grads_GPU takes host models / traces and returns host gradients: the "
                              "timed region IS the end-to-end call (per shot H2D of model+srcv+obs, D2H of the loss; gradient D2H once)",
                         h2d_bytes_per_step=int(sum((np.asarray(w["model"]).size + s["srcv"].size + Rs[k].size) * 8
                                                    for k, s in enumerate(w["shots"]))),
                         d2h_bytes_per_step=int(8 * len(srcs) + world * np.asarray(w["model"]).size * 8)),
                gpu_launches=int(launches), roofline=roof, loss=loss, grad_checksum=checksum(g),
                config=dict(grid=[p.NX, p.NY], nstep=p.NSTEP, shots=len(srcs), history_slots=info["hist_slots"],
                            segments=info["segments"], parallelism="shots round-robin over %d GPU(s) + NCCL all-reduce" % world))


def acoustic_sub_record(A, ctx, w, steps, warmup):
    r = run_acoustic(A, ctx, w, steps, warmup)
    p = w["param"]
    return dict(workload=w["name"], metric=METRIC, unit=UNIT, value=r["value"], ms_per_step=r["ms_per_step"], e2e=r["e2e"],
                gpu_launches=r["launches"] // steps, roofline=acoustic_roofline(A, w, r), loss=r["loss"],
                grad_checksum=r["grad_checksum"],
                config=dict(grid=[p.NX, p.NY], nstep=p.NSTEP, PropagatorKernel=p.PropagatorKernel,
                            history_slots=r["info"]["hist_slots"], segments=r["info"]["segments"]))


# ------------------------------------------------------------------------------------------------------------
# main
# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=4096)
    ap.add_argument("--nstep", type=int, default=5000)
    ap.add_argument("--extra", default="c5,c2,c1,c3", help="comma list of extra workloads (c1,c2,c3,c5) or 'none'")
    ap.add_argument("--extra-nstep", type=int, default=0, help="shorten the extra workloads (profiling only)")
    ap.add_argument("--extra-shots", type=int, default=0, help="shots of the C3 extra (0 = 64, the BASELINE config)")
    ap.add_argument("--cpu-steps", type=int, default=0, help="time steps of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--hist-slots", type=int, default=0,
                    help="force a history window of this many snapshots (profiling: reproduces the checkpoint/replay "
                         "mix of the full workload at a small --nstep); 0 = as many as fit")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    extras = [] if args.extra in ("none", "") else [x.strip() for x in args.extra.split(",") if x.strip()]

    # -------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        import adseis_b200 as A              # workload builders only; no GPU call on this arm
        w = A.workloads.c4(args.nstep, args.nx, args.ny)
        p = w["param"]
        nstep_s = args.cpu_steps or 16
        sec, desc = reference_arm_c4(w, args.steps, args.warmup, nstep_s)
        val = p.NX * p.NY * (nstep_s - 1) / sec / 1e9
        cfg = dict(workload=w["name"], grid=[p.NX, p.NY], nstep=p.NSTEP, shots=1, dx=p.DELTAX, dt=p.DELTAT,
                   npml=p.NPOINTS_PML, nrcv=len(w["shots"][0]["rcvi"]), parallelism="host cores only")
        out = dict(impl="reference", metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                   warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="strong",
                   vs_baseline=None, dtype="f64", data="synthetic", config=cfg,
                   cpu_baseline=dict(value=val, unit=UNIT, **desc),
                   e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(out))
        return

    # -------------------------------------------------------------------------------------------------------
    import torch
    import adseis_b200 as A
    W = A.workloads
    w = W.c4(args.nstep, args.nx, args.ny)
    p = w["param"]
    cfg = dict(workload=w["name"], grid=[p.NX, p.NY], nstep=p.NSTEP, shots=1, dx=p.DELTAX, dt=p.DELTAT,
               npml=p.NPOINTS_PML, nrcv=len(w["shots"][0]["rcvi"]),
               l2_policy="working set (>=1 GB per step) far exceeds the 126 MB L2; no explicit flush")

    def build_extra(name):
        kw = {}
        if args.extra_nstep:
            kw["nstep"] = args.extra_nstep
        if name == "c3" and args.extra_shots:
            kw["shots"] = args.extra_shots
        return W.BUILDERS[name](**kw)

    if world > 1:
        from adseis_b200 import parallel
        res = parallel.bench_domain_decomposed(A, w, args, rank, world, local_rank)
        ctx = res.pop("_ctx")
        dev = torch.device("cuda", local_rank)
        ex = {}
        for name in extras:
            try:
                if name == "c3":
                    ex["c3"] = run_shots(A, ctx, build_extra("c3"), 1, 1, rank, world, dev)
                elif name == "c5":
                    ex["c5"] = parallel.bench_elastic_domain_decomposed(A, build_extra("c5"), 1, 1, rank, world, ctx)
            except Exception as e:                      # an extra must never take the headline line down
                ex[name] = dict(error="%s: %s" % (type(e).__name__, e))
        res["extra"] = ex
        if rank == 0:
            print(json.dumps(res), flush=True)
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        return

    ctx = A.Context(local_rank)
    r = run_acoustic(A, ctx, w, args.steps, args.warmup, hist_slots=args.hist_slots, sampler=ClockSampler(local_rank))
    info = r["info"]
    cfg.update(parallelism="1 GPU", history_slots=info["hist_slots"], segments=info["segments"],
               recomputed_forward_steps=info["recomputed_steps"],
               note="recomputed forward steps are overhead and are not counted in the metric")
    out = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=1, steps=args.steps, warmup=args.warmup,
               ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
               data="synthetic", config=cfg, clocks=r["clocks"], e2e=r["e2e"], gpu_launches=r["launches"],
               roofline=acoustic_roofline(A, w, r), loss=r["loss"], loss_e2e=r["loss_e2e"],
               grad_checksum=r["grad_checksum"])
    # replay-free 1-GPU rate on the same grid (the whole tape resident): what an N-GPU run, which never replays,
    # should be compared with
    if p.NSTEP > 1200 and not args.hist_slots:
        try:
            ws = W.c4(min(1000, p.NSTEP), p.NX, p.NY)
            rs = run_acoustic(A, ctx, ws, 1, 1, e2e=False)
            out["replay_free_1gpu"] = dict(value=rs["value"], unit=UNIT, nstep=ws["param"].NSTEP,
                                           segments=rs["info"]["segments"],
                                           note="same grid, whole tape resident (no checkpoint replay)")
        except Exception as e:
            out["replay_free_1gpu"] = dict(error="%s: %s" % (type(e).__name__, e))
    ex = {}
    dev = torch.device("cuda", local_rank)
    for name in extras:
        try:
            we = build_extra(name)
            if name == "c3":
                ex[name] = run_shots(A, ctx, we, 1, 1, 0, 1, dev)
            elif we["kind"] == "elastic":
                ex[name] = run_elastic(A, ctx, we, 2, 1)
            elif name == "c1":
                ex[name] = acoustic_sub_record(A, ctx, we, 3, 2)
                ex["c1_kernel1"] = acoustic_sub_record(A, ctx, W.c1(we["param"].NSTEP, kernel=1), 3, 2)
            else:
                ex[name] = acoustic_sub_record(A, ctx, we, 3, 2)
        except Exception as e:
            ex[name] = dict(error="%s: %s" % (type(e).__name__, e))
    if not args.no_cpu:
        nstep_s = args.cpu_steps or 16
        out["cpu_baseline"] = cpu_acoustic_single_thread(w, nstep_s)
        for name in extras:
            if name not in ex or "error" in ex[name]:
                continue
            try:
                we = build_extra(name)
                if we["kind"] == "elastic":
                    ex[name]["cpu_baseline"] = cpu_elastic_reference(we)
                elif name == "c1":
                    ex[name]["cpu_baseline"] = cpu_acoustic_single_thread(W.c1(we["param"].NSTEP, kernel=1), min(600, we["param"].NSTEP))
                    ex[name]["cpu_baseline"]["note"] = "timed with the custom-op bodies (PropagatorKernel=1), the reference's fastest CPU path"
                elif name == "c3":
                    one = dict(we, shots=we["shots"][:1])
                    ex[name]["cpu_baseline"] = cpu_acoustic_single_thread(one, 10)
            except Exception as e:
                ex[name]["cpu_baseline"] = dict(error="%s: %s" % (type(e).__name__, e))
    out["extra"] = ex
    print(json.dumps(out))


if __name__ == "__main__":
    main()
