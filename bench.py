#!/usr/bin/env python
"""bench.py -- headline benchmark of the FDTD forward+adjoint hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric: Gcell-updates/s of ONE full gradient evaluation (forward + exact adjoint) =
        NX*NY*(NSTEP-1) / seconds / 1e9                                   
Workload ("C4", BASELINE.json configs[3], the one the target is quoted on): 2-D acoustic, 4096 x 4096 cells,
NSTEP=5000, PML on all sides, one Ricker source, 4058 receivers, MPIAcousticPropagatorSolver conventions
(examples/mpi_acoustic_optimized/MPI_forward.jl:11-38 scaled up).  A "step" is one gradient evaluation of that
shot.  On one GPU the 672 GB wavefield history does not fit in 180 GB of HBM, so the reverse sweep uses segment
checkpointing with one bit-identical forward recomputation per segment; the recomputed cell-updates are NOT
counted in the metric (they are overhead) and are reported in `config`.

value    : inputs already resident in HBM, CUDA events on the launching stream, max over ranks.
e2e      : the same gradient through the host-buffer API (pinned host model/srcv/obs -> H2D -> gradient -> loss
           and grad_c D2H), wall clock with a device synchronise on both sides.
roofline : dominant kernel (ac_adj_kernel), algorithmic bytes per launch / its mean launch duration measured with
           CUDA events inside the timed region, against MEASURED_PEAKS.json:hbm_gbs.
cpu_baseline / --impl reference : the reference's own C++ op bodies (oracle/_ref, built from /root/reference) on
           the host cores, on a bounded sample of the same workload.
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


# ------------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------------
def workload_c4(NX=4096, NY=4096, NSTEP=5000):
    """SURVEY.md 8d "C4": c^2 = 1000 with a 2000 square inclusion, Rcoef 0.2, source (NX/5, NY/2), Ricker(100,500),
    receivers j=20..NY-19 at i=NX/5; MPI convention (c given as c^2 on the unpadded grid, unpadded indices)."""
    w = dict(name="C4 acoustic %dx%d nt=%d (mpi_acoustic_optimized analogue)" % (NX, NY, NSTEP), NX=NX, NY=NY,
             NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=0.05, Rcoef=0.2, vp_ref=1000.0, NPOINTS_PML=12)
    c2 = np.full((NX, NY), 1000.0)
    cx, cy, wx, wy = NX // 2, NY // 2, NX // 8, NY // 8
    c2[cx - wx - 1:cx + wx, cy - wy - 1:cy + wy] = 2000.0
    w["c2"] = c2
    w["c2_background"] = np.full((NX, NY), 1100.0)  # "observed" data come from this model
    w["srci"] = np.array([NX // 5], dtype=np.int64)
    w["srcj"] = np.array([NY // 2], dtype=np.int64)
    w["rcvj"] = np.arange(20, NY - 18, dtype=np.int64)
    w["rcvi"] = np.full(len(w["rcvj"]), NX // 5, dtype=np.int64)
    return w


def n_pml_cells(w):
    n = w["NPOINTS_PML"] + 1
    return w["NX"] * w["NY"] - max(w["NX"] - 2 * n, 0) * max(w["NY"] - 2 * n, 0)


def algorithmic_bytes(w):
    """forward 32 B/cell (+32 in the PML frame), adjoint 56 B/cell (+32 in the frame)."""
    N, Np = w["NX"] * w["NY"], n_pml_cells(w)
    return dict(forward=32 * (N - Np) + 64 * Np, adjoint=56 * (N - Np) + 88 * Np)


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
        # "under load": samples drawing more than half of the maximum observed power
        pmax = max(power)
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
        for k in (kernel + "<1>", kernel):        # the step kernels are templates on PropagatorKernel (1 = custom-op scheme)
            if k in d:
                return d[k]["dram_bytes_per_launch"]
        return None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------------
# CPU reference arms
# ------------------------------------------------------------------------------------------------------------
def cpu_sample_single_thread(w, nstep_s):
    """The single-process reference path on ONE thread (what TF executes for PropagatorKernel=1 on CPU, minus TF's
    own overhead): AcousticOneStepCpuForward/Backward + ScatterAddOps bodies via oracle/_ref; falls back to the
    plain-C port (oracle/liboracle.so) when oracle/_ref is absent."""
    from oracle import pyoracle as po
    NX, NY = w["NX"], w["NY"]
    kind = "reference" if po.has_ref() else "port"
    which = "ref" if po.has_ref() else "oracle"
    sig, tau = po.acoustic_pml(NX, NY, w["DELTAX"], w["DELTAY"], npml=w["NPOINTS_PML"], Rcoef=w["Rcoef"],
                               vp_ref=w["vp_ref"])
    c = np.zeros((NX + 2, NY + 2))
    c[1:-1, 1:-1] = np.sqrt(w["c2"])
    srcv = po.ricker(nstep_s, 100.0, 500.0).reshape(-1, 1) * 1e6
    srci, srcj, rcvi, rcvj = w["srci"] + 1, w["srcj"] + 1, w["rcvi"] + 1, w["rcvj"] + 1  # padded 1-based
    t0 = time.perf_counter()
    u, r = po.acoustic_forward(NX, NY, nstep_s, w["DELTAT"], w["DELTAX"], w["DELTAY"], sig, tau, c, srci, srcj, srcv,
                               rcvi, rcvj, which=which)
    po.acoustic_misfit_grad(NX, NY, nstep_s, w["DELTAT"], w["DELTAX"], w["DELTAY"], sig, tau, c, srci, srcj, rcvi,
                            rcvj, np.zeros_like(r), u, which=which)
    dt = time.perf_counter() - t0
    cells = NX * NY * (nstep_s - 1)
    return dict(value=cells / dt / 1e9, unit=UNIT, cores=1, kind=kind, seconds=dt,
                sample="%dx%d grid, %d of %d time steps, forward+adjoint, single-process op path (1 thread)" %
                       (NX, NY, nstep_s - 1, w["NSTEP"] - 1))


def reference_arm(w, steps, warmup, nstep_s):
    """`--impl reference`: the reference's block-decomposed (MPI) path with ranks emulated by OpenMP threads on all
    host cores: MpiAcousticOneStepCpuForward/Backward bodies + halo copies (oracle/_ref)."""
    from oracle import pyoracle as po
    NX, NY = w["NX"], w["NY"]
    if not po.has_ref():
        # the plain-C port, single thread (oracle/_ref was not built where /root/reference exists)
        ts = []
        for k in range(warmup + steps):
            r = cpu_sample_single_thread(w, nstep_s)
            if k >= warmup:
                ts.append(r["seconds"])
        sec = sum(ts) / len(ts)
        return sec, dict(kind="port", cores=1, sample=r["sample"])
    threads = po.ref_threads()
    kblk = 1
    while kblk * kblk < 2 * threads and NX % (2 * kblk) == 0 and NY % (2 * kblk) == 0 and NX // (2 * kblk) >= 64:
        kblk *= 2
    n = NX // kblk
    assert NX % n == 0 and NY % n == 0 and NX == NY, "reference arm needs a square grid divisible into blocks"
    sig, tau = po.acoustic_pml(NX, NY, w["DELTAX"], w["DELTAY"], npml=w["NPOINTS_PML"], Rcoef=w["Rcoef"],
                               vp_ref=w["vp_ref"])
    srcv = po.ricker(nstep_s, 100.0, 500.0).reshape(-1, 1) * 1e6
    obs = np.zeros((nstep_s + 1, len(w["rcvi"])))
    ts = []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        u = po.ref_mpi_acoustic_forward(NX, NY, n, nstep_s, w["DELTAT"], w["DELTAX"], w["DELTAY"], sig, tau, w["c2"],
                                        w["srci"], w["srcj"], srcv, nthreads=threads)
        po.ref_mpi_acoustic_gradient(NX, NY, n, nstep_s, w["DELTAT"], w["DELTAX"], w["DELTAY"], sig, tau, w["c2"],
                                     w["srci"], w["srcj"], w["rcvi"], w["rcvj"], obs, u, nthreads=threads)
        if k >= warmup:
            ts.append(time.perf_counter() - t0)
        del u
    sec = sum(ts) / len(ts)
    return sec, dict(kind="reference", cores=threads,
                     sample="%dx%d grid, %d of %d time steps per step, forward+adjoint, MPIAcoustic block "
                            "decomposition %dx%d blocks of %d^2 emulated with %d OpenMP threads (no MPI runtime in "
                            "the image)" % (NX, NY, nstep_s - 1, w["NSTEP"] - 1, kblk, kblk, n, threads))


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
    ap.add_argument("--cpu-steps", type=int, default=0, help="time steps of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--hist-slots", type=int, default=0,
                    help="force a history window of this many snapshots (profiling: reproduces the checkpoint/replay "
                         "mix of the full workload at a small --nstep); 0 = as many as fit")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = workload_c4(args.nx, args.ny, args.nstep)
    cfg = dict(workload=w["name"], grid=[w["NX"], w["NY"]], nstep=w["NSTEP"], shots=1,
               dx=w["DELTAX"], dt=w["DELTAT"], npml=w["NPOINTS_PML"], nrcv=len(w["rcvi"]),
               l2_policy="working set (>=1 GB per step) far exceeds the 126 MB L2; no explicit flush")

    # -------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        nstep_s = args.cpu_steps or 16
        sec, desc = reference_arm(w, args.steps, args.warmup, nstep_s)
        cells = w["NX"] * w["NY"] * (nstep_s - 1)
        val = cells / sec / 1e9
        cfg["parallelism"] = "host cores only"
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
    if world > 1:
        from adseis_b200 import parallel
        res = parallel.bench_domain_decomposed(A, w, args, rank, world, local_rank)
        if rank == 0:
            print(json.dumps(res), flush=True)
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        return

    ctx = A.Context(local_rank)
    p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=w["NX"], NY=w["NY"], NSTEP=w["NSTEP"], DELTAX=w["DELTAX"], DELTAY=w["DELTAY"],
                                   DELTAT=w["DELTAT"], Rcoef=w["Rcoef"], vp_ref=w["vp_ref"],
                                   NPOINTS_PML=w["NPOINTS_PML"], mpi_convention=True)
    srcv_np = (A.Ricker(p, 100.0, 500.0) * 1e6).reshape(-1, 1)
    pitch = (w["NY"] + 2 + 15) // 16 * 16
    plan = A.AcousticPlan(p, w["srci"], w["srcj"], w["rcvi"], w["rcvj"], ctx=ctx,
                          hist_bytes_budget=args.hist_slots * (w["NX"] + 2) * pitch * 8)
    nrcv = len(w["rcvi"])

    # pinned host buffers (the e2e leg copies from / to these)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_c2, h_srcv = pin(w["c2"]), pin(srcv_np)
    # observed data = traces of the background model (device-resident forward, then kept on the host, pinned)
    plan.set_model(w["c2_background"]); plan.set_srcv(srcv_np); plan.forward()
    h_obs = torch.empty((p.NSTEP + 1, nrcv), dtype=torch.float64).pin_memory()
    plan.rcvv(out=h_obs.numpy())
    h_grad = torch.empty((w["NX"], w["NY"]), dtype=torch.float64).pin_memory()

    # ---- value: inputs resident in HBM ----------------------------------------------------------------
    plan.set_model(h_c2.numpy()); plan.set_srcv(h_srcv.numpy()); plan.set_obs(h_obs.numpy())
    for _ in range(args.warmup):
        plan.gradient()
    ctx.sync()
    clocks = ClockSampler(local_rank); clocks.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        plan.gradient()
    ms = ctx.timer_stop_ms()
    launches = ctx.launch_count() - l0
    tm = plan.timings()       # CUDA-event spans of the last timed gradient, per kernel family
    info = plan.info()
    loss = plan.loss()
    sec = ms / 1e3 / args.steps
    cells = w["NX"] * w["NY"] * (w["NSTEP"] - 1)
    value = cells / sec / 1e9

    # ---- e2e: host buffers in, host results out, every step --------------------------------------------
    h2d = h_c2.numel() * 8 + h_srcv.numel() * 8 + h_obs.numel() * 8
    d2h = h_grad.numel() * 8 + 8
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.set_model(h_c2.numpy()); plan.set_srcv(h_srcv.numpy()); plan.set_obs(h_obs.numpy())
        plan.gradient()
        loss_e2e = plan.loss()                 # D2H (synchronises)
        plan.grad_c(out=h_grad.numpy())        # D2H
    ctx.sync()
    sec_e2e = (time.perf_counter() - t0) / args.steps
    clk = clocks.stop()

    # ---- roofline of the dominant kernel ----------------------------------------------------------------
    peak, peak_src = measured_peaks()
    ab = algorithmic_bytes(w)
    adj_us = tm["adjoint_ms"] * 1e3 / max(tm["adjoint_launches"], 1)
    fwd_us = (tm["forward_ms"] + tm["recompute_ms"]) * 1e3 / max(tm["forward_launches"] + tm["recompute_launches"], 1)
    achieved = ab["adjoint"] / (adj_us * 1e-6) / 1e9
    roof = dict(bound="hbm", kernel="ac_adj_kernel", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                traffic=ncu_traffic("ac_adj_kernel"), peak_source=peak_src, bytes_per_launch=ab["adjoint"],
                us_per_launch=adj_us, share_of_step=tm["adjoint_ms"] / (ms / args.steps),
                other_kernels=dict(ac_fwd_kernel=dict(achieved=ab["forward"] / (fwd_us * 1e-6) / 1e9,
                                                      frac=ab["forward"] / (fwd_us * 1e-6) / 1e9 / peak,
                                                      bytes_per_launch=ab["forward"], us_per_launch=fwd_us,
                                                      traffic=ncu_traffic("ac_fwd_kernel"),
                                                      share_of_step=(tm["forward_ms"] + tm["recompute_ms"]) /
                                                                    (ms / args.steps))))
    cfg.update(parallelism="1 GPU", history_slots=info["hist_slots"], segments=info["segments"],
               recomputed_forward_steps=info["recomputed_steps"],
               note="recomputed forward steps are overhead and are not counted in the metric")

    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=1, steps=args.steps, warmup=args.warmup,
               ms_per_step=sec * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
               data="synthetic", config=cfg, clocks=clk,
               e2e=dict(value=cells / sec_e2e / 1e9, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                        ms_per_step=sec_e2e * 1e3),
               gpu_launches=launches, roofline=roof, loss=loss, loss_e2e=loss_e2e)
    plan.close()
    if not args.no_cpu:
        nstep_s = args.cpu_steps or 16
        out["cpu_baseline"] = cpu_sample_single_thread(w, nstep_s)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
