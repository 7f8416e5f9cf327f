"""Round-2 probe (torchrun, >= 2 GPUs): per-launch completion times of a slab-decomposed acoustic gradient.

Needs the timeline variant of the library:
  ADSEIS_NVCC_EXTRA=-DADSEIS_TIMELINE ADSEIS_LIB_SUFFIX=_tl python -c "import adseis_b200; adseis_b200.build()"
Run (emulates the per-GPU work of C4 on 8 GPUs with 2: NX = 1024 rows -> two 512-row slabs):
  ADSEIS_LIB_SUFFIX=_tl ADSEIS_TIMELINE=400 ADSEIS_TIMELINE_SKIP=600 PNX=1024 torchrun --nproc-per-node 2 ... this file
Prints per-kind statistics of the gaps between consecutive completions on each stream, and the sweep rates."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A
from adseis_b200 import parallel as P
from adseis_b200 import _lib

P.init_process_group("nccl")
import torch.distributed as dist
rank, world = dist.get_rank(), dist.get_world_size()
ctx = A.Context(int(os.environ.get("LOCAL_RANK", "0")))
NX, NY, NSTEP = int(os.environ.get("PNX", "1024")), int(os.environ.get("PNY", "4096")), int(os.environ.get("PT", "401"))
w = A.workloads.c4(nstep=NSTEP, nx=NX, ny=NY)
p, s = w["param"], w["shots"][0]
srcv = (A.Ricker(p, 20.0, 40.0) * 1e6).reshape(-1, 1)
dd = P.DomainDecomposedAcoustic(p, s["srci"], s["srcj"], s["rcvi"], s["rcvj"], ctx=ctx)
dd.set_model(w["model"]); dd.set_srcv(srcv); dd.forward()
obs = 0.5 * dd.rcvv()
dd.set_obs(obs)
for rep in range(3):
    ctx.sync(); dist.barrier(); t0 = time.perf_counter(); dd.gradient(); ctx.sync(); t1 = time.perf_counter()
tm, info = dd.plan.timings(), dd.plan.info()
loss = dd.loss()
g = dd.plan.grad_c_owned()[1].astype(np.longdouble)
gs = P.all_reduce_scalar(float(g.sum()), "sum", device=dd.dev)
if rank == 0:
    print("N=%d %dx%d nt=%d TB_SLAB=%s: gradient %.2f ms, %.2f us per step pair, %.1f Gcell-upd/s; fwd %.2f adj %.2f us/step;"
          " loss %.17g gsum %.17g" % (world, NX, NY, NSTEP, os.environ.get("ADSEIS_AC_TB_SLAB", "-"), (t1 - t0) * 1e3,
                                      (t1 - t0) * 1e6 / (NSTEP - 1), NX * NY * (NSTEP - 1) / (t1 - t0) / 1e9,
                                      tm["forward_ms"] * 1e3 / max(tm["forward_launches"], 1),
                                      tm["adjoint_ms"] * 1e3 / max(tm["adjoint_launches"], 1), loss, gs), flush=True)
if os.environ.get("ADSEIS_TIMELINE"):
    lib = _lib.load()
    path = "gpurun_out/timeline_%s_r%d.txt" % (os.environ.get("PTAG", "x"), rank)
    os.makedirs("gpurun_out", exist_ok=True)
    n = lib.adseis_debug_timeline_dump(path.encode())
    if rank == 0 and n > 0:
        rows = np.loadtxt(path).reshape(-1, 9)
        names = {0: "fwd full", 1: "fwd narrow frame", 2: "fwd wide frame", 3: "fwd box pair", 10: "adj full",
                 11: "adj narrow frame", 12: "adj wide frame", 13: "adj box pair"}
        print("  first 24 marks (kind, step, t_us):", " | ".join("%s %d %.1f" % (names[int(r[1])][:9], r[2], r[3]) for r in rows[:24]))
        for kinds, label in (((0,), "fwd one-step"), ((1, 2), "fwd frames (stream B)"), ((3,), "fwd box pairs (stream A)"),
                             ((10,), "adj one-step"), ((11, 12), "adj frames (stream B)"), ((13,), "adj box pairs (stream A)")):
            t = rows[np.isin(rows[:, 1].astype(int), kinds), 3]
            if len(t) > 2:
                d = np.diff(t)
                print("  %-28s n=%4d  gap between completions: median %.2f  mean %.2f  p90 %.2f us" %
                      (label, len(t), np.median(d), d.mean(), np.percentile(d, 90)))
        # device time stamps of the forward launches: [4] first CTA entry, [5] last CTA entry, [6] last edge CTA past its
        # halo wait, [7] last CTA done computing, [8] last edge CTA has published its rows
        for kinds, label in (((0,), "fwd one-step"), ((2,), "fwd wide frame"), ((1,), "fwd narrow frame"),
                             ((10,), "adj one-step"), ((12,), "adj wide frame"), ((11,), "adj narrow frame")):
            m = np.isin(rows[:, 1].astype(int), kinds) & (rows[:, 4] >= 0)
            r = rows[m]
            if len(r) > 4:
                med = lambda x: float(np.median(x))
                print("  %-18s entry spread %.2f | last entry -> past halo wait %.2f | -> compute done %.2f | -> published %.2f"
                      " | first entry -> published %.2f us" % (label, med(r[:, 5] - r[:, 4]), med(r[:, 6] - r[:, 5]),
                                                             med(r[:, 7] - r[:, 6]), med(r[:, 8] - r[:, 7]), med(r[:, 8] - r[:, 4])))
        fr = rows[np.isin(rows[:, 1].astype(int), (0, 1, 2, 10, 11, 12)) & (rows[:, 4] >= 0)]
        if len(fr) > 4:
            print("  previous launch published -> next launch's first CTA enters: median %.2f us" %
                  float(np.median(fr[1:, 4] - fr[:-1, 8])))
        # pair structure of the two-step path from device stamps alone (run with ADSEIS_TIMELINE_EVENTS=0): for each pair the
        # start / end of the box launch, the wide and the narrow frame launch, relative to the start of the pair's box launch
        for box, wide, nar, label in ((3, 2, 1, "fwd"), (13, 12, 11, "adj")):
            B_ = rows[(rows[:, 1] == box) & (rows[:, 4] >= 0)]
            out = []
            for b in B_[2:-2]:
                s0 = b[2]
                wd = rows[(rows[:, 1] == wide) & (rows[:, 2] == s0)]
                nr = rows[(rows[:, 1] == nar) & (rows[:, 2] == (s0 + 1 if box == 3 else s0 - 1))]
                if len(wd) == 1 and len(nr) == 1 and wd[0, 4] >= 0 and nr[0, 4] >= 0:
                    e = lambda r: max(r[7], r[8]) if r[8] >= 0 else r[7]
                    out.append([wd[0, 4] - b[4], e(wd[0]) - b[4], nr[0, 4] - b[4], e(nr[0]) - b[4], b[7] - b[4]])
            if len(out) > 4:
                o = np.median(np.array(out), axis=0)
                per = np.median(np.diff(B_[2:-2, 4]))
                print("  %s pair (us after the box launch's first CTA): wide %.1f..%.1f  narrow %.1f..%.1f  box ..%.1f   period %.1f" %
                      (label, o[0], o[1], o[2], o[3], o[4], per))
dd.close()
