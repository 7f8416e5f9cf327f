"""Development probe (torchrun): per-step device time of the slab-decomposed acoustic path.  PNX x PNY grid split over
the ranks, e.g. 1024 x 4096 on 2 GPUs reproduces the per-GPU slab (512 rows) of the C4 workload on 8 GPUs.
Tuning knobs are read by the library at load time: ADSEIS_PDL, ADSEIS_AC_RB."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A
from adseis_b200 import parallel

rank, world, local_rank = parallel.init_process_group("nccl")
ctx = A.Context(local_rank)
import torch.distributed as dist

NX, NY, NSTEP = int(os.environ.get("PNX", "1024")), int(os.environ.get("PNY", "4096")), int(os.environ.get("PT", "600"))
pa = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=0.05,
                                Rcoef=0.2, vp_ref=1000.0, mpi_convention=True)
c2 = np.full((NX, NY), 1000.0)
rj = np.arange(20, NY - 19); ri = np.full(len(rj), NX // 5)
da = parallel.DomainDecomposedAcoustic(pa, [NX // 5], [NY // 2], ri, rj, ctx=ctx)
da.set_model(c2); da.set_srcv(A.Ricker(pa, 100.0, 500.0).reshape(-1, 1)); da.set_obs(np.zeros((NSTEP + 1, len(rj))))
da.gradient(); ctx.sync(); dist.barrier()
best = None
for _ in range(3):
    ctx.timer_start(); da.gradient(); ms = ctx.timer_stop_ms(); dist.barrier()
    ms = parallel.all_reduce_scalar(ms, "max")
    tm = da.plan.timings()
    if best is None or ms < best[0]:
        best = (ms, tm)
if rank == 0:
    ms, tm = best
    print("acoustic DD x%d %dx%d nt=%d PDL=%s RB=%s: %.2f ms/gradient, %.1f us/step-pair, fwd %.2f us/launch, adj %.2f us/launch, info %s"
          % (world, NX, NY, NSTEP, os.environ.get("ADSEIS_PDL", "1"), os.environ.get("ADSEIS_AC_RB", "auto"), ms,
             ms * 1e3 / (NSTEP - 1), tm["forward_ms"] * 1e3 / max(tm["forward_launches"], 1),
             tm["adjoint_ms"] * 1e3 / max(tm["adjoint_launches"], 1), da.plan.info()), flush=True)
da.close()
dist.barrier(); dist.destroy_process_group()
