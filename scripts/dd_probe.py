"""Development probe (torchrun): device time of the slab-decomposed elastic (C5 grid) and acoustic paths per step."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A
from adseis_b200 import parallel

rank, world, local_rank = parallel.init_process_group("nccl")
ctx = A.Context(local_rank)
import torch.distributed as dist


def timed(fn, reps=2):
    fn(); ctx.sync(); dist.barrier()
    ts = []
    for _ in range(reps):
        ctx.timer_start(); fn(); ts.append(ctx.timer_stop_ms()); dist.barrier()
    return parallel.all_reduce_scalar(min(ts), "max")


NX = int(os.environ.get("PN", "2000")); NY = int(os.environ.get("PNY", str(NX))); NSTEP = int(os.environ.get("PT", "60"))
p = A.ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=1.0, DELTAY=1.0, DELTAT=5e-5, vp_ref=3300.0, variant=1)
rho = np.full((NX, NY), 2800.0); vp = np.full((NX, NY), 3300.0); vs = vp / 1.732
lam, mu, rho = A.compute_lame_parameters(vp, vs, rho)
srcv = A.Ricker(p, 15.0, 100.0, 1e6).reshape(-1, 1)
rcvi = np.arange(20, NX - 20); rcvj = np.full(len(rcvi), NY // 2 + 7); rcvt = np.zeros(len(rcvi), dtype=np.int64)
dd = parallel.DomainDecomposedElastic(p, [NX // 5], [NY // 2], [0], rcvi, rcvj, rcvt, ctx=ctx)
dd.set_model(rho, lam, mu); dd.set_srcv(srcv); dd.set_obs(np.zeros((len(rcvi), NSTEP + 1)))
cells = NX * NY * NSTEP
for name, fn in (("forward", dd.forward), ("grad srcv", lambda: dd.gradient(False)), ("grad mat", lambda: dd.gradient(True))):
    ms = timed(fn)
    if rank == 0:
        print("elastic DD x%d %dx%d nt=%d %-10s %8.2f ms  %7.1f us/step  %6.2f Gcell/s" %
              (world, NX, NY, NSTEP, name, ms, ms * 1e3 / NSTEP, cells / ms / 1e6), flush=True)
dd.close()

NXa = int(os.environ.get("PNA", "4096"))
if NXa == 0:
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
pa = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NXa, NY=NXa, NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=0.05, Rcoef=0.2,
                                vp_ref=1000.0, mpi_convention=True)
c2 = np.full((NXa, NXa), 1000.0)
rj = np.arange(20, NXa - 19); ri = np.full(len(rj), NXa // 5)
da = parallel.DomainDecomposedAcoustic(pa, [NXa // 5], [NXa // 2], ri, rj, ctx=ctx)
da.set_model(c2); da.set_srcv(A.Ricker(pa, 100.0, 500.0).reshape(-1, 1)); da.set_obs(np.zeros((NSTEP + 1, len(rj))))
for name, fn in (("forward", da.forward), ("gradient", da.gradient)):
    ms = timed(fn)
    if rank == 0:
        print("acoustic DD x%d %d^2 nt=%d %-10s %8.2f ms  %7.1f us/step" % (world, NXa, NSTEP, name, ms, ms * 1e3 / NSTEP), flush=True)
da.close()
dist.barrier(); dist.destroy_process_group()
