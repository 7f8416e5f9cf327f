"""Round-2 probe (torchrun, 2+ GPUs): elastic C5 slab decomposition, leg by leg (forward only, source gradient, material
gradient), us per step; ADSEIS_EL_LL=0/1 switches the halo protocol."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A
from adseis_b200 import parallel as P
P.init_process_group("nccl")
import torch.distributed as dist
rank, world = dist.get_rank(), dist.get_world_size()
ctx = A.Context(int(os.environ.get("LOCAL_RANK", "0")))
w = A.workloads.c5(nstep=int(os.environ.get("PT", "200")))
p, sh = w["param"], w["shots"][0]
slots = int(os.environ.get("PSLOTS", "0")) or None
dd = P.DomainDecomposedElastic(p, sh["srci"], sh["srcj"], sh["srctype"], sh["rcvi"], sh["rcvj"], sh["rcvtype"], ctx=ctx, hist_slots=slots)
dd.set_model(*w["model_obs"]); dd.set_srcv(sh["srcv"])
def leg(name, fn, reps=3):
    for _ in range(reps):
        ctx.sync(); dist.barrier(); t0 = time.perf_counter(); fn(); ctx.sync(); t1 = time.perf_counter()
    if rank == 0:
        print("EL_LL=%s x%d %-16s %.1f us/step" % (os.environ.get("ADSEIS_EL_LL", "1"), world, name, (t1 - t0) * 1e6 / p.NSTEP), flush=True)
leg("forward", dd.forward)
obs = dd.rcvv()
dd.set_model(*w["model"]); dd.set_obs(obs)
leg("source gradient", lambda: dd.gradient(False))
leg("material gradient", lambda: dd.gradient(True))
info = dd.plan.info()
gl = np.asarray(dd.plan.grad_lambda()).astype(np.longdouble)
gs = P.all_reduce_scalar(float(gl.sum()), "sum", device=dd.dev)
if rank == 0:
    print("   loss %.17g  sum(grad_lambda) %.17g  segments %d" % (dd.loss(), gs, info["segments"]), flush=True)
dd.close()
