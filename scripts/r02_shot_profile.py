"""cProfile of compute_loss_and_grads_GPU on a few C3 shots (where do the ~200 ms per shot beyond the kernels go?)."""
import os, sys, time, cProfile, pstats
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A
from adseis_b200 import parallel
ctx = A.default_context()
w = A.workloads.c3(nstep=3000, shots=int(os.environ.get("NSHOT", "6")))
p = w["param"]
srcs = [A.AcousticSource(s["srci"], s["srcj"], s["srcv"]) for s in w["shots"]]
rcvs = [A.AcousticReceiver(s["rcvi"], s["rcvj"]) for s in w["shots"]]
cache = parallel.ShotPlanCache()
Rs = parallel.compute_forward_GPU(p, srcs, rcvs, w["model_obs"], ctx=ctx, plan_cache=cache)
parallel.compute_loss_and_grads_GPU(p, srcs, rcvs, Rs, w["model"], ctx=ctx, plan_cache=cache)
ctx.sync(); t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
L, g = parallel.compute_loss_and_grads_GPU(p, srcs, rcvs, Rs, w["model"], ctx=ctx, plan_cache=cache)
pr.disable(); ctx.sync()
print("per shot %.1f ms" % ((time.perf_counter() - t0) * 1e3 / len(srcs)))
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
