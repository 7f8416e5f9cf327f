"""Round-2 probe: elastic C5 forward sweep with the nvidia-smi clock sampler running (is the rise of the step time with
the age of the wavefield a clock / power effect?)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A
import bench as B
ctx = A.default_context()
nstep = int(os.environ.get("PT", "2000"))
w = A.workloads.c5(nstep=nstep)
p, s = w["param"], w["shots"][0]
plan = A.ElasticPlan(p, s["srci"], s["srcj"], s["srctype"], s["rcvi"], s["rcvj"], s["rcvtype"], ctx=ctx)
plan.set_model(*w["model"]); plan.set_srcv(s["srcv"])
plan.forward(); ctx.sync()
cs = B.ClockSampler(0); cs.start()
ctx.sync(); t0 = time.perf_counter(); plan.forward(); ctx.sync(); t1 = time.perf_counter()
print("C5 forward nstep %d: %.1f us/step" % (nstep, (t1 - t0) * 1e6 / nstep), cs.stop(), flush=True)
if os.environ.get("PGRAD"):
    r = plan.rcvv()
    plan.set_obs(0.5 * r)
    plan.gradient(True); ctx.sync()
    t0 = time.perf_counter(); plan.gradient(True); ctx.sync(); t1 = time.perf_counter()
    print("C5 material gradient nstep %d: %.1f us/step (fwd+adj), loss %.17g" % (nstep, (t1 - t0) * 1e6 / nstep, plan.loss()), flush=True)
