"""Round-2 diagnostics: (a) per-call host overhead of one C3 shot through the plan API, (b) elastic C5 forward time per
step as a function of NSTEP / history window (round 1 measured 117 us/step on short runs; the full workload shows 281)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A

ctx = A.default_context()
what = sys.argv[1] if len(sys.argv) > 1 else "all"

def tic():
    ctx.sync(); return time.perf_counter()

if what in ("all", "shots"):
    w = A.workloads.c3(nstep=3000, shots=4)
    p = w["param"]
    sh = w["shots"]
    t = tic(); plan = A.AcousticPlan(p, sh[0]["srci"], sh[0]["srcj"], sh[0]["rcvi"], sh[0]["rcvj"], ctx=ctx); print("plan create %.1f ms" % ((tic() - t) * 1e3), plan.info())
    obs = np.zeros((p.NSTEP + 1, len(sh[0]["rcvi"])))
    for k in range(3):
        s = sh[k + 1]
        t0 = tic(); plan.set_points(s["srci"], s["srcj"], s["rcvi"], s["rcvj"]); t1 = tic()
        plan.set_model(w["model"]); t2 = tic()
        plan.set_srcv(s["srcv"]); t3 = tic()
        plan.set_obs(obs); t4 = tic()
        plan.gradient(); t5 = tic()
        L = plan.loss(); t6 = tic()
        g = plan.grad_c(); t7 = tic()
        r = plan.rcvv(); t8 = tic()
        tm = plan.timings()
        print("shot %d: set_points %.1f set_model %.1f set_srcv %.1f set_obs %.1f gradient %.1f (kernels fwd %.1f adj %.1f) loss %.1f grad_c %.1f rcvv %.1f ms" %
              (k, *(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)), tm["forward_ms"], tm["adjoint_ms"], *(1e3 * x for x in (t6 - t5, t7 - t6, t8 - t7))), flush=True)
    plan.close()

if what in ("all", "elastic"):
    for nstep, budget_slots in ((100, 0), (400, 0), (2000, 0), (2000, 8)):
        w = A.workloads.c5(nstep=nstep)
        p, s = w["param"], w["shots"][0]
        plan = A.ElasticPlan(p, s["srci"], s["srcj"], s["srctype"], s["rcvi"], s["rcvj"], s["rcvtype"], ctx=ctx,
                             hist_bytes_budget=budget_slots * 6 * 2004 * 2016 * 8)
        plan.set_model(*w["model"]); plan.set_srcv(s["srcv"]); plan.set_obs(np.zeros((len(s["rcvi"]), nstep + 1)))
        for rep in range(2):
            t0 = tic(); plan.forward(); t1 = tic()
        i = plan.info()
        t2 = tic(); plan.gradient(False); t3 = tic()
        print("C5 nstep %d budget %d: forward %.1f us/step (launches %d, slots %d segs %d) ; gradient(False) %.1f us/step total" %
              (nstep, budget_slots, (t1 - t0) * 1e6 / nstep, i["launches"], i["hist_slots"], i["segments"], (t3 - t2) * 1e6 / nstep), flush=True)
        plan.close()
