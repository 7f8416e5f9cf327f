"""Round-2 probe: one launch per step vs the whole-sweep (resident) kernel on small acoustic grids.
ADSEIS_AC_PERSIST=0/1 python scripts/r02_small_grid_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A
ctx = A.default_context()
NSTEP = int(os.environ.get("PT", "1000"))
for (NX, NY) in ((133, 401), (300, 300), (400, 600), (500, 1000), (1000, 1000)):
    for kernel in (1, 0):
        p = A.AcousticPropagatorParams(PropagatorKernel=kernel, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=1e-3,
                                       vp_ref=2500.0, NPOINTS_PML=12)
        c = 2500.0 * np.ones((NX + 2, NY + 2))
        srcv = (A.Ricker(p, 15.0, 0.1, 1e6)).reshape(-1, 1)
        rj = np.arange(20, NY - 18)
        plan = A.AcousticPlan(p, [NX // 2], [NY // 2], np.full(len(rj), NX // 3), rj, ctx=ctx)
        plan.set_model(c); plan.set_srcv(srcv); plan.forward()
        plan.set_obs(0.5 * plan.rcvv())
        for rep in range(3):
            ctx.sync(); t0 = time.perf_counter(); plan.gradient(); ctx.sync(); t1 = time.perf_counter()
        tm = plan.timings()
        print("PERSIST=%s %4dx%-4d kernel %d: %.2f us per step pair (fwd %.2f adj %.2f), %.2f Gcell-upd/s, launches %d" %
              (os.environ.get("ADSEIS_AC_PERSIST", "auto"), NX, NY, kernel, (t1 - t0) * 1e6 / (NSTEP - 1),
               tm["forward_ms"] * 1e3 / max(tm["forward_launches"], 1), tm["adjoint_ms"] * 1e3 / max(tm["adjoint_launches"], 1),
               NX * NY * (NSTEP - 1) / (t1 - t0) / 1e9, plan.info()["launches"]), flush=True)
        plan.close()
