"""Round-2 probe: forward sweep time per step at 4096^2 with and without temporal blocking (ADSEIS_AC_TB), and the
bitwise equality of the two paths (traces, last snapshot, gradient)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A
ctx = A.default_context()
NX = NY = int(os.environ.get("PN", "4096")); NSTEP = int(os.environ.get("PT", "201"))
w = A.workloads.c4(nstep=NSTEP, nx=NX, ny=NY)
p, s = w["param"], w["shots"][0]
srcv = (A.Ricker(p, 20.0, 40.0) * 1e6).reshape(-1, 1)     # a wavelet that lives inside the short run
plan = A.AcousticPlan(p, s["srci"], s["srcj"], s["rcvi"], s["rcvj"], ctx=ctx)
plan.set_model(w["model"]); plan.set_srcv(srcv)
for rep in range(3):
    ctx.sync(); t0 = time.perf_counter(); plan.forward(); ctx.sync(); t1 = time.perf_counter()
tm = plan.timings()
r, u = plan.rcvv(), plan.snapshot(NSTEP)
plan.set_obs(0.5 * r)
for rep in range(2):
    plan.gradient()
tg = plan.timings()
print("TB=%s RB2=%s: forward %.2f us/step (wall %.2f), adjoint %.2f us/step, launches %d, |rcvv|max %.3e" %
      (os.environ.get("ADSEIS_AC_TB", "1"), os.environ.get("ADSEIS_AC_RB2", "-"), tm["forward_ms"] * 1e3 / tm["forward_launches"], (t1 - t0) * 1e6 / (NSTEP - 1),
       tg["adjoint_ms"] * 1e3 / tg["adjoint_launches"], plan.info()["launches"], np.abs(r).max()), flush=True)
out = os.environ.get("POUT")
if out:
    np.savez(out, r=r, u=u, g=plan.grad_c(), L=plan.loss())
cmp_ = os.environ.get("PCMP")
if cmp_:
    d = np.load(cmp_)
    print("   vs %s: traces equal %s, snapshot equal %s, grad equal %s, loss equal %s" %
          (cmp_, np.array_equal(d["r"], r), np.array_equal(d["u"], u), np.array_equal(d["g"], plan.grad_c()), float(d["L"]) == plan.loss()), flush=True)
