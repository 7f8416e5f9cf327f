"""Development probe: run the same acoustic forward sweep repeatedly on one plan and report any run-to-run difference
(a deterministic kernel must reproduce its traces and snapshots bit for bit)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A

NX, NY, NSTEP = int(os.environ.get("PNX", "4096")), int(os.environ.get("PNY", "4096")), int(os.environ.get("PT", "120"))
ctx = A.default_context()
p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=1e-3, vp_ref=2500.0)
srci, srcj = np.array([NX // 2, NX // 3]), np.array([NY // 2, 40])
rcvi = np.linspace(20, NX - 20, 256).astype(np.int64); rcvj = np.full(256, NY // 2 + 30)
plan = A.AcousticPlan(p, srci, srcj, rcvi, rcvj, ctx=ctx)
rng = np.random.default_rng(1)
z = np.linspace(0, 1, NY + 2)[None, :]
c = (1500.0 + 2000.0 * np.floor(z * 4) / 4) * (1 + 0.02 * rng.random((NX + 2, NY + 2)))
srcv = np.stack([A.Ricker(p, 30.0, 40.0, 1e6), A.Ricker(p, 25.0, 50.0, 5e5)], 1)
plan.set_model(c); plan.set_srcv(srcv)
GRAD = os.environ.get("PGRAD", "0") == "1"   # also repeat the reverse sweep (adjoint kernel) and compare the gradient
if GRAD:
    plan.forward()
    plan.set_obs(0.7 * plan.rcvv())
ref_r = ref_u = None
nbad = 0
for k in range(int(os.environ.get("REPS", "12"))):
    if GRAD:
        plan.gradient()
        r, u = plan.rcvv(), plan.grad_c()
    else:
        plan.forward()
        r, u = plan.rcvv(), plan.snapshot(NSTEP)
    if ref_r is None:
        ref_r, ref_u = r, u
        continue
    dr, du = (r != ref_r), (u != ref_u)
    if dr.any() or du.any():
        nbad += 1
        t, q = np.argwhere(dr)[0] if dr.any() else (-1, -1)
        ij = np.argwhere(du)
        print("run %d differs: %d trace entries (first at step %d receiver %d: %r vs %r), %d snapshot cells, rows %s cols %s"
              % (k, dr.sum(), t, q, r[t, q] if t >= 0 else None, ref_r[t, q] if t >= 0 else None, du.sum(),
                 (ij[:, 0].min(), ij[:, 0].max()) if len(ij) else None, (ij[:, 1].min(), ij[:, 1].max()) if len(ij) else None),
              flush=True)
print("lib%s PDL=%s %s: %d of %d repeat runs differ from the first" % (os.environ.get("ADSEIS_LIB_SUFFIX", ""), os.environ.get("ADSEIS_PDL", "1"), "gradient" if GRAD else "forward", nbad, k), flush=True)

import ctypes
_l = A._lib.load()
if hasattr(_l, "adseis_debug_ring"):
    buf = np.zeros((64, 12))
    n = _l.adseis_debug_ring(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    print("ring debug: %d mismatching stage reads" % n)
    for r in buf[:min(n, 64)]:
        what = {1: "w", 2: "wold", 4: "c2"}.get(int(r[6]), str(int(r[6])))
        tag = "OLD row (read before the copy landed)" if r[7] == r[9] else ("LATER row (overwritten early)" if r[7] == r[10] else "other")
        print("  cta %d it %d/%d thread %d row %d col %d array %s: stage %r expected %r -> %s" % (r[0], r[1], r[2], r[3], r[4], r[5], what, r[7], r[8], tag))
