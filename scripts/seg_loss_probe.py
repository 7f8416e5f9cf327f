import sys, os, numpy as np
sys.path.insert(0, "/root/repo")
import adseis_b200 as A
ctx = A.default_context()
NX=NY=int(os.environ.get("PN","1024")); NSTEP=int(os.environ.get("PT","1500"))
p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=0.05, Rcoef=0.2, vp_ref=1000.0, mpi_convention=True)
c2 = np.full((NX, NY), 1000.0); c2[NX//2-NX//8:NX//2+NX//8, NY//2-NY//8:NY//2+NY//8]=2000.0
srcv = (A.Ricker(p, 100.0, 500.0)*1e6).reshape(-1, 1)
rcvj = np.arange(20, NY - 19); rcvi = np.full(len(rcvj), NX // 5)
res=[]
for budget in (0,):
    plan = A.AcousticPlan(p, [NX // 5], [NY // 2], rcvi, rcvj, ctx=ctx, hist_bytes_budget=budget)
    plan.set_model(np.full((NX,NY),1100.0)); plan.set_srcv(srcv); plan.forward(); obs = plan.rcvv().copy()
    plan.set_model(c2); plan.set_obs(obs); plan.gradient()
    g = plan.grad_c()
    print(budget, plan.info(), plan.loss(), np.abs(obs).max(), np.abs(plan.rcvv()).max(), np.abs(g).max(), flush=True)
    res.append((plan.loss(), g.copy(), plan.rcvv().copy()))
    plan.close()
