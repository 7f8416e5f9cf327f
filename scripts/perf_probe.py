"""Quick device-side timing of the acoustic kernels (not the bench contract; a development probe)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A

def run(NX, NY, NSTEP, reps=3, budget=0):
    ctx = A.default_context()
    p = A.AcousticPropagatorParams(PropagatorKernel=int(os.environ.get('PK', '1')), NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=0.05, Rcoef=0.2,
                                   vp_ref=1000.0, mpi_convention=True)
    c2 = np.full((NX, NY), 1000.0); c2[NX//2-NX//8:NX//2+NX//8, NY//2-NY//8:NY//2+NY//8] = 2000.0
    srcv = A.Ricker(p, 100.0, 500.0).reshape(-1, 1)
    rcvj = np.arange(20, NY - 19); rcvi = np.full(len(rcvj), NX // 5)
    plan = A.AcousticPlan(p, [NX // 5], [NY // 2], rcvi, rcvj, ctx=ctx, hist_bytes_budget=budget)
    plan.set_model(c2); plan.set_srcv(srcv); plan.set_obs(np.zeros((NSTEP + 1, len(rcvj))))
    cells = NX * NY * (NSTEP - 1)
    for name, fn in (("forward", plan.forward), ("gradient", plan.gradient)):
        fn(); ctx.sync()
        ts = []
        for _ in range(reps):
            ctx.timer_start(); fn(); ts.append(ctx.timer_stop_ms())
        t = min(ts)
        info = plan.info()
        bytes_ = cells * (32 if name == "forward" else 88 + 32 * (info["recomputed_steps"] / max(NSTEP - 1, 1)))
        tm = plan.timings()
        per = lambda k: tm[k + "_ms"] * 1e3 / max(tm[k + "_launches"], 1)
        print("%dx%d nt=%d %-8s %8.2f ms  %7.2f Gcell/s  %7.1f GB/s(alg) | us/launch fwd %.1f adj %.1f | slots %d segs %d" %
              (NX, NY, NSTEP, name, t, cells / t / 1e6, bytes_ / t / 1e6, per("forward"), per("adjoint"),
               info["hist_slots"], info["segments"]), flush=True)
    plan.close()

if __name__ == "__main__" and not any(a.startswith("--elastic") for a in sys.argv):
    if "--quick" in sys.argv:
        print("variant", os.environ.get("ADSEIS_LIB_SUFFIX", "(default)"))
        run(4096, 4096, 60)
        sys.exit(0)
    if "--rb" in sys.argv:
        for nx, ny, nt, rbs in ((512, 4096, 200, (6, 8, 10, 14, 20, 28)), (4096, 4096, 40, (24, 28, 37, 56, 74))):
            for rb in rbs:
                os.environ["ADSEIS_AC_RB"] = str(rb)
                print("rb", rb, end=" ")
                run(nx, ny, nt, reps=2)
        sys.exit(0)
    if "--pk" in sys.argv:
        for pk in ("1", "0"):
            os.environ["PK"] = pk
            print("PropagatorKernel", pk)
            run(4096, 4096, 60, reps=2)
            run(401, 133, 600, reps=2)
        sys.exit(0)
    if "--one" in sys.argv:
        run(4096, 4096, 24, reps=1)
        sys.exit(0)
    run(4096, 4096, 60)
    run(4096, 4096, 120, budget=40 * 4098 * 4112 * 8)
    run(2000, 1000, 200)
    run(401, 133, 1000)


def run_elastic(NX, NY, NSTEP, variant, reps=2, mat=True):
    ctx = A.default_context()
    p = A.ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=1.0, DELTAY=1.0, DELTAT=1e-4, vp_ref=3300.0,
                                  variant=variant)
    sh = p.model_shape()
    rho = np.full(sh, 2800.0); vp = np.full(sh, 3000.0); vp[:, sh[1] // 2:] = 3300.0; vs = vp / 1.732
    lam, mu, rho = A.compute_lame_parameters(vp, vs, rho)
    srcv = A.Ricker(p, 15.0, 100.0, 1e6).reshape(-1, 1)
    rcvi = np.arange(20, NX - 20); rcvj = np.full(len(rcvi), 20); rcvt = np.arange(len(rcvi)) % 2
    plan = A.ElasticPlan(p, [NX // 2], [NY // 2], [0], rcvi, rcvj, rcvt, ctx=ctx)
    plan.set_model(rho, lam, mu); plan.set_srcv(srcv); plan.set_obs(np.zeros((len(rcvi), NSTEP + 1)))
    cells = NX * NY * NSTEP
    for name, fn, byt in (("forward", plan.forward, 144), ("gradient", lambda: plan.gradient(mat), 144 + (232 if mat else 152))):
        fn(); ctx.sync()
        ts = []
        for _ in range(reps):
            ctx.timer_start(); fn(); ts.append(ctx.timer_stop_ms())
        t = min(ts)
        print("elastic v%d %dx%d nt=%d %-8s mat=%d %8.2f ms  %7.2f Gcell/s  %7.1f GB/s(est)  %s" %
              (variant, NX, NY, NSTEP, name, mat, t, cells / t / 1e6, cells * byt / t / 1e6, plan.info()), flush=True)
    plan.close()


if __name__ == "__main__" and "--elastic-rb" in sys.argv:
    for rb in sys.argv[sys.argv.index("--elastic-rb") + 1].split(","):
        os.environ["ADSEIS_EL_RB"] = rb
        print("EL_RB", rb)
        if "--big" in sys.argv:
            run_elastic(4096, 4096, 16, 0, reps=2)
        else:
            run_elastic(2000, 2000, 40, 1, reps=2)
    sys.exit(0)

if __name__ == "__main__" and "--elastic-one" in sys.argv:
    run_elastic(2000, 2000, 12, 1, reps=1)
    sys.exit(0)

if __name__ == "__main__" and "--elastic-quick" in sys.argv:
    print("variant", os.environ.get("ADSEIS_LIB_SUFFIX", "(default)"))
    run_elastic(2000, 2000, 60, 1)
    run_elastic(2000, 2000, 60, 1, mat=False)
    sys.exit(0)

if __name__ == "__main__" and "--elastic" in sys.argv:
    run_elastic(500, 500, 200, 0)
    run_elastic(2000, 2000, 60, 1)
    run_elastic(2000, 2000, 60, 1, mat=False)
    run_elastic(4096, 4096, 20, 0)
