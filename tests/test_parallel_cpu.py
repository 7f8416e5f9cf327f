"""CPU tests (gloo, world_size 2) of the multi-GPU host logic in adseismic.jl_b200/parallel.py, with the CPU oracle
standing in for the per-rank GPU compute:
  * shot sharding (reference rule k % n_gpu) + all-reduce of loss / gradient == the sum over all shots
  * slab geometry / ownership / halo-row addressing: an emulated slab-decomposed time loop (oracle step per slab,
    halo rows exchanged with gloo send/recv exactly as the GPU path addresses them) == the undecomposed loop --
    the reference's distributed invariant (test/verify_forward.jl:32-90, decomposed == undecomposed)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    return [ret[r] for r in range(world)]


def test_shot_assignment_matches_reference_rule():
    import adseis_b200 as A
    from adseis_b200 import parallel
    for n, w in ((8, 8), (64, 8), (5, 2), (3, 4)):
        jobs = parallel.shot_assignment(n, w)
        assert sorted(sum(jobs, [])) == list(range(n))
        for r, js in enumerate(jobs):      # src/Utils.jl:326: jobs = [k for k=1:n if k%n_gpu==i-1]
            assert js == [k - 1 for k in range(1, n + 1) if k % w == r]


def test_slab_geometry_covers_grid():
    from adseis_b200 import parallel
    for NX, NY, w in ((4096, 4096, 8), (97, 600, 2), (10, 33, 3)):
        prev = 0
        for r in range(w):
            g = parallel.slab_geometry(NX, NY, w, r)
            assert g["row0"] == prev and g["own1"] - g["own0"] == g["row1"] - g["row0"]
            assert g["goff"] == g["row0"] - (1 if r > 0 else 0) and g["ld"] % 16 == 0 and g["ld"] >= NY + 2
            prev = g["row1"]
        assert prev == NX + 2


def _shots_fn(rank, world):
    import torch
    from adseis_b200 import parallel
    from oracle import pyoracle as po
    rng = np.random.default_rng(3)
    NX, NY, NSTEP, dx, dt, vp, nshots = 30, 40, 40, 10.0, 1e-3, 2000.0, 5
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=5, vp_ref=vp)
    c = vp * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    L, g = 0.0, torch.zeros((NX + 2, NY + 2), dtype=torch.float64)
    Lall, gall = 0.0, np.zeros((NX + 2, NY + 2))
    mine = parallel.shot_assignment(nshots, world)[rank]
    for k in range(nshots):
        si, sj = np.array([6 + 4 * k]), np.array([8 + 5 * k])
        sv = po.ricker(NSTEP, 7.0, 15.0, 1e6).reshape(-1, 1)
        ri, rj = np.full(10, 4), np.arange(5, 35, 3)
        u, r = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, si, sj, sv, ri, rj)
        Lk, gk, _ = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, si, sj, ri, rj, 0.5 * r, u)
        Lall += Lk; gall += gk
        if k in mine:
            L += Lk; g += torch.from_numpy(gk)
    parallel.all_reduce_sum_(g)
    L = parallel.all_reduce_scalar(L, "sum", device="cpu")
    return abs(L - Lall) / Lall, float(np.abs(g.numpy() - gall).max() / np.abs(gall).max())


def test_shot_parallel_reduction_gloo():
    for dl, dg in _spawn(_shots_fn, 2):
        assert dl < 1e-14 and dg < 1e-14


def _dd_fn(rank, world):
    """Emulated slab time loop: each rank advances its rows with the oracle step on its local (halo-padded) arrays
    and swaps the edge rows of u and phi with its neighbours after every step."""
    import torch
    import torch.distributed as dist
    from adseis_b200 import parallel
    from oracle import pyoracle as po
    rng = np.random.default_rng(8)
    NX, NY, NSTEP, dx, dt, vp = 23, 30, 50, 10.0, 1e-3, 2000.0
    W = NY + 2
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=4, vp_ref=vp)
    c = vp * (1 + 0.1 * rng.random((NX + 2, W)))
    srci, srcj = np.array([5, 12, 13, 20]), np.array([7, 15, 15, 22])   # rows around the slab boundary
    srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 12.0, 1e6) for k in range(4)], 1)
    u_ref, _ = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci, srcj, srcv, [], [])
    g = parallel.slab_geometry(NX, NY, world, rank)
    Hl, goff, own0, own1 = g["Hl"], g["goff"], g["own0"], g["own1"]
    rows = slice(goff, goff + Hl)
    c2 = (c * c)[rows]
    sg, ta = sig.reshape(NX + 2, W)[rows], tau.reshape(NX + 2, W)[rows]
    mine = parallel.owned_points(srci, g["row0"], g["row1"], False)
    u = [np.zeros((Hl, W)), np.zeros((Hl, W))]
    phi, psi = np.zeros((Hl, W)), np.zeros((Hl, W))

    def swap(a):
        t = torch.from_numpy(a)
        reqs = []
        if rank > 0:
            reqs += [dist.isend(t[own0].clone(), rank - 1), dist.irecv(t[0], rank - 1)]
        if rank < world - 1:
            reqs += [dist.isend(t[own1 - 1].clone(), rank + 1), dist.irecv(t[Hl - 1], rank + 1)]
        for r in reqs:
            r.wait()

    for s in range(2, NSTEP + 1):
        # oracle step on the local array: its own "ring" rows are my halo rows, so take the interior result only
        un, pn, qn = po.acoustic_step_fwd(u[-1].ravel(), u[-2].ravel(), phi.ravel(), psi.ravel(), sg.ravel(),
                                          ta.ravel(), c2.ravel(), dt, dx, dx, Hl - 2, NY)
        un, pn, qn = un.reshape(Hl, W), pn.reshape(Hl, W), qn.reshape(Hl, W)
        new = np.zeros((Hl, W)); nphi = np.zeros((Hl, W)); npsi = np.zeros((Hl, W))
        lo = max(own0, 1 - goff)                     # global ring rows 0 / NX+1 stay zero
        hi = min(own1, NX + 1 - goff)
        new[lo:hi], nphi[lo:hi], npsi[lo:hi] = un[lo:hi], pn[lo:hi], qn[lo:hi]
        for k in np.nonzero(mine)[0]:
            new[srci[k] - 1 - goff, srcj[k] - 1] += srcv[s - 1, k] * (dt * dt)
        swap(new); swap(nphi)
        u.append(new); phi, psi = nphi, npsi
    err = 0.0
    for s in (2, NSTEP // 2, NSTEP):
        err = max(err, float(np.abs(u[s][own0:own1] - u_ref[s][g["row0"]:g["row1"]]).max()))
    return err


def test_slab_halo_protocol_gloo():
    for err in _spawn(_dd_fn, 2):
        assert err == 0.0


def test_elastic_slab_partition_and_ownership():
    """Host logic of the elastic slab decomposition (MPIElastic.jl block ownership, :71-86, 116-131): the slabs tile
    the internal rows exactly, interior rows are balanced, every source / receiver is owned by exactly one slab."""
    import adseis_b200 as A
    from adseis_b200 import parallel
    rng = np.random.default_rng(3)
    for variant, ghost in ((0, 1), (1, 2)):
        for NX, world in ((2000, 8), (37, 2), (103, 4), (16, 4)):
            p = A.ElasticPropagatorParams(NX=NX, NY=50, NSTEP=4, variant=variant)
            bounds = [parallel.elastic_slab_partition(p, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == NX + 2 * ghost
            for (a0, a1), (b0, b1) in zip(bounds[:-1], bounds[1:]):
                assert a1 == b0 and a1 > a0
            interior = [min(b1, NX + ghost) - max(b0, ghost) for b0, b1 in bounds]
            assert sum(interior) == NX and max(interior) - min(interior) <= 1
            lo, hi = (1, NX + 2) if variant == 0 else (1, NX)       # the caller's 1-based index range
            pi = rng.integers(lo, hi + 1, 200)
            own = np.stack([parallel.elastic_owned_points(p, pi, b0, b1) for b0, b1 in bounds])
            assert np.all(own.sum(0) == 1)


def test_plan_window_budget():
    """The slab window chosen on the host (parallel.plan_window) obeys the budget the C side checks: W snapshots +
    4 checkpoint planes per extra segment fit, W is maximal, and the whole tape is kept when it fits."""
    from adseis_b200.parallel import plan_window
    assert plan_window(5000, 6000) == 5001 and plan_window(50, 51) == 51
    for NSTEP, slots in ((5000, 2700), (5000, 1390), (1000, 150), (60, 30), (200, 120)):
        W = plan_window(NSTEP, slots)
        nseg = -(-(NSTEP - 1) // (W - 2))
        assert 6 <= W <= slots and W + 4 * (nseg - 1) <= slots
        if W + 1 <= slots:                         # a larger window would not fit
            nseg1 = -(-(NSTEP - 1) // (W - 1))
            assert W + 1 + 4 * (nseg1 - 1) > slots
    with pytest.raises(MemoryError):
        plan_window(1000, 5)


def test_packed_halo_words_roundtrip():
    """The slab halo rows travel as 8-byte words = 32 data bits + the sender's 32-bit launch epoch, two words per double
    (csrc/common.cuh: ll_put / ll_get).  NumPy restatement of that packing: every bit pattern survives (NaN payloads, -0.0,
    denormals), and a pair whose words carry different epochs -- a torn or stale read -- is never accepted."""
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 2**64, 4096, dtype=np.uint64)
    bits[:6] = np.array([0.0, -0.0, 5e-324, np.inf, -np.inf, np.nan]).view(np.uint64)
    ep = np.uint64(0x9E3779B9)

    def put(b, e):
        return (e << np.uint64(32)) | (b & np.uint64(0xFFFFFFFF)), (e << np.uint64(32)) | (b >> np.uint64(32))

    def get(x, y, e):
        ok = ((x >> np.uint64(32)) == e) & ((y >> np.uint64(32)) == e)
        return ok, (x & np.uint64(0xFFFFFFFF)) | (y << np.uint64(32))

    x, y = put(bits, ep)
    ok, back = get(x, y, ep)
    assert ok.all() and np.array_equal(back, bits)
    # the receiver of launch n+2 polls the same parity buffer that still holds epoch n: not accepted
    ok_old, _ = get(x, y, ep + np.uint64(2))
    assert not ok_old.any()
    # torn pair: low word already rewritten by the next sender, high word still the old one
    x2, _ = put(bits ^ np.uint64(1), ep + np.uint64(2))
    ok_torn, _ = get(x2, y, ep + np.uint64(2))
    assert not ok_torn.any()
    # epoch 0 is "never written" (the arena is zero-filled): a launch that expects epoch >= 1 never accepts zeros
    ok_zero, _ = get(np.zeros(4, np.uint64), np.zeros(4, np.uint64), np.uint64(1))
    assert not ok_zero.any()
