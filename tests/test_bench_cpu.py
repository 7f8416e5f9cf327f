"""bench.py contract on the CPU side: the reference arm (`--impl reference`) times the reference's own op bodies
(oracle/_ref when it was built here, else the plain-C port) on the host cores and prints ONE JSON line with the keys
the driver reads.  Small grid so that the test takes seconds."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nx", "256", "--ny", "256",
           "--nstep", "40", "--cpu-steps", "8", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gcell-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] == 0 and d["config"]["workload"].startswith("C4 acoustic 256x256")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_match_survey_8d():
    """roofline.achieved is computed from SURVEY.md 8(d)'s per-cell figures: acoustic 32 / 56 B (+32 in the PML frame),
    elastic 104 / 192 B (fused ideal) -- not from what the kernels happen to move."""
    sys.path.insert(0, ROOT)
    import adseis_b200 as A
    W = A.workloads
    w = W.c4(nstep=8)
    ab = W.algorithmic_bytes(w)
    N = 4096 * 4096
    Np = N - (4096 - 26) ** 2
    assert ab["forward"] == 32 * N + 32 * Np and ab["adjoint"] == 56 * N + 32 * Np
    assert abs(ab["forward"] - 543.7e6) < 0.1e6 and abs(ab["adjoint"] - 946.3e6) < 0.1e6
    e = W.algorithmic_bytes(W.c5(nstep=4))
    n = 2000 * 2000
    assert 104 * n <= e["forward"] <= 105 * n and 192 * n <= e["adjoint"] <= 193 * n


def test_reference_arm_ignores_inherited_omp_num_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must still use every core it may run on."""
    sys.path.insert(0, ROOT)
    import bench
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nx", "256", "--ny", "256",
           "--nstep", "40", "--cpu-steps", "6", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    if d["cpu_baseline"]["kind"] == "reference":
        assert d["cpu_baseline"]["cores"] == bench.host_threads()


def test_workload_builders_are_deterministic():
    sys.path.insert(0, ROOT)
    import numpy as np
    import adseis_b200 as A
    for name in ("c1", "c2", "c5"):
        a, b = A.workloads.BUILDERS[name](nstep=20), A.workloads.BUILDERS[name](nstep=20)
        ma, mb = a["model_obs"], b["model_obs"]
        ma, mb = (ma, mb) if isinstance(ma, np.ndarray) else (ma[1], mb[1])
        assert np.array_equal(ma, mb) and a["shots"][0]["srcv"].shape[0] == 20
    w = A.workloads.c3(nstep=10, shots=64)
    assert len(w["shots"]) == 64 and w["shots"][0]["srci"][0] == 40 and w["shots"][-1]["srci"][0] == 1960
