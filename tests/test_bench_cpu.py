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


def test_algorithmic_bytes_match_design():
    sys.path.insert(0, ROOT)
    import bench
    w = bench.workload_c4()
    ab = bench.algorithmic_bytes(w)
    N, Np = 4096 * 4096, bench.n_pml_cells(w)
    assert Np == N - (4096 - 26) ** 2
    assert ab["forward"] == 32 * (N - Np) + 64 * Np and ab["adjoint"] == 56 * (N - Np) + 88 * Np
    assert abs(ab["forward"] - 543.7e6) < 0.1e6 and abs(ab["adjoint"] - 946.3e6) < 0.1e6
