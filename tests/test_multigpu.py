"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): launches tests/mgpu_worker.py under torchrun."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_domain_decomposition_and_shots():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (have %d)" % n)
    world = 8 if n >= 8 else (4 if n >= 4 else 2)     # decomposition invariance (gradients too) at up to 8 ranks
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "shot-parallel x%d ok" % world in r.stdout
