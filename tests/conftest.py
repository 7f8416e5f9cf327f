import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The two-steps-per-launch kernels (temporal blocking) switch on automatically only for large grids; the parity tests
    # run on small ones, so they force the path on (tests that compare it with the one-step path set "0" themselves).
    os.environ.setdefault("ADSEIS_AC_TB", "1")


@pytest.fixture(scope="session")
def po():
    """The CPU checkers (oracle/liboracle.so and, when present, oracle/_ref)."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def A():
    """The product package (adseismic.jl_b200/), imported through the adseis_b200 shim."""
    import adseis_b200
    return adseis_b200


@pytest.fixture(scope="session")
def ctx(A):
    return A.default_context()


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name))


def relerr(a, b):
    import numpy as np
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    d = np.abs(b).max()
    return float(np.abs(a - b).max() / d) if d > 0 else float(np.abs(a - b).max())


def on_both_records():
    """Decorator for CPU-only tests that must ALSO appear on the GPU box's `-m gpu` record (the chain
    CUDA -> oracle.c -> reference op bodies is then closed in one run): the test is collected twice, once plain
    (runs under -m "not gpu") and once carrying the gpu marker."""
    return pytest.mark.parametrize("record", ["cpu", pytest.param("gpubox", marks=pytest.mark.gpu)])
