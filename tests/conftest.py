import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The step uses the one-pass `all` sweep (predicted re-binning) only for patches whose tile-kernel grid fills the GPU
# (IonUpdater.predict_min_cells); the test problems are small, so the threshold is lifted here: the solver-level tests then
# exercise the path the benchmark runs.  tests/test_predict_gpu.py checks the threshold itself and the two-pass path.
os.environ.setdefault("PHB_PREDICT_MIN_CELLS", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cpu_oracle():
    import oracle
    oracle.build(ref=False)
    return oracle.Cpu("oracle")


@pytest.fixture(scope="session")
def cpu_ref():
    import oracle
    if not oracle.have_ref():
        if os.path.isdir("/root/reference/src"):
            oracle.build(ref=True)
        else:
            pytest.skip("oracle/_ref/libphare_ref.so absent and /root/reference not available")
    return oracle.Cpu("ref")
