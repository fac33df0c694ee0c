import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


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
