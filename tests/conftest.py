import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "refonly: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def oracle_libs():
    """Build (if needed) the CPU oracle libraries; building the checker is not using it."""
    from oracle import kernels as ok
    ok.build(ref=True, port=True)
    return ok
