import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """One FockEngine per test session (GPU tests only)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sqcc_b200.backend import FockEngine
    eng = FockEngine()
    yield eng
    eng.close()
