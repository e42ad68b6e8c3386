import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def vsf_ctx():
    """A session-wide context for the GPU parity tests (32-byte descriptors)."""
    import vision_slam_frontend_b200 as vsf
    ctx = vsf.Context(device=0, max_features=20480, desc_bytes=32, window=10)
    yield ctx
    ctx.close()
