import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a machine without a CUDA device: gpu-marked tests are skipped instead of failing in rem2d_create
    (the product has no CPU fallback by design). `-m gpu` on such a machine skips them all - visibly."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200 box): librem2d_cuda.so has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
