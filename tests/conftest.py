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
def oracle_lib():
    """Builds (if needed) and returns the CPU oracle of the point ops."""
    from oracle import point_ops
    point_ops.build()
    return point_ops


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests must run on it (no fallback)."""
    import torch
    assert torch.cuda.is_available()
    from butd_detr_b200 import _lib, build
    build.build()
    _lib.load()
    return _lib
