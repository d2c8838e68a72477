import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_runtest_logreport(report):
    """Failures are also appended to gpurun_out/test_failures.log: that directory comes back from a GPU box even when
    only the tail of the terminal output does (one full run of the GPU suite on a fresh box stopped at a
    failure whose text was lost and which did not repeat — DESIGN.md section 9)."""
    if report.failed:
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "test_failures.log"), "a") as f:
                f.write(f"==== {report.nodeid} [{report.when}]\n{report.longrepr}\n")
        except OSError:
            pass


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
