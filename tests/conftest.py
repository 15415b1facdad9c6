import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def synth(pkg):
    import importlib
    return importlib.import_module("tpose_b200.synth")


@pytest.fixture(scope="session")
def O():
    return entry.load_oracle()


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda_pkg(pkg):
    """The CUDA library, or a loud failure: GPU tests never fall back to anything else."""
    if not os.path.exists(pkg.LIB_PATH):
        pytest.fail("libtpose_cuda.so is not built (run __graft_entry__.build())")
    assert has_cuda(), "gpu-marked test running without a CUDA device"
    return pkg


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(params=["table", "tiles"])
def kernel_path(request, monkeypatch):
    """Both edge-integration paths of the library (row-prefix table gathers / TMA-staged tiles) must give
    the oracle's bits; contexts created inside the test pick the path up from the environment."""
    monkeypatch.setenv("TPOSE_CUDA_PATH", request.param)
    return request.param
