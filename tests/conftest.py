import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `pytest -m gpu` on the B200 box)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def backend_params():
    """Backends a test can run on: the host twin everywhere, the CUDA C-ABI on GPU boxes."""
    return [pytest.param("hostsim", id="hostsim"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def backend_factory(name):
    if name == "hostsim":
        from hostsim.backend import HostSimBackend
        return HostSimBackend
    from rware_b200._capi import CudaBackend
    return CudaBackend


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    oracle_py.build()
    yield
