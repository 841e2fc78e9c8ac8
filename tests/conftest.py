import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def hostsim():
    """The CUDA library's source compiled for the CPU under the CUDA-semantics emulator."""
    from backends import hostsim_backend

    return hostsim_backend()


@pytest.fixture(scope="session")
def gpu():
    """The product library on cuda:0.  Fails (does not skip) if CUDA or the library is missing."""
    import torch

    assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
    from backends import GpuBackend

    return GpuBackend()


def backend_params():
    return [pytest.param("hostsim", id="hostsim"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture(params=backend_params())
def backend(request):
    return request.getfixturevalue(request.param)
