import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "gpu_next: GPU tests of code that has not had its first run on a B200 yet "
                                       "(validated with tools/gpu_check first, then promoted to `gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native libraries once (no-op when up to date)."""
    import __graft_entry__ as g

    g.build(verbose=False)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kokkos_kernels_b200 as kk

    assert kk._lib.sparse().b200sp_device_ok() == 1, "libb200sparse needs a compute-capability 10.x device"
    return torch.device("cuda:0")
