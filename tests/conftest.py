import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gnn-ops-benchmark_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run under gpurun on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the C-ABI library and the oracle exist (both are built by __graft_entry__.build)."""
    lib = os.path.join(PKG, "lib", "libgno_b200.so")
    if not os.path.exists(lib):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    import oracle
    oracle.oracle.build()
    yield


@pytest.fixture
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
