import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def corc():
    import oracle
    oracle.build()
    return oracle.corc()


@pytest.fixture(scope="session")
def reflib():
    import oracle
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return r


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from srcdsp_b200 import build as b
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        b.build()
    from srcdsp_b200 import _capi
    return _capi.lib()
