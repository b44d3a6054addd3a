import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: full-size (10M-row) checks; opt in with SHODH_SLOW=1")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have = _have_gpu()
    for item in items:
        if "gpu" in item.keywords and not have:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "slow" in item.keywords and os.environ.get("SHODH_SLOW") != "1":
            item.add_marker(pytest.mark.skip(reason="set SHODH_SLOW=1"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def sg():
    """The product library (C-ABI over the CUDA kernels), built in-tree if stale."""
    import __graft_entry__ as ge
    ge.build()
    import shodh_memory_b200 as pkg
    return pkg
