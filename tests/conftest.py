import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libtsdref.so not built (needs /root/reference)")
    return oracle.ref()


@pytest.fixture(scope="session")
def cpu_oracle():
    """The strongest CPU oracle available: the reference build if present, else the C port."""
    import oracle
    return oracle.ref() if oracle.have_ref() else oracle.port()
