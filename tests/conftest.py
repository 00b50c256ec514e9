import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_available() -> bool:
    try:
        from needle_b200.engine import Context
        Context(0).close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not errored) on a machine without a CUDA device."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _cuda_device_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (sm_100a); run on the B200 box with -m gpu")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; compiled on demand with gcc)."""
    from oracle import oracle as orc
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def ctx():
    """One library context on cuda:0 for the whole GPU session."""
    from needle_b200.engine import Context
    c = Context(0)
    yield c
    c.close()
