import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes a minute or more (still part of -m gpu)")


def _have_gpu() -> bool:
    return os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; parity unpinned — see oracle/voidray_oracle.cpp)."""
    from oracle import oracle as O
    O.load()
    return O


@pytest.fixture(scope="session")
def ctx():
    """A vr_context on cuda:0. Fails loudly if the CUDA extension is missing."""
    from voidray_b200.render import Context
    return Context(0)
