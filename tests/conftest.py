import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The native pieces are built by __graft_entry__.build(); build on demand
    when a test run finds them missing (e.g. a fresh checkout)."""
    need = [os.path.join(ROOT, "oracle", "liboracle.so"),
            os.path.join(ROOT, "sleipnir_b200", "lib", "libslpb.so"),
            os.path.join(ROOT, "sleipnir_b200", "lib", "libslpb_host.so"),
            os.path.join(ROOT, "tests", "emu", "libslpb_emu.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
