import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    # a test marked gpu must never silently pass on a CPU-only box
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (builds it if the .so is missing or stale)."""
    from mamba_b200 import build as _b
    _b.build()
    from mamba_b200 import _lib
    return _lib.lib()
