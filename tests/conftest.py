import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ensure_library():
    """The CUDA library is built in-tree (nvcc cross-compiles without a GPU); a fresh checkout has no .so yet."""
    try:
        from dfdm_b200 import build as _build
        if _build.stale():
            _build.build()
    except Exception as exc:          # the tests that need the library will fail loudly on import
        print("dfdm_b200: could not build libdfdm_b200.so: %s" % exc)


_ensure_library()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
