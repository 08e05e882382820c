"""pytest configuration: registers the ``gpu`` marker and makes the repo root importable."""
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_SRC = "/root/reference/src"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def live_reference():
    """The unmodified reference package when it is importable (build container), else skip."""
    if not os.path.isdir(REFERENCE_SRC):
        pytest.skip("/root/reference not present (GPU box): live pinning runs in the build container")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import recsys_lite
    return recsys_lite
