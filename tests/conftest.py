import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout: wall-clock limit of a test (pytest-timeout)")


def pytest_collection_modifyitems(config, items):
    # every GPU test runs under a wall-clock limit (pytest-timeout, thread method: it also ends a test stuck inside a CUDA
    # call): a kernel that never returns fails its test instead of hanging the session
    if config.pluginmanager.hasplugin("timeout"):
        for item in items:
            if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
                item.add_marker(pytest.mark.timeout(600, method="thread"))
    _skip_gpu_without_cuda(items)


def _skip_gpu_without_cuda(items):
    """Tests marked ``gpu`` need a CUDA device AND the built library: skip them elsewhere (CPU CI runs plain
    ``pytest tests``).  On a GPU box a missing library is a hard error, not a skip."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
