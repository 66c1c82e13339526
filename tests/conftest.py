import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def set_deno_env(monkeypatch, name, value):
    """Set a DENO_* switch for this test and make the library read it (it reads the environment once)."""
    monkeypatch.setenv(name, value)
    from deno4pytorch_b200.functional import reload_env
    reload_env()


@pytest.fixture(autouse=True)
def _deno_env_is_fresh(request):
    """Every GPU test starts with the library's switches re-read from the (restored) environment, so a switch set by
    the previous test cannot leak.  DENO_POISON=1 in the environment poisons every output / scratch buffer."""
    if "gpu" in request.keywords:
        import torch
        if torch.cuda.is_available():
            from deno4pytorch_b200.functional import reload_env
            reload_env()
    yield
