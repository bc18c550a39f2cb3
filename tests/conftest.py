import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _live_tuning_knobs(monkeypatch):
    """libdfa_b200.so reads its DFA_* tuning environment variables once per process; tests that switch
    kernels with monkeypatch.setenv need the library to look again (dfa_debug_reload_knobs)."""
    if not _has_cuda():
        yield
        return
    from simpb_b200 import cabi
    cabi.reload_knobs()      # whatever the previous test left behind is gone
    set_env, del_env = monkeypatch.setenv, monkeypatch.delenv

    def setenv(name, value, *a, **k):
        set_env(name, value, *a, **k)
        cabi.reload_knobs()

    def delenv(name, *a, **k):
        del_env(name, *a, **k)
        cabi.reload_knobs()

    monkeypatch.setenv, monkeypatch.delenv = setenv, delenv
    yield
