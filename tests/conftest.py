import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_engine_device() -> bool:
    """lwse_create succeeds only on an sm_100 device (the engine has no CPU path)."""
    try:
        from lws_b200.engine import Engine

        Engine(0).close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a B200 skips the gpu-marked tests instead of
    failing at the first lwse_create.  `-m gpu` states the intent to run them: nothing is skipped
    then, and a missing device or library fails loudly."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _have_engine_device():
        skip = pytest.mark.skip(reason="no sm_100 CUDA device: lwse_create returns LWSE_ERR_NO_DEVICE")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_sweep():
    """sweep(tables, flags) backed by the CPU oracle."""
    import oracle

    oracle.build()

    def sweep(tables, flags=0):
        lws_out, group_out, _ = oracle.sweep_lws(
            tables.lws, tables.groups, tables.pod_state, tables.pod_ident, tables.nodes, flags=flags
        )
        return lws_out, group_out

    return sweep
