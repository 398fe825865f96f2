import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
