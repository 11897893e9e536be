import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


GPU_TEST_TIMEOUT_S = 300  # a hung kernel must cost minutes, not the whole run (round 1 lost a 1200 s driver run to one)


def pytest_collection_modifyitems(config, items):
    """Every GPU test gets a hard per-test timeout (pytest-timeout, thread method: a thread blocked inside
    cudaStreamSynchronize cannot be interrupted by a signal, so the watchdog dumps all stacks and ends the process)."""
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(GPU_TEST_TIMEOUT_S, method="thread"))


@pytest.fixture(scope="session")
def oracle_backend():
    from refcases import OracleBackend
    from oracle import gfo
    gfo.set_workers(1)  # small grids: threads only add overhead
    return OracleBackend()
