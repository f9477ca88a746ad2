import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    """One zkb context on cuda:0 for the whole GPU session.  Fails (never skips) when the
    extension or the device is missing: a green GPU run must mean the CUDA path ran."""
    from ckb_zkp_b200.backend import Context
    c = Context(0)
    yield c
    c.close()
