import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """ZKB_TEST_ORDER=reverse / a seed: run the collected tests in another order (all GPU tests share ONE context, so
    state that one test leaves behind -- cached domains, the Groth16 stage, pooled scratch -- must not matter)"""
    order = os.environ.get("ZKB_TEST_ORDER")
    if not order:
        return
    if order == "reverse":
        items.reverse()
    else:
        import random
        random.Random(int(order)).shuffle(items)


@pytest.fixture(scope="session")
def ctx():
    """One zkb context on cuda:0 for the whole GPU session.  Fails (never skips) when the
    extension or the device is missing: a green GPU run must mean the CUDA path ran."""
    from ckb_zkp_b200.backend import Context
    c = Context(0)
    yield c
    c.close()
