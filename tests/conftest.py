import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (checker only)."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def vf():
    import gst_plugins_rs_b200 as g
    return g


@pytest.fixture()
def ctx(vf):
    import torch
    assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
    c = vf.Context(0)
    yield c
    c.close()
