import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The built libraries are git-ignored: on a fresh checkout build them once (nvcc + gcc,
    about a minute) so the suite is self-sufficient.  On the GPU box they arrive prebuilt."""
    needed = [os.path.join(ROOT, "gst-plugins-rs_b200", "libb200vf.so"),
              os.path.join(ROOT, "gst-plugins-rs_b200", "libb200vf_elements.so"),
              os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in needed):
        import __graft_entry__
        __graft_entry__.build()


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
