import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session")
def ref_harness():
    """oracle/_ref/vic_ref_harness: the reference's own physics built by oracle/Makefile (travels to the GPU box)."""
    p = os.path.join(ROOT, "oracle", "_ref", "vic_ref_harness")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/vic_ref_harness not built (needs /root/reference; see oracle/Makefile)")
    return p


@pytest.fixture(scope="session")
def vicport():
    p = os.path.join(ROOT, "oracle", "_ref", "vicport")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/vicport not built (oracle/Makefile)")
    return p
