import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure; see oracle/ns3d_oracle.c)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ns():
    import navierstokes3d_b200 as ns
    return ns


@pytest.fixture()
def ctx(ns):
    """A fresh PARITY-mode context on cuda:0 (gpu tests only)."""
    c = ns.Context(0, ns.PARITY)
    yield c
    c.close()
