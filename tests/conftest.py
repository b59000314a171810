import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def rts():
    from powersystemsreliabilityassessment_b200 import rts79
    cap, mttf, mttr = rts79.units()
    return dict(cap=cap, mttf=mttf, mttr=mttr, load_mw=rts79.load_curve_mw(), load_int=rts79.load_curve_int())


@pytest.fixture(scope="session")
def engine():
    """One CUDA engine for the GPU tests; fails loudly (no fallback) if the library or GPU is missing."""
    from powersystemsreliabilityassessment_b200 import Engine
    eng = Engine()
    yield eng
    eng.close()
