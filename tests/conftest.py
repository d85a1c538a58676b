# -*- coding: utf-8 -*-
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # docstrings of the reference tree (imported in place by tests/golden/refload.py) carry
    # invalid escape sequences
    config.addinivalue_line("filterwarnings", "ignore::SyntaxWarning")


def golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def stencils_golden():
    return golden("stencils")


@pytest.fixture(scope="session")
def stencils_1d_golden():
    return golden("stencils_1d")
