# -*- coding: utf-8 -*-
"""pytest plugin for the GPU-less container: every test runs with the oracle-backed C-ABI stub in
place of the library (host storages), so the host side of the ``-m gpu`` tests can be exercised
without a GPU:

    python -m pytest tests/test_gpu_moist_model.py -m gpu -p experiments.stub_plugin -q

Tests that need CUDA itself (streams, graphs, peer access, the library context) fail there by design."""
import pytest

from tests.abi_oracle import OracleStub
from tests.abi_stub import stubbed_library


@pytest.fixture(autouse=True)
def _oracle_stub():
    with stubbed_library(OracleStub) as stub:
        yield stub
