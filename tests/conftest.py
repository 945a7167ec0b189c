"""pytest configuration: markers, import path, shared fixtures.

`-m "not gpu"` : oracle vs golden vectors / libjpeg-turbo / the reference's own
                 parser+kernels, host logic, C-ABI symbol checks (CPU only).
`-m gpu`       : parity tests proper — the CUDA path through the C ABI against
                 the oracle (needs a B200).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "slow: larger CPU cases")


@pytest.fixture(scope="session")
def orc():
    import oracle

    return oracle.Oracle()


@pytest.fixture(scope="session")
def ljt():
    import oracle

    return oracle.LibJpegTurbo()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
