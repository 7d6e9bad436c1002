import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def blood_arrays():
    from metabodecon_rust_b200.readers import read_bruker_arrays
    x, y, _ = read_bruker_arrays(os.path.join(GOLDEN, "bruker", "blood_01"), 10, 10)
    return x, y


@pytest.fixture(scope="session")
def sim_arrays():
    from metabodecon_rust_b200.readers import read_bruker_arrays
    x, y, _ = read_bruker_arrays(os.path.join(GOLDEN, "bruker", "sim_01"), 10, 10)
    return x, y
