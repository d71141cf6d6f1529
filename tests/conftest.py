import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def h1_model_path():
    return os.path.join(ROOT, "configs", "h1.model")


@pytest.fixture(scope="session")
def oracle_h1(h1_model_path):
    from oracle.pyoracle import Oracle, build
    build()
    return Oracle(h1_model_path)
