import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")


def golden_cases():
    return ["gemat11_k1", "gemat11_k2", "gemat11_k3_hp", "gemat11_k3_rp", "karate_k3_hp", "karate_k3_stchp"]


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
