import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def b200():
    import __graft_entry__ as g

    return g.load_package()


@pytest.fixture(scope="session")
def gpu(b200):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    b200.init(0)
    return b200
