import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import stark_rs_b200 as S
    c = S.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def S():
    import stark_rs_b200 as S
    return S
