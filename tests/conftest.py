import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 and the compiled libbd_b200.so")
    config.addinivalue_line("markers", "slow: long-running CPU oracle comparison")


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not have_gpu():
        pytest.skip("no CUDA device")
    return True


@pytest.fixture(scope="session")
def parity_models():
    """name -> engine.Model carrying the parity-test weights (oracle.nets.parity_weights), built once per session:
    the calibration pass is a CPU forward of the oracle."""
    from building_detection_b200.predict_model import CTORS
    from oracle import nets
    cache = {}

    def get(name):
        if name not in cache:
            m = CTORS[name]()
            m.set_weights(nets.parity_weights(name, m.spec))
            cache[name] = m
        return cache[name]
    return get
