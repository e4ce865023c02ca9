import os
import sys

import pytest

# the single-device tests of the sharded MSM run up to 8 ranks (contexts) on one GPU, each spinning on flags another one sets:
# every stream needs its own hardware queue (default 8 are shared round-robin), set before CUDA initialises
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """One engine for the whole session; fails loudly when the CUDA library or the GPU is missing."""
    from kyber_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.close()
