import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def handle():
    from quadruped_gait_generation_ismpc_b200 import binding
    import torch
    if not torch.cuda.is_available():
        # a box without a GPU: the GPU tests skip (run with -m "not gpu" there); on a GPU box a handle that cannot be
        # created is a failure, not a skip
        pytest.skip("no CUDA device: GPU test skipped (the product has no CPU fallback)")
    h = binding.Handle(device=0, max_batch=1 << 17)
    yield h
    h.close()
