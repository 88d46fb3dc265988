import os
import sys

import pytest

# Several ranks may share one GPU (tests/loopback.py, tests/test_multi_rank_dropin.py: the rank
# processes inherit this environment): every stream gets its own hardware queue, so that a kernel
# spinning on a peer's flag never sits in front of that peer's work.  Must be set before CUDA
# initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and no kernel is loaded lazily at its first launch (which waits for the device) while a
# peer already spins on this rank
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
