import os
import sys

import pytest

# tests/test_peer_gpu.py emulates several ranks on ONE device: every rank's stream needs its own hardware queue,
# or a rank's spinning barrier kernel blocks the kernels of the rank it waits for (default: 8 connections)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and no kernel may be loaded lazily while another rank's barrier kernel spins on the same device (a module load
# synchronises the device): load every module when the library is opened.  One rank per GPU does not need this.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
