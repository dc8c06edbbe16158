"""pytest configuration: registers the `gpu` marker, puts the repo root on sys.path, and skips `@pytest.mark.gpu`
tests on a machine without a usable CUDA device (the product itself has no CPU fallback: it raises)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def _cuda_usable() -> bool:
    try:
        import mcmcdiag_b200 as m
        m.get_context(0)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device / libmcmcdiag_b200.so (GPU tests run on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)
