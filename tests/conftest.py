import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


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


@pytest.fixture
def ffvc_options():
    """set kernel-selection switches (ffvc_set_option) for one test; restored afterwards"""
    from feed_forward_vqgan_clip_b200 import _lib
    lib = _lib.load()
    saved = {}

    def set_(**kw):
        for k, v in kw.items():
            old = lib.ffvc_set_option(k.encode(), int(v))
            assert old >= 0, "unknown option %s" % k
            saved.setdefault(k, old)

    yield set_
    for k, v in saved.items():
        lib.ffvc_set_option(k.encode(), v)
