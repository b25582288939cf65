"""pytest configuration: the `gpu` marker and shared fixtures."""
import os
import sys

import pytest

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _have_gpu() -> bool:
    try:
        from skypilot_b200 import _native
        return _native.device_count() > 0
    except Exception:  # pylint: disable=broad-except
        return False


def pytest_collection_modifyitems(config, items):
    del config
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
