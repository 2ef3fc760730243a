import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


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


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _restore_eos_constants():
    """SpamComplete sets the module-global equation-of-state constants, as the reference does with the Fortran eos
    module (spam_complete_force.py:49-51); tests must not leak them into each other."""
    from pyticles_b200 import properties
    saved = (properties.ADASH, properties.BDASH, properties.KBDASH)
    yield
    properties.ADASH, properties.BDASH, properties.KBDASH = saved
