import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


def _has_gpu():
    try:
        from whisper_apr_b200 import _lib
        return _lib.lib().wb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests FAIL (not skip) when selected on a box without the extension; they are only skipped
    # when no device exists and they were not explicitly selected with -m gpu.
    if _has_gpu():
        return
    selected = config.getoption("-m") or ""
    if "gpu" in selected and "not gpu" not in selected:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_audio():
    return np.fromfile(os.path.join(GOLDEN, "ref_a_audio.bin"), "<f4")


@pytest.fixture(scope="session")
def golden_mel():
    return np.fromfile(os.path.join(GOLDEN, "ref_c_mel_numpy.bin"), "<f4").reshape(148, 80)


@pytest.fixture(scope="session")
def fb80():
    return np.fromfile(os.path.join(GOLDEN, "mel_80.bin"), "<f4").reshape(80, 201)


@pytest.fixture(scope="session")
def fb128():
    return np.fromfile(os.path.join(GOLDEN, "mel_128.bin"), "<f4").reshape(128, 201)
