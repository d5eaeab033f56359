import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, (re)built in-tree if stale.  nvcc cross-compiles without a GPU."""
    from sbb_textline_detection_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def textline_weights():
    from sbb_textline_detection_b200.detector import synthetic_weights
    return synthetic_weights("textline")


@pytest.fixture(scope="session")
def region_weights():
    from sbb_textline_detection_b200.detector import synthetic_weights
    return synthetic_weights("region")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def iou(a, b, cls=1):
    inter = np.logical_and(a == cls, b == cls).sum()
    union = np.logical_or(a == cls, b == cls).sum()
    return inter / max(union, 1)
