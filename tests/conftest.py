import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def has_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "models"))


needs_reference = pytest.mark.skipif(not has_reference(), reason="/root/reference not mounted (GPU box)")
