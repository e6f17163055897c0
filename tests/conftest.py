import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: longer CPU test")


@pytest.fixture(scope="session")
def reference_results():
    with open(os.path.join(GOLDEN, "reference_results.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def dense_results():
    with open(os.path.join(GOLDEN, "dense_results.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def built_library():
    """The C-ABI library, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    from interiorpointmethod_b200 import build

    return build.build()
