import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running")


@pytest.fixture(scope="session")
def golden():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "sessions.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_wrappers():
    """412 games of the reference's own wrapper stacks (oracle/make_golden.py --wrappers)."""
    import json

    with open(os.path.join(ROOT, "tests", "golden", "wrappers.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the CUDA library; GPU tests must run the native path."""
    import importlib

    build = importlib.import_module("pikazoo_b200.build")
    build.build()
    import pikazoo_b200

    return pikazoo_b200.load_library()
