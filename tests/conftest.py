import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gpu_lib():
    """Build (if stale) and load the CUDA extension; GPU tests fail loudly if it is missing."""
    from protein_gibbs_sampler_b200 import _lib
    return _lib.load()
