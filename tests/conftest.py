import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        return json.load(f)


GOLDEN_SETS = ["cfg1_small", "cfg2_small", "cfg3_small", "cfg4_small", "cfg5_small", "cfg5_small_fast", "crafted"]


@pytest.fixture(scope="session")
def oracle():
    from oracle import nr_oracle
    nr_oracle.build()
    return nr_oracle


@pytest.fixture(scope="session")
def engine():
    """The CUDA library; GPU tests fail loudly (no skip, no fallback) when it is missing."""
    from nanorepeat_b200 import engine as e
    e.lib()
    e.init(0)
    return e
