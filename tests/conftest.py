"""pytest config: registers the `gpu` marker; everything unmarked must pass on a CPU-only box."""
import glob
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "golden_n*_s*.json")))


def load_golden(path):
    with open(path) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle, build
    build(ref=True)
    return Oracle()


def fhex(s):
    return float.fromhex(s)
