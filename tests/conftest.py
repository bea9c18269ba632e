"""pytest configuration: the `gpu` marker and shared helpers.

`-m "not gpu"` covers the oracle against the golden vectors, the host logic and the C-ABI
symbol check (no compute calls); `-m gpu` tests are the parity tests proper: they call the
CUDA path through the C ABI and compare with oracle/ on the same seeded inputs.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "golden_firstparty.npz")))


@pytest.fixture(scope="session")
def model_cfgs():
    import json
    with open(os.path.join(GOLDEN, "configs.json")) as f:
        return json.load(f)
