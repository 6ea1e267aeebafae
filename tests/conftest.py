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


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) where no device is visible, e.g. the build container."""
    try:
        import torch

        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def highway_config():
    from highway_rope_ppo_b200.config.base_config import HIGHWAY_CONFIG

    return HIGHWAY_CONFIG


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import highway

    highway.build()
