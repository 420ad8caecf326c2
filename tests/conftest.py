import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


CONFIG_KEYS_ISO = ["Al", "CH2", "H2O", "YAG"]


@pytest.fixture(scope="session")
def configs():
    from __graft_entry__ import CONFIGS
    return CONFIGS


def golden(key):
    import numpy as np
    return np.load(os.path.join(HERE, "golden", "iso_%s.npz" % key), allow_pickle=False)
