import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def state_bundle():
    """Seed-0 random-init weights with the reference's state_dict names (the goldens use seed 0)."""
    from mmdx_b200 import synth
    return synth.make_state_bundle(seed=0)


@pytest.fixture(scope="session")
def g1():
    return dict(np.load(os.path.join(GOLDEN, "g1_samples.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
