"""pytest configuration: the ``gpu`` marker and shared fixtures.

CPU tests (no marker) must pass without a GPU; tests marked ``gpu`` run on a real B200 and fail
loudly if the CUDA library cannot be loaded there (no silent PyTorch fallback exists).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def rng():
    return np.random.default_rng(1234)


def random_image(rng, h, w, c):
    shape = (h, w) if c == 1 else (h, w, c)
    return rng.integers(0, 256, shape, dtype=np.uint8)
