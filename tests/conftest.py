"""pytest configuration: the ``gpu`` marker and import paths.

``-m "not gpu"`` tests run in the authoring container (no GPU): the oracle against the golden
fixtures, host-side logic, world_size-2 gloo paths and the C-ABI export check.
``-m gpu`` tests are the parity tests proper: they call the CUDA kernels through the C ABI and
compare with the oracle / fixtures on a B200.
"""
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import torch

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = torch.load(GOLDEN / name, weights_only=False, map_location="cpu")
        return cache[name]

    return load


@pytest.fixture(scope="session")
def cuda_dev():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("CUDA device required")
    return torch.device("cuda", 0)
