import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def orc():
    from oracle import binding
    if not os.path.exists(binding.ORC_PATH):
        binding.build()
    return binding.Oracle("orc")


@pytest.fixture(scope="session")
def ref():
    from oracle import binding
    if not binding.have_ref():
        pytest.skip("oracle/_ref/libcattle_ref.so not built (needs /root/reference)")
    return binding.Oracle("ref")
