import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), built on demand."""
    from oracle import pyoracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))


@pytest.fixture(scope="session")
def cuda_lib_path():
    """librptr_cuda.so, built in-tree on demand (nvcc cross-compiles without a GPU)."""
    from realtimepathtracingresearchframework_b200 import build
    if build.needs_build():
        build.build()
    return build.LIB


@pytest.fixture(scope="session")
def hostsim():
    so = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostsim")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so
