import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """liblst_b200.so, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from laser_spot_track_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def oracle_clib():
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return os.path.join(ROOT, "oracle", "_build", "liblst_oracle_c.so")
