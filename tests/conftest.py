import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The oracle is test infrastructure: build it once per session if it is not there."""
    so = os.path.join(ROOT, "oracle", "liborb_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liborb_oracle.so"])
    # the product library (nvcc cross-compiles without a GPU); normally built by __graft_entry__.build()
    from multiagent_orb_slam2_b200 import build as b
    if not os.path.exists(b.LIB):
        b.build()
    return so
