import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def emu_cdll():
    """The CPU functional simulator build of the kernels (tests only)."""
    import ctypes
    import subprocess

    from cpmd_b200 import lib

    d = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(d, "libcpb200_emu.so")
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", d, "-j", "8"], env=env, stdout=subprocess.DEVNULL)
    return lib.declare(ctypes.CDLL(so))
