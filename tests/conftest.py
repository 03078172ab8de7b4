import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ffi():
    """ctypes bindings; builds the host front end and the oracle twin if they are missing."""
    from parity_util import load_ffi
    mod = load_ffi()
    if not os.path.exists(mod.HOST_LIB):
        subprocess.run(["make", "-j8"], cwd=os.path.join(ROOT, "pbrt-v1_b200", "host"), check=True)
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libpbrt_twin.so")) or not os.path.exists(
            os.path.join(ROOT, "oracle", "_build", "libpbrt_twin_libm.so")):
        subprocess.run(["make", "-j2"], cwd=os.path.join(ROOT, "oracle"), check=True)
    return mod


@pytest.fixture(scope="session")
def scene_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("scenes"))
