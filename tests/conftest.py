import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle on the CPU oracle (test infrastructure; built on demand)."""
    import oracle_binding
    return oracle_binding.load()


@pytest.fixture(scope="session")
def ref_bin():
    """The unmodified reference CLI compiled by oracle/Makefile, or None where /root/reference is absent and no prebuilt binary travelled."""
    p = os.path.join(ROOT, "oracle", "_ref", "fastF_ref")
    if not os.path.exists(p) and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=False)
    return p if os.path.exists(p) else None


@pytest.fixture(scope="session")
def synth():
    import synth_binding
    return synth_binding.load()


@pytest.fixture(scope="session")
def gpu_ctx():
    from fastf_b200 import _lib
    ctx = _lib.Context(0)   # raises without a CUDA device: -m gpu tests never fall back
    yield ctx
    ctx.close()
