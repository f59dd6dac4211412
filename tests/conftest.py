import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout; the SIMT tests deadlock on a malformed kernel)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Checkers (oracle, twins, the kernel sources over the SIMT shim) are built on demand; the product .so is
    built by __graft_entry__.build() and travels with the snapshot."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "twin")])
    # (not fatal here: only tests/test_simt.py needs these, and it fails on its own if they are missing)
    subprocess.call(["make", "-s", "-C", os.path.join(ROOT, "tests", "simt")])
    yield
