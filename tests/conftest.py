import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port, have_port
    if not have_port():
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return Ref()
