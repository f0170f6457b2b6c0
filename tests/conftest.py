import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    import oracle
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libbinary_ref.so not built (needs /root/reference at build time)")
    return oracle.Oracle("reference")
