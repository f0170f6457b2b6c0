import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built libraries (they are git-ignored): build them once, as the driver's
    # __graft_entry__.build() does (nvcc cross-compiles sm_100a without a GPU)
    needed = [os.path.join(ROOT, "binary_b200", "libbinary_cuda.so"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in needed):
        import __graft_entry__
        __graft_entry__.build()


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
