import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "acg-alp-ldpc_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/ldpc_oracle.c) -- the checker."""
    from oracle.oracle import Oracle, build
    build(ref=True)
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref), when it was built in the dev container."""
    from oracle.oracle import Ref, have_ref, build
    build(ref=True)
    if not have_ref():
        pytest.skip("oracle/_ref/libref_oracle.so not built (no /root/reference here)")
    return Ref()


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; GPU tests must run the CUDA path or fail loudly."""
    import ldpc_b200
    ldpc_b200.lib()
    assert ldpc_b200.device_count() >= 1, "no CUDA device: there is no CPU fallback"
    return ldpc_b200
