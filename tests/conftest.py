import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PRODUCT_LIB = os.path.join(ROOT, "salviarenderer_b200", "csrc", "libsalvia_b200.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "libsalvia_oracle.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libsalvia_ref.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Builds the checkers (and the product, which cross-compiles without a GPU) once per session."""
    import __graft_entry__ as g
    g.build_oracle()
    g.build_product()
    g.build_reference()
    return g


@pytest.fixture(scope="session")
def oracle(built):
    from salviarenderer_b200 import abi
    return abi.Backend(ORACLE_LIB)


@pytest.fixture(scope="session")
def reference(built):
    from salviarenderer_b200 import abi
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libsalvia_ref.so not built (needs /root/reference)")
    return abi.Backend(REF_LIB)


@pytest.fixture(scope="session")
def cuda(built):
    """The product on cuda:0. Never falls back: a missing library or GPU is an error for a gpu test."""
    import salviarenderer_b200 as pkg
    return pkg.load(0)
