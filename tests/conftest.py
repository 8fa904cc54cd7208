import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    """Both shared libraries exist (built in-tree by __graft_entry__.build())."""
    import __graft_entry__ as g
    from sosba_loader import load_package
    pkg = load_package()
    orc = os.path.join(ROOT, "oracle", "_build", "liborc_parity.so")
    if not (os.path.exists(pkg.LIB_PATH) and os.path.exists(orc)):
        g.build()
    return pkg


@pytest.fixture(scope="session")
def orc(built):
    from sos_slam_b200 import binding
    return binding.Lib(os.path.join(ROOT, "oracle", "_build", "liborc_parity.so"), "orc")


@pytest.fixture(scope="session")
def gpu(built):
    return built.load()
