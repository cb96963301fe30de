import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the in-tree libraries once (CUDA library needs nvcc, not a GPU)."""
    import __graft_entry__ as g
    need = [os.path.join(ROOT, "uhsdr_b200", "csrc", "libuhsdr_b200.so"),
            os.path.join(ROOT, "uhsdr_b200", "csrc", "libuhsdr_b200_exact.so"),
            os.path.join(ROOT, "oracle", "liboracle_port.so")]
    if not all(os.path.exists(p) for p in need):
        g.build()
    return True


def oracle_channel(cfg):
    """The strongest oracle available: the compiled reference if oracle/_ref travelled with the
    snapshot, else the plain-C port (which is pinned against it by test_oracle_pin.py)."""
    from oracle import refchain
    from oracle.port import PortChannel
    return refchain.RefChannel(cfg) if refchain.available() else PortChannel(cfg)
