"""pytest configuration: the `gpu` marker and shared helpers.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors / compiled reference, host logic, C-ABI exports.
`-m gpu` runs on a B200: the CUDA path (through the C ABI) vs the oracle and the golden vectors.
Nothing marked gpu reads /root/reference (it does not exist on the GPU box)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def built():
    """Makes sure libmag.so and the restated oracle exist AND that libmag.so was built from the sources as they are now
    (a hash of the sources is stamped at build time: file times do not survive the copy to the GPU box, hashes do)."""
    import __graft_entry__ as g
    from core_b200._lib import LIB_PATH
    from oracle import mao
    stale = True
    try:
        stale = open(g.STAMP).read().strip() != g.source_hash()
    except OSError:
        pass
    if stale or not os.path.exists(LIB_PATH) or not os.path.exists(mao.LIB_PATH):
        g.build()
    return True
