import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the oracle is test infrastructure: only tests/, smoke() and bench.py's cpu legs import it
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library and the C oracle are build artefacts (git-ignored): build them once if a clean
    checkout has not run __graft_entry__.build() yet. nvcc cross-compiles without a GPU."""
    lib = os.path.join(ROOT, "ark_bulletproofs_b200", "libbp_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()
    yield
