import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The tests load the in-tree CUDA library (built artefacts are git-ignored): build it once when it is missing or
    older than its sources and a compiler is at hand (nvcc cross-compiles without a GPU)."""
    import shutil
    if shutil.which("nvcc") is None:
        return
    import __graft_entry__ as g
    try:
        g.build()
    except Exception as e:                      # the tests that need the library will say what is missing
        print("conftest: build() failed:", e)
