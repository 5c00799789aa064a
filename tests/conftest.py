import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The tests load the in-tree CUDA library (built artefacts are git-ignored): build it once when it is MISSING and a
    compiler is at hand (nvcc cross-compiles without a GPU).  A library that is present is used as it is - file times
    do not survive the copy to the GPU box, so "older than its sources" means nothing there."""
    import shutil
    import __graft_entry__ as g
    if os.path.exists(g.LIB) or shutil.which("nvcc") is None:
        return
    try:
        g.build()
    except Exception as e:                      # the tests that need the library will say what is missing
        print("conftest: build() failed:", e)
