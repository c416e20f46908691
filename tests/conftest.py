import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

PRM = (64, 150, 1, 0)  # default deterministic-schedule knobs (include/gds.h gds_params)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def O():
    import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def R():
    import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    pyref.lib()
    return pyref


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope="session")
def solver(pkg):
    s = pkg.Solver(0)  # raises without a GPU: no CPU fallback
    yield s
    s.close()


@pytest.fixture(scope="session")
def golden():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
