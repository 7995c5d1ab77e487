import os
import sys

import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "openke-putranse_b200"))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def wn18_dir(tmp_path_factory):
    import util
    return util.materialize_wn18(str(tmp_path_factory.mktemp("wn18")))


@pytest.fixture(scope="session")
def golden():
    import util
    return util.Golden()
