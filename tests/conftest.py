import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    ge.build()  # no-op when up to date; compiles with nvcc otherwise
    return ge.load_package()


@pytest.fixture(scope="session")
def repo_dir():
    import __graft_entry__ as ge
    ge.ensure_fixtures()
    return os.path.join(ROOT, "models")


@pytest.fixture(scope="session")
def densenet_path(repo_dir):
    return os.path.join(repo_dir, "densenet_onnx", "1", "model.onnx")
