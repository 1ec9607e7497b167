import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def _native_library_is_current():
    """The ABI / SASS tests (CPU) and every GPU test load librloa_b200.so from the tree; it is git-ignored, so a fresh
    checkout (or an edited kernel) needs a build first.  nvcc cross-compiles without a GPU; without nvcc the tests that
    need the library fail loudly on their own."""
    from robotic_manipulator_rloa_b200 import build_native
    if build_native._stale() and os.path.isfile(build_native.NVCC):
        build_native.build()
    yield


@pytest.fixture(autouse=True)
def _tmp_cwd(tmp_path, monkeypatch):
    """The package (like the reference) writes training_logs.log / checkpoints/ into the CWD."""
    monkeypatch.chdir(tmp_path)
    yield
