import pathlib
import sys

import pytest

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs oracle/_ref/ref_harness (the compiled unmodified reference)")


@pytest.fixture(scope="session")
def orc():
    import kit
    return kit.oracle()


@pytest.fixture(scope="session")
def rbc():
    import kit
    return kit.load_rbc_fixture()
