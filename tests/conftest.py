import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle_lib():
    import helpers
    return helpers.oracle_lib()


@pytest.fixture(scope="session")
def hostsim_lib():
    import helpers
    return helpers.hostsim_lib()


@pytest.fixture(scope="session")
def golden_cases(tmp_path_factory):
    import helpers
    base = tmp_path_factory.mktemp("golden")
    return {name: helpers.load_golden(name, str(base)) for name in helpers.GOLDEN_NAMES}
