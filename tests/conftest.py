import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "mitsuba-path-guiding_b200")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def load_package():
    """Import the package directory (its name has a dash) under the module name ``b200pg``."""
    if "b200pg" in sys.modules:
        return sys.modules["b200pg"]
    spec = importlib.util.spec_from_file_location(
        "b200pg", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b200pg"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def oracle(pkg):
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    import oracle_lib

    return oracle_lib.Oracle(so)
