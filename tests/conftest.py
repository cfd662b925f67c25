import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu)')


def load_pkg():
    """Import the product package (directory name is not a valid identifier) as `pgpp_b200`."""
    if 'pgpp_b200' in sys.modules:
        return sys.modules['pgpp_b200']
    pkg_dir = os.path.join(ROOT, 'pasta-gan-plusplus_b200')
    spec = importlib.util.spec_from_file_location('pgpp_b200', os.path.join(pkg_dir, '__init__.py'),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules['pgpp_b200'] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope='session')
def pkg():
    return load_pkg()


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
