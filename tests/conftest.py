import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


class Fixture:
    """A golden .npz with prefix helpers: fx.sub('sd/') -> {name: array}."""

    def __init__(self, name):
        self.name = name
        self.data = dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    def __getitem__(self, k):
        return self.data[k]

    def __contains__(self, k):
        return k in self.data

    def sub(self, prefix):
        return {k[len(prefix):]: v for k, v in self.data.items() if k.startswith(prefix)}

    @property
    def cfg(self):
        c = self.data["cfg"]
        keys = ["n_items", "n_users", "L", "D", "Z", "hidden", "phidden", "B", "no_user"]
        return dict(zip(keys, [int(v) for v in c]))


_cache = {}


def load_golden(name):
    if name not in _cache:
        _cache[name] = Fixture(name)
    return _cache[name]


@pytest.fixture(scope="session")
def golden():
    return load_golden
