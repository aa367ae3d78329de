"""Shared test plumbing.

``-m "not gpu"`` (CPU container): oracle vs golden vectors, host logic, C-ABI
symbol exports.  ``-m gpu`` (B200 box): CUDA path vs oracle / golden through
the C-ABI.  Nothing here reads /root/reference at run time.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG_NAME = "multimodal-sensor-fusion-with-attention-rajeevatla_b200"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_pkg():
    """The product package (its directory name has hyphens, hence importlib)."""
    return importlib.import_module(PKG_NAME)


def dropin_src():
    """Directory holding the drop-in ``fusion`` / ``attention`` / ``encoders`` /
    ``uncertainty`` modules; the reference's tests put it first on sys.path."""
    return os.path.join(ROOT, PKG_NAME, "src")


class Golden:
    """A tests/golden/*.npz fixture with '/'-separated groups."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    def __getitem__(self, key):
        return self.z[key]

    def __contains__(self, key):
        return key in self.z.files

    def t(self, key, dtype=None):
        a = torch.from_numpy(np.array(self.z[key]))
        return a if dtype is None else a.to(dtype)

    def group(self, prefix, dtype=None):
        pre = prefix + "/"
        return {k[len(pre):]: self.t(k, dtype) for k in self.z.files if k.startswith(pre)}

    @property
    def names(self):
        return [str(s) for s in self.z["names"]]


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()
