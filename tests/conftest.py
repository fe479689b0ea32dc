"""Shared pytest configuration.

Markers
-------
gpu : test needs a CUDA device (run with ``-m gpu`` on the B200 box).
Everything else runs on CPU only (``-m "not gpu"``).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


def rel_linf(x, ref):
    """max|x-ref| / max|ref| with NaNs required to coincide."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    nx, nr = np.isnan(x), np.isnan(ref)
    assert np.array_equal(nx, nr), "NaN pattern differs"
    d = np.abs(np.where(nr, 0.0, x - ref))
    scale = np.max(np.abs(np.where(nr, 0.0, ref)))
    return float(d.max() / scale) if scale > 0 else float(d.max())
