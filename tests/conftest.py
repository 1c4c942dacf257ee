"""Shared pytest configuration: markers, paths, fixture loaders."""
from __future__ import annotations

import pathlib
import sys

import numpy as np
import pytest

REPO = pathlib.Path(__file__).resolve().parent.parent
GOLDEN = REPO / "tests" / "golden"
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / name, allow_pickle=False) as data:
        return {k: data[k] for k in data.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_norm(a, b):
    """max|a-b| / max|b|  -- the norm-wise criterion of SURVEY.md section 4."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.max(np.abs(b))
    if scale == 0.0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / scale)
