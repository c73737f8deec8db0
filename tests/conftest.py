"""Shared pytest configuration.

``-m "not gpu"``: oracle vs the committed golden vectors, host logic, and that the C-ABI library
loads and exports every symbol ``include/tvbf.h`` declares (no compute calls without a GPU).
``-m gpu``: the parity tests proper -- they call the sm_100a kernels through the C-ABI.
"""

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def load_golden(name):
    from oracle.make_golden import unpack_catalogue

    z = np.load(GOLDEN / f"{name}.npz")
    return z, unpack_catalogue(z)


# ---- the reference's own fixtures, restated (tests/conftest.py:137-206 of the reference) -------
@pytest.fixture
def sample_genre_features():
    return np.array([[1, 1, 1, 0, 0], [1, 1, 0, 0, 0], [0, 0, 0, 1, 0]], dtype=float)


@pytest.fixture
def sample_text_features():
    from scipy.sparse import csr_matrix

    data = np.array([0.5, 0.3, 0.7, 0.4, 0.6, 0.8])
    row = np.array([0, 0, 1, 1, 2, 2])
    col = np.array([0, 3, 1, 4, 2, 5])
    return csr_matrix((data, (row, col)), shape=(3, 10))


@pytest.fixture
def sample_platform_features():
    return np.array([[1, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=float)


@pytest.fixture
def sample_type_features():
    return np.array([[1, 0], [1, 0], [1, 0]], dtype=float)


@pytest.fixture
def sample_language_features():
    return np.array([[1, 0], [1, 0], [1, 0]], dtype=float)


@pytest.fixture
def sample_similarity_matrix():
    return np.array([[1.0, 0.8, 0.2], [0.8, 1.0, 0.3], [0.2, 0.3, 1.0]])
