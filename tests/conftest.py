import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def clustered(n, d, n_centres=32, seed=0, scale=2.0, dtype=np.float32):
    """Seeded clustered data: neighbourhood structure like real embeddings."""
    cent = np.random.default_rng(777 + d + n_centres).standard_normal((n_centres, d)).astype(np.float32) * scale
    rng = np.random.default_rng(seed)   # centres are shared by base sets and queries
    x = cent[rng.integers(0, n_centres, n)] + rng.standard_normal((n, d)).astype(np.float32)
    return x.astype(dtype)


def sift_like(n, d=128, seed=0):
    """SIFT-shaped: small non-negative integers stored as fp32 -> exact distance ties are common."""
    cent = np.random.default_rng(555 + d).integers(0, 120, (64, d))
    rng = np.random.default_rng(seed)
    x = cent[rng.integers(0, 64, n)] + rng.integers(-12, 13, (n, d))
    return np.clip(x, 0, 255).astype(np.float32)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def pkg():
    import pgvector_hnsw_partitioning_b200 as p
    return p


def has_gpu():
    try:
        import pgvector_hnsw_partitioning_b200 as p
        return p.load_library().hb_device_count() > 0
    except Exception:
        return False
