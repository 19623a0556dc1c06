import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build_c()
    return O


@pytest.fixture(scope="session")
def pkg():
    import psd_b200
    return psd_b200.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def make_clouds(kind, b, n, m, seed):
    """Seeded synthetic clouds: created on the CPU so that every implementation sees the
    same bits.  kinds: uniform U[0,1)^3, clustered (16 Gaussians clamped to [0,1]), lattice (integer grid /8:
    many exact distance ties), dup (targets drawn with repetition), offset (uniform + 100: large |coords|)."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        x = rng.random((b, n, 3), dtype=np.float32)
        y = rng.random((b, m, 3), dtype=np.float32)
    elif kind == "clustered":
        c = rng.random((b, 16, 3))
        def draw(k):
            ci = rng.integers(0, 16, size=(b, k))
            pts = np.take_along_axis(c, ci[..., None].repeat(3, -1), 1) + 0.03 * rng.standard_normal((b, k, 3))
            return np.clip(pts, 0, 1).astype(np.float32)
        x, y = draw(n), draw(m)
    elif kind == "lattice":
        x = (rng.integers(0, 9, size=(b, n, 3)) / 8.0).astype(np.float32)
        y = (rng.integers(0, 9, size=(b, m, 3)) / 8.0).astype(np.float32)
    elif kind == "dup":
        base = rng.random((b, max(m // 8, 1), 3), dtype=np.float32)
        sel = rng.integers(0, base.shape[1], size=(b, m))
        y = np.take_along_axis(base, sel[..., None].repeat(3, -1), 1)
        x = rng.random((b, n, 3), dtype=np.float32)
        k = min(n // 4, m)
        x[:, :k] = y[:, :k]  # a quarter of the queries coincide with (duplicated) targets
    elif kind == "offset":
        x = (rng.random((b, n, 3), dtype=np.float32) + np.float32(100.0)).astype(np.float32)
        y = (rng.random((b, m, 3), dtype=np.float32) + np.float32(100.0)).astype(np.float32)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
