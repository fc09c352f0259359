import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

S4_SHAPES = [(7, 7), (14, 14), (28, 28)]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def level_start(shapes):
    a = np.asarray(shapes, dtype=np.int64)
    sizes = a[:, 0] * a[:, 1]
    return np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)


def make_inputs(N, shapes, M, D, Lq, P, seed=0, lo=0.0, hi=1.0, dtype=np.float32, normal=False):
    """Inputs with the distributions of the reference's test.py:36-39 (value = U*0.01, loc ~ U[lo,hi), attn normalised)."""
    rng = np.random.default_rng(seed)
    shapes = np.asarray(shapes, dtype=np.int64)
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    L = len(shapes)
    if normal:
        value = rng.standard_normal((N, S, M, D))
        attn = rng.standard_normal((N, Lq, M, L, P))
    else:
        value = rng.random((N, S, M, D)) * 0.01
        attn = rng.random((N, Lq, M, L, P)) + 1e-5
        attn /= attn.sum(axis=(-1, -2), keepdims=True)
    loc = rng.random((N, Lq, M, L, P, 2)) * (hi - lo) + lo
    grad_out = rng.random((N, Lq, M * D))
    return dict(value=value.astype(dtype), shapes=shapes, start=level_start(shapes), loc=loc.astype(dtype),
                attn=attn.astype(dtype), grad_out=grad_out.astype(dtype))


def rel_err_elementwise(a, b):
    """max over the elements of |a-b| / |b| -- the strict reading of "relative error"; only meaningful where b has no zero crossings
    (the forward output on the reference's test distribution: positive values, positive weights)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1e-300)).max())


def close_elementwise(a, b, tol):
    """|a-b| <= tol * (|b| + rms(b)) for every element: relative where the element is large, absolute at the scale of the tensor
    where it crosses zero (gradients)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return a.size == 0
    rms = float(np.sqrt(np.mean(b * b)))
    return bool((np.abs(a - b) <= tol * (np.abs(b) + rms)).all())


def rel_err(a, b):
    """max-norm relative error: max|a-b| / max|b| (elementwise relative error is meaningless near zero crossings)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
