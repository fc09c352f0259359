"""ctypes/numpy front-end of oracle/msda_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

forward()/backward() take numpy arrays laid out exactly as the reference's CUDA entry points take
tensors (ops/src/cuda/ms_deform_attn_cuda.cu:25-31, 88-95): value[N,S,M,D], spatial_shapes[L,2] int64
(H,W), level_start_index[L] int64, sampling_loc[N,Lq,M,L,P,2] (x,y), attn_weight[N,Lq,M,L,P].
dtype float64 -> the _f64 instantiation, float32 -> _f32.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmsda_oracle.so")
_lib = None


def build(force=False):
    """Compile libmsda_oracle.so with gcc if it is missing or older than its source."""
    src = os.path.join(_HERE, "msda_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        r = subprocess.run(["make", "-s", "-C", _HERE, "-B", "all"], capture_output=True, text=True)
        if r.returncode != 0:  # no libgomp on this host: single-threaded oracle
            subprocess.run(["make", "-s", "-C", _HERE, "-B", "all", "OMP="], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        p = ctypes.c_void_p
        i = ctypes.c_int
        for sfx in ("f64", "f32"):
            f = getattr(lib, "msda_oracle_forward_" + sfx)
            f.argtypes = [p, p, p, p, p, p] + [i] * 7
            f.restype = i
            b = getattr(lib, "msda_oracle_backward_" + sfx)
            b.argtypes = [p, p, p, p, p, p, p, p, p] + [i] * 7
            b.restype = i
        lib.msda_oracle_max_threads.restype = i
        lib.msda_oracle_set_threads.argtypes = [i]
        _lib = lib
    return _lib


def max_threads():
    return _load().msda_oracle_max_threads()


def set_threads(n):
    _load().msda_oracle_set_threads(int(n))


def _prep(value, shapes, start, loc, attn):
    dt = value.dtype
    if dt not in (np.float64, np.float32):
        raise TypeError("oracle supports float32/float64, got %s" % dt)
    value = np.ascontiguousarray(value, dtype=dt)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    start = np.ascontiguousarray(start, dtype=np.int64)
    N, S, M, D = value.shape
    L = shapes.shape[0]
    Lq, P = loc.shape[1], loc.shape[4]
    assert loc.shape == (N, Lq, M, L, P, 2), loc.shape
    assert attn.shape == (N, Lq, M, L, P), attn.shape
    assert int((shapes[:, 0] * shapes[:, 1]).sum()) == S
    sfx = "f64" if dt == np.float64 else "f32"
    return value, shapes, start, loc, attn, (N, S, M, D, L, Lq, P), sfx


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def forward(value, shapes, start, loc, attn):
    """-> out[N, Lq, M*D] (same dtype as value)."""
    value, shapes, start, loc, attn, dims, sfx = _prep(value, shapes, start, loc, attn)
    N, S, M, D, L, Lq, P = dims
    out = np.empty((N, Lq, M * D), dtype=value.dtype)
    rc = getattr(_load(), "msda_oracle_forward_" + sfx)(
        _ptr(value), _ptr(shapes), _ptr(start), _ptr(loc), _ptr(attn), _ptr(out), *dims)
    if rc != 0:
        raise ValueError("msda_oracle_forward: bad dimensions")
    return out


def backward(value, shapes, start, loc, attn, grad_out):
    """-> (grad_value, grad_loc, grad_attn) with the shapes of value / loc / attn."""
    value, shapes, start, loc, attn, dims, sfx = _prep(value, shapes, start, loc, attn)
    N, S, M, D, L, Lq, P = dims
    grad_out = np.ascontiguousarray(grad_out, dtype=value.dtype).reshape(N, Lq, M * D)
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    rc = getattr(_load(), "msda_oracle_backward_" + sfx)(
        _ptr(value), _ptr(shapes), _ptr(start), _ptr(loc), _ptr(attn), _ptr(grad_out),
        _ptr(gv), _ptr(gl), _ptr(ga), *dims)
    if rc != 0:
        raise ValueError("msda_oracle_backward: bad dimensions")
    return gv, gl, ga
