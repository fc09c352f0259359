"""y = LayerNorm(a + b) and the bias-gradient column sum as single-pass HBM-bound kernels -- row f2 of the scope table, the elementwise
epilogue of the deformable encoder layer (reference models/modeling/pixel_decoder/msdeformattn.py:119-134: `src = norm1(src +
dropout1(src2))`, `src = norm2(src + dropout3(src2))`).  C ABI: msda_add_layernorm_forward / msda_layernorm_backward / msda_colsum
(include/msda_b200.h), kernels: csrc/msda_layernorm.cu.

add_layer_norm(a, b, weight, bias, eps) == F.layer_norm(a + b, (C,), weight, bias, eps) with gradients for a, b, weight and bias;
column_sum(g) == g.sum(0).  Both fall back to PyTorch when the kernels do not apply (CPU tensors, other dtypes, C not a multiple of 128)."""
import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from .functions import _ptr


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _dev(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def supported(a, b, weight, bias):
    """True when the kernels take this call: contiguous-able CUDA float32 rows of C in {128, 256, 384, 512} with affine parameters."""
    if weight is None or bias is None or not (a.is_cuda and a.dtype == torch.float32 and weight.dtype == torch.float32 and bias.dtype == torch.float32):
        return False
    if b is not None and (b.shape != a.shape or b.dtype != a.dtype or b.device != a.device):
        return False
    c = a.shape[-1]
    return bool(a.numel() > 0 and weight.numel() == c and bias.numel() == c and _lib.load().msda_layernorm_supported(c))


class AddLayerNormFunction(Function):
    @staticmethod
    def forward(ctx, a, b, weight, bias, eps):
        c = a.shape[-1]
        a2 = a.reshape(-1, c).contiguous()
        b2 = b.reshape(-1, c).contiguous() if b is not None else None
        m = a2.shape[0]
        y = torch.empty_like(a2)
        stats = torch.empty((2, m), dtype=torch.float32, device=a2.device)
        dev = _dev(a2)
        w, bt = weight.contiguous(), bias.contiguous()
        rc = _lib.load().msda_add_layernorm_forward(_ptr(a2), _ptr(b2) if b2 is not None else None, _ptr(w), _ptr(bt), _ptr(y), _ptr(stats[0]),
                                                    _ptr(stats[1]), m, c, float(eps), dev, _stream(dev))
        _lib.check(rc)
        ctx.save_for_backward(a2, b2, w, stats)
        ctx.shape = a.shape
        ctx.has_b = b is not None
        return y.view(a.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        a2, b2, w, stats = ctx.saved_tensors
        m, c = a2.shape
        g = grad_y.reshape(m, c).contiguous()
        dx = torch.empty_like(a2)
        dgb = torch.empty((2, c), dtype=torch.float32, device=a2.device)
        dev = _dev(a2)
        lib = _lib.load()
        ws_bytes = lib.msda_layernorm_backward_workspace_bytes(m, c, dev)
        ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.float32, device=a2.device)
        rc = lib.msda_layernorm_backward(_ptr(g), _ptr(a2), _ptr(b2) if b2 is not None else None, _ptr(w), _ptr(stats[0]), _ptr(stats[1]), _ptr(dx),
                                         _ptr(dgb[0]), _ptr(dgb[1]), _ptr(ws), ws_bytes, m, c, dev, _stream(dev))
        _lib.check(rc)
        dx = dx.view(ctx.shape)
        return dx, (dx if ctx.has_b else None), dgb[0], dgb[1], None


def add_layer_norm(a, b, weight, bias, eps=1e-5):
    """LayerNorm over the last dimension of a + b (b may be None)."""
    if supported(a, b, weight, bias):
        return AddLayerNormFunction.apply(a, b, weight, bias, eps)
    x = a if b is None else a + b
    return torch.nn.functional.layer_norm(x, (a.shape[-1],), weight, bias, eps)


def column_sum(g2d):
    """g2d[M, N].sum(0) (the bias gradient of a Linear layer) in one pass; PyTorch's reduction otherwise."""
    if g2d.is_cuda and g2d.dtype == torch.float32 and g2d.dim() == 2 and g2d.is_contiguous() and g2d.shape[0] > 0:
        lib = _lib.load()
        m, n = g2d.shape
        if lib.msda_colsum_supported(n):
            dev = _dev(g2d)
            out = torch.empty((n,), dtype=torch.float32, device=g2d.device)
            ws_bytes = lib.msda_colsum_workspace_bytes(m, n, dev)
            ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.float32, device=g2d.device)
            _lib.check(lib.msda_colsum(_ptr(g2d), _ptr(out), _ptr(ws), ws_bytes, m, n, dev, _stream(dev)))
            return out
    return g2d.sum(0)
