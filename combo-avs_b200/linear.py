"""torch.nn.functional.linear on the tcgen05 tensor cores at float32 accuracy (3xTF32) -- row f2 of the scope table: the Linear
layers of the deformable encoder layer (reference msdeformattn.py:119-134, ops/modules/ms_deform_attn.py:105-128), which the reference
runs as strict-fp32 cuBLAS GEMMs.  C ABI: msda_linear_forward (include/msda_b200.h), kernel: csrc/msda_linear.cu.

linear(x, weight, bias) == F.linear(x, weight, bias) to float32-level accuracy (measured 6e-7 .. 8e-6 max-norm relative against float64,
cuBLAS fp32: 2e-7 .. 1e-6).  Forward, dgrad and wgrad all run on the tensor cores; only the bias gradient (a column sum) is PyTorch's."""
import ctypes

import torch
from torch.autograd import Function

from . import _lib
from .functions import _ptr


def supported(x, weight):
    """True when the tensor-core kernel takes this call: CUDA float32, K % 4 == 0, N % 16 == 0 for the forward AND the dgrad shape."""
    if not (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.is_cuda):
        return False
    n, k = weight.shape
    m = x.numel() // max(k, 1)
    lib = _lib.load()
    return bool(m > 0 and x.shape[-1] == k and lib.msda_linear_supported(m, n, k) and lib.msda_linear_supported(m, k, n))


def linear_forward(x2d, weight, bias=None, relu=False, tf32_only=False):
    """y[M, N] = x2d[M, K] @ weight[N, K]^T (+ bias) (optionally max(., 0)); contiguous CUDA float32 tensors."""
    m, k = x2d.shape
    n = weight.shape[0]
    y = torch.empty((m, n), dtype=torch.float32, device=x2d.device)
    dev = x2d.device.index if x2d.device.index is not None else torch.cuda.current_device()
    flags = (1 if relu else 0) | (2 if tf32_only else 0)
    rc = _lib.load().msda_linear_forward(_ptr(x2d), _ptr(weight), _ptr(bias) if bias is not None else None, _ptr(y), m, n, k, flags, dev,
                                         ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc)
    return y


def linear_wgrad(g2d, x2d):
    """grad_weight[N, K] = g2d[M, N]^T @ x2d[M, K] on the tensor cores, operands read in place (no transposes materialised)."""
    m, n = g2d.shape
    k = x2d.shape[1]
    gw = torch.empty((n, k), dtype=torch.float32, device=g2d.device)
    dev = g2d.device.index if g2d.device.index is not None else torch.cuda.current_device()
    rc = _lib.load().msda_linear_wgrad(_ptr(g2d), _ptr(x2d), _ptr(gw), m, n, k, dev, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc)
    return gw


class TensorCoreLinearFunction(Function):
    """y = [relu](x @ weight^T + bias) with all three GEMMs of the training step on the tensor cores:
    forward, dgrad (grad_x = g @ weight) and wgrad (grad_weight = g^T @ x, its reduction over the rows split over the SMs)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu=False):
        x2d = x.reshape(-1, x.shape[-1]).contiguous()
        y = linear_forward(x2d, weight.contiguous(), bias.contiguous() if bias is not None else None, relu=relu)
        ctx.save_for_backward(x2d, weight, y if relu else None)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, grad_y):
        x2d, weight, y = ctx.saved_tensors
        g2d = grad_y.reshape(-1, grad_y.shape[-1])
        if y is not None:  # ReLU': one vectorised pass (a bool mask + a type-promoting multiply are two, the second one unvectorised)
            g2d = torch.ops.aten.threshold_backward(g2d, y, 0.0)
        g2d = g2d.contiguous()
        grad_x = grad_w = grad_b = None
        if ctx.needs_input_grad[0]:  # dgrad: grad_x[M, K] = g[M, N] @ W[N, K] = linear(g, W^T)
            grad_x = linear_forward(g2d, weight.t().contiguous()).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:  # wgrad: grad_W[N, K] = g^T[N, M] @ x[M, K] = linear(g^T, x^T)
            if _lib.load().msda_linear_wgrad_supported(g2d.shape[0], g2d.shape[1], x2d.shape[1]):
                grad_w = linear_wgrad(g2d, x2d)
            else:
                grad_w = g2d.t() @ x2d
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from .layernorm import column_sum
            grad_b = column_sum(g2d)
        return grad_x, grad_w, grad_b, None


def linear(x, weight, bias=None, relu=False):
    """Drop-in for F.linear (optionally followed by ReLU): the tensor-core kernel when it applies (see supported()), else PyTorch."""
    if supported(x, weight):
        return TensorCoreLinearFunction.apply(x, weight, bias, relu)
    y = torch.nn.functional.linear(x, weight, bias)
    return torch.relu(y) if relu else y
