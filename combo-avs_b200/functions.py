"""Drop-in replacements for the reference's extension entry points and autograd function.

Reference interface (paths under /root/reference/models/modeling/pixel_decoder/ops/):
  * module `MultiScaleDeformableAttention`: ms_deform_attn_forward / ms_deform_attn_backward  (src/vision.cpp:18-21,
    host checks in src/cuda/ms_deform_attn_cuda.cu:25-85, 88-158)
  * class MSDeformAttnFunction(torch.autograd.Function)                                         (functions/ms_deform_attn_func.py:32-50)
Same names, argument order, shapes, dtypes of the results and error behaviour; the work is done by the sm_100a
kernels in libmsda_b200.so through its C ABI (include/msda_b200.h).  There is no CPU path and no PyTorch fallback:
non-CUDA tensors raise, exactly as the reference's extension does (src/ms_deform_attn.h:43,65).
"""
import ctypes
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float64: _lib.F64}

# host mirrors of (spatial_shapes, level_start_index), keyed by tensor identity + version counter, so that the
# one device->host read happens once per geometry tensor instead of once per call (6 layers share one tensor).
_geometry_cache = {}


def _version(t):
    # inference tensors (created under torch.inference_mode()) have no version counter: in-place changes to them cannot be seen,
    # so their host mirror is re-read on every call instead of cached (-1 never hits)
    return -1 if t.is_inference() else t._version


def _host_geometry(shapes, start):
    key = (id(shapes), id(start))
    hit = _geometry_cache.get(key)
    version = (_version(shapes), _version(start))
    if hit is not None and hit[0]() is shapes and hit[1]() is start and hit[2] == version and -1 not in version:
        return hit[3]
    hs = shapes.detach().to("cpu", torch.int64).contiguous()
    hl = start.detach().to("cpu", torch.int64).contiguous()
    geo = ((ctypes.c_int64 * hs.numel())(*hs.flatten().tolist()), (ctypes.c_int64 * hl.numel())(*hl.tolist()),
           tuple(map(tuple, hs.tolist())), tuple(hl.tolist()))

    def _drop(_ref, key=key):
        _geometry_cache.pop(key, None)

    _geometry_cache[key] = (weakref.ref(shapes, _drop), weakref.ref(start, _drop), version, geo)
    return geo


def _check_inputs(named, value):
    # ms_deform_attn_cuda.cu:33-43 / 98-110: contiguity and device asserts, in the reference's words
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError("%s tensor has to be contiguous" % name)
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")  # ms_deform_attn.h:43
    for name, t in named:
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor" % name)
        if t.device != value.device:
            raise RuntimeError("%s is on %s but value is on %s" % (name, t.device, value.device))


def _problem(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    if value.dim() != 4 or sampling_loc.dim() != 6 or attn_weight.dim() != 5 or spatial_shapes.dim() != 2:
        raise RuntimeError("expected value[N,S,M,D], spatial_shapes[L,2], sampling_loc[N,Lq,M,L,P,2], attn_weight[N,Lq,M,L,P]")
    if value.dtype not in _DTYPES:
        raise RuntimeError('"ms_deform_attn_cuda" not implemented for \'%s\'' % str(value.dtype).replace("torch.", ""))
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64 tensors")
    if attn_weight.dtype != sampling_loc.dtype:
        raise RuntimeError("sampling_loc and attn_weight must have the same dtype")
    if sampling_loc.dtype != value.dtype and not (value.dtype == torch.bfloat16 and sampling_loc.dtype == torch.float32):
        raise RuntimeError("sampling_loc/attn_weight dtype %s does not match value dtype %s" % (sampling_loc.dtype, value.dtype))
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    if tuple(sampling_loc.shape) != (N, Lq, M, L, P, 2) or tuple(attn_weight.shape) != (N, Lq, M, L, P):
        raise RuntimeError("inconsistent shapes: value %s, sampling_loc %s, attn_weight %s, %d levels" %
                           (tuple(value.shape), tuple(sampling_loc.shape), tuple(attn_weight.shape), L))
    if level_start_index.numel() != L or spatial_shapes.shape[1] != 2:
        raise RuntimeError("spatial_shapes must be [L,2] and level_start_index [L]")
    return (N, S, M, D, L, Lq, P), _DTYPES[value.dtype], _DTYPES[sampling_loc.dtype]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else None


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step, _geometry=None):
    """== MultiScaleDeformableAttention.ms_deform_attn_forward (vision.cpp:19). Returns output[N, Lq, M*D]."""
    _check_inputs((("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)), value)
    dims, dt, ldt = _problem(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    N, S, M, D, L, Lq, P = dims
    hs, hl = (_geometry or _host_geometry(spatial_shapes, level_start_index))[:2]
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    dev = value.device.index if value.device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = _lib.load().msda_forward(dt, ldt, _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_loc),
                                  _ptr(attn_weight), _ptr(out), N, S, M, D, L, Lq, P, int(im2col_step), hs, hl, dev,
                                  ctypes.c_void_p(stream))
    _lib.check(rc)
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step,
                            _geometry=None):
    """== MultiScaleDeformableAttention.ms_deform_attn_backward (vision.cpp:20).
    Returns [grad_value, grad_sampling_loc, grad_attn_weight]."""
    _check_inputs((("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)), value)
    dims, dt, ldt = _problem(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    N, S, M, D, L, Lq, P = dims
    if grad_output.dtype != value.dtype or grad_output.numel() != N * Lq * M * D:
        raise RuntimeError("grad_output must be %s with %d elements" % (value.dtype, N * Lq * M * D))
    hs, hl = (_geometry or _host_geometry(spatial_shapes, level_start_index))[:2]
    grad_value = torch.empty_like(value)
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    lib = _lib.load()
    ws_bytes = lib.msda_backward_workspace_bytes(dt, ldt, N, S, M, D, L, Lq, P)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=value.device) if ws_bytes else None
    dev = value.device.index if value.device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.msda_backward(dt, ldt, _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_loc),
                           _ptr(attn_weight), _ptr(grad_output), _ptr(grad_value), _ptr(grad_loc), _ptr(grad_attn),
                           N, S, M, D, L, Lq, P, int(im2col_step), hs, hl, _ptr(ws), ws_bytes, dev, ctypes.c_void_p(stream))
    _lib.check(rc)
    return [grad_value, grad_loc, grad_attn]


class MSDeformAttnFunction(Function):
    """Same contract as the reference class (functions/ms_deform_attn_func.py:32-50):
    apply(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step)
    -> output[N, Lq, M*D]; backward returns (grad_value, None, None, grad_sampling_loc, grad_attn_weight, None)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.geometry = _host_geometry(value_spatial_shapes, value_level_start_index) if value.is_cuda else None
        output = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                        attention_weights, ctx.im2col_step, ctx.geometry)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = ms_deform_attn_backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            grad_output.contiguous(), ctx.im2col_step, ctx.geometry)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


# ------------------------------------------------------------------------------------------------ fused module prologue
def fused_prologue_supported(value, spatial_shapes, level_start_index, reference_points, sampling_offsets):
    """True when msda_fused_forward / msda_fused_backward (include/msda_b200.h) can take this call: CUDA float32 tensors,
    2-d reference points that need no gradient, and a shape the slab / sorted kernels hold in shared memory."""
    if not (value.is_cuda and value.dtype == torch.float32 and sampling_offsets.dtype == torch.float32):
        return False
    if reference_points.shape[-1] != 2 or reference_points.requires_grad or reference_points.dtype != torch.float32:
        return False
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_offsets.shape[1], sampling_offsets.shape[4]
    # the reference's unfused arithmetic broadcasts reference points (e.g. [1, Lq, L, 2]); the kernels index them per frame
    if tuple(reference_points.shape) != (N, Lq, L, 2):
        return False
    hs, hl = _host_geometry(spatial_shapes, level_start_index)[:2]
    dev = value.device.index if value.device.index is not None else torch.cuda.current_device()
    return bool(_lib.load().msda_fused_supported(_lib.F32, N, S, M, D, L, Lq, P, hs, hl, dev))


class MSDeformAttnPrologueFunction(Function):
    """The op with the module's prologue folded in (ops/modules/ms_deform_attn.py:102-116 + functions/ms_deform_attn_func.py:32-50):
    apply(value[N,S,M,D], spatial_shapes, level_start_index, reference_points[N,Lq,L,2], sampling_offsets[N,Lq,M,L,P,2],
          attention_logits[N,Lq,M,L*P], im2col_step) == MSDeformAttnFunction.apply(value, shapes, start,
          reference_points[:, :, None, :, None, :] + sampling_offsets / (W_l, H_l), softmax(attention_logits, -1), im2col_step).
    An optional trailing padding_mask [N, S] (bool) stands for value.masked_fill(mask[..., None], 0) before the op.
    backward returns (grad_value, None, None, None, grad_sampling_offsets, grad_attention_logits, None, None)."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits, im2col_step,
                padding_mask=None):
        named = (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                 ("reference_points", reference_points), ("sampling_offsets", sampling_offsets), ("attention_logits", attention_logits))
        _check_inputs(named, value)
        N, S, M, D = value.shape
        L = spatial_shapes.shape[0]
        Lq, P = sampling_offsets.shape[1], sampling_offsets.shape[4]
        if tuple(sampling_offsets.shape) != (N, Lq, M, L, P, 2) or attention_logits.numel() != N * Lq * M * L * P or \
                tuple(reference_points.shape) != (N, Lq, L, 2):
            raise RuntimeError("inconsistent shapes: value %s, reference_points %s, sampling_offsets %s, attention_logits %s" %
                               (tuple(value.shape), tuple(reference_points.shape), tuple(sampling_offsets.shape), tuple(attention_logits.shape)))
        if not (value.dtype == sampling_offsets.dtype == attention_logits.dtype == reference_points.dtype == torch.float32):
            raise RuntimeError("the fused prologue takes float32 tensors")
        if padding_mask is not None:
            # replaces value.masked_fill(mask[..., None], 0) of the module (ms_deform_attn.py:103-104): `value` comes in unmasked
            if padding_mask.dtype != torch.bool or tuple(padding_mask.shape) != (N, S) or padding_mask.device != value.device:
                raise RuntimeError("padding_mask must be a bool tensor of shape [N, S] on the device of value")
            padding_mask = padding_mask.contiguous()
        ctx.im2col_step = im2col_step
        ctx.geometry = _host_geometry(spatial_shapes, level_start_index)
        hs, hl = ctx.geometry[:2]
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        attn = torch.empty((N, Lq, M, L, P), dtype=value.dtype, device=value.device)
        dev = value.device.index if value.device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = _lib.load().msda_fused_forward(_lib.F32, _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(reference_points),
                                            _ptr(sampling_offsets), _ptr(attention_logits), _ptr(padding_mask), _ptr(out), _ptr(attn), N, S, M, D, L, Lq, P,
                                            int(im2col_step), hs, hl, dev, ctypes.c_void_p(stream))
        _lib.check(rc)
        ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn)
        ctx.padding_mask = padding_mask
        ctx.logits_shape = attention_logits.shape
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn = ctx.saved_tensors
        N, S, M, D = value.shape
        L = spatial_shapes.shape[0]
        Lq, P = sampling_offsets.shape[1], sampling_offsets.shape[4]
        grad_output = grad_output.contiguous()
        hs, hl = ctx.geometry[:2]
        grad_value = torch.empty_like(value)
        grad_off = torch.empty_like(sampling_offsets)
        grad_logits = torch.empty(ctx.logits_shape, dtype=value.dtype, device=value.device)
        dev = value.device.index if value.device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = _lib.load().msda_fused_backward(_lib.F32, _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(reference_points),
                                             _ptr(sampling_offsets), _ptr(attn), _ptr(ctx.padding_mask), _ptr(grad_output), _ptr(grad_value), _ptr(grad_off),
                                             _ptr(grad_logits), N, S, M, D, L, Lq, P, int(ctx.im2col_step), hs, hl, dev,
                                             ctypes.c_void_p(stream))
        _lib.check(rc)
        return grad_value, None, None, None, grad_off, grad_logits, None, None
