"""Host-buffer entry point: the MSDeformAttn core forward+backward on tensors that live in (pinned) HOST memory.

This is the end-to-end form of the reference-facing call (what `bench.py`'s `e2e` leg times): the inputs of one
step are copied host->device, the sm_100a kernels run, and every result is copied back device->host.  Frames are
independent in this op (ops/src/cuda/ms_deform_im2col_cuda.cuh:268-274: the batch index only selects the value
slab), so the batch is cut into frame chunks that are pipelined over three CUDA streams.  The pipeline itself is
native (`msda_forward_backward_host` in libmsda_b200.so, include/msda_b200.h): Python only hands over the host
pointers, so small chunks (short fill / drain) do not pay a per-chunk interpreter cost.
"""
import ctypes

import torch

from . import _lib
from .functions import _DTYPES, _host_geometry


class HostPipeline:
    """Owns one native pipeline object: three streams + `slots` rotating device staging buffers on `device`."""

    def __init__(self, device, chunk_frames=4, slots=3):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Not implemented on the CPU")  # ms_deform_attn.h:43
        self.chunk = int(chunk_frames)
        self.slots = int(slots)
        self._lib = _lib.load()
        self._handle = ctypes.c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = self._lib.msda_host_pipeline_create(ctypes.byref(self._handle), index, self.chunk, self.slots)
        if rc != 0:
            raise RuntimeError("libmsda_b200: msda_host_pipeline_create failed (status %d)" % rc)

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle:
            self._lib.msda_host_pipeline_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward_backward(self, value, shapes_dev, start_dev, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn,
                         im2col_step=128):
        """All float tensors are HOST tensors (pin them for asynchronous copies); shapes_dev / start_dev are the int64
        geometry tensors (device or host).  Results are written into the host tensors out / grad_value / grad_loc /
        grad_attn.  Returns (h2d_bytes, d2h_bytes).  The call returns after everything has landed on the host."""
        named = (("value", value), ("sampling_loc", loc), ("attn_weight", attn), ("grad_output", grad_out), ("output", out),
                 ("grad_value", grad_value), ("grad_sampling_loc", grad_loc), ("grad_attn_weight", grad_attn))
        for name, t in named:
            if t.is_cuda:
                raise RuntimeError("%s must be a host tensor for the host-buffer entry point" % name)
            if not t.is_contiguous():
                raise RuntimeError("%s tensor has to be contiguous" % name)
        if value.dtype not in _DTYPES or loc.dtype not in _DTYPES:
            raise RuntimeError('"ms_deform_attn_cuda" not implemented for \'%s\'' % str(value.dtype).replace("torch.", ""))
        N, S, M, D = value.shape
        L = shapes_dev.shape[0]
        Lq, P = loc.shape[1], loc.shape[4]
        if tuple(loc.shape) != (N, Lq, M, L, P, 2) or tuple(attn.shape) != (N, Lq, M, L, P) or grad_out.numel() != N * Lq * M * D:
            raise RuntimeError("inconsistent shapes: value %s, sampling_loc %s, attn_weight %s, grad_output %s" %
                               (tuple(value.shape), tuple(loc.shape), tuple(attn.shape), tuple(grad_out.shape)))
        for t, like in ((out, grad_out), (grad_value, value), (grad_loc, loc), (grad_attn, attn)):
            if t.dtype != like.dtype or t.numel() != like.numel():
                raise RuntimeError("result buffers must match their counterparts in dtype and size")
        hs, hl = _host_geometry(shapes_dev, start_dev)[:2]
        up, down = ctypes.c_uint64(0), ctypes.c_uint64(0)
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t.numel() else None
        rc = self._lib.msda_forward_backward_host(self._handle, _DTYPES[value.dtype], _DTYPES[loc.dtype], p(value), hs, hl, p(loc), p(attn),
                                                  p(grad_out), p(out), p(grad_value), p(grad_loc), p(grad_attn), N, S, M, D, L, Lq, P,
                                                  int(im2col_step), ctypes.byref(up), ctypes.byref(down))
        if rc != 0:
            msg = self._lib.msda_host_pipeline_error(self._handle).decode("utf-8", "replace")
            if rc in (-2, -3, -5):
                raise ValueError("libmsda_b200: %s (status %d)" % (msg, rc))
            raise RuntimeError("libmsda_b200: %s (status %d)" % (msg, rc))
        return int(up.value), int(down.value)


def msda_forward_backward_host(value, shapes_dev, start_dev, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn,
                               im2col_step=128, pipeline=None, chunk_frames=4):
    """One-shot convenience wrapper around HostPipeline (creates and destroys a pipeline when none is passed)."""
    pipe = pipeline or HostPipeline(shapes_dev.device if shapes_dev.is_cuda else "cuda", chunk_frames)
    try:
        return pipe.forward_backward(value, shapes_dev, start_dev, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn, im2col_step)
    finally:
        if pipeline is None:
            pipe.close()
