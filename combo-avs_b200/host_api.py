"""Host-buffer entry point: the MSDeformAttn core forward+backward on tensors that live in (pinned) HOST memory.

This is the end-to-end form of the reference-facing call (what `bench.py`'s `e2e` leg times): the inputs of one
step are copied host->device, the sm_100a kernels run through the C ABI, and every result is copied back
device->host.  Frames are independent in this op (ops/src/cuda/ms_deform_im2col_cuda.cuh:268-274: the batch index
only selects the value slab), so the batch is cut into frame chunks and pipelined over three CUDA streams --
H2D of chunk i+1, compute of chunk i and D2H of chunk i-1 overlap, and PCIe runs full duplex.
"""
import torch

from .functions import ms_deform_attn_backward, ms_deform_attn_forward


class HostPipeline:
    """Reusable device staging buffers + streams for msda_forward_backward_host."""

    def __init__(self, device, chunk_frames=8):
        self.device = torch.device(device)
        self.chunk = int(chunk_frames)
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._bufs = {}

    def _stage(self, key, like, n):
        shape = (n,) + tuple(like.shape[1:])
        buf = self._bufs.get(key)
        if buf is None or buf.shape != shape or buf.dtype != like.dtype:
            buf = torch.empty(shape, dtype=like.dtype, device=self.device)
            self._bufs[key] = buf
        return buf

    def forward_backward(self, value, shapes_dev, start_dev, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn,
                         im2col_step=128):
        """All float tensors are HOST tensors (pinned for async copies); shapes_dev / start_dev are the int64 geometry
        tensors on the device.  Results are written into the host tensors out / grad_value / grad_loc / grad_attn.
        Returns (h2d_bytes, d2h_bytes).  The call returns after everything has landed on the host."""
        N = value.shape[0]
        h2d = d2h = 0
        prev = None
        nchunks = (N + self.chunk - 1) // self.chunk
        slots = 2  # double-buffered staging
        done_out = [None] * slots
        for i in range(nchunks + 1):
            cur = None
            if i < nchunks:
                lo, hi = i * self.chunk, min(N, (i + 1) * self.chunk)
                n, slot = hi - lo, i % slots
                with torch.cuda.stream(self.s_in):
                    if done_out[slot] is not None:  # staging buffers of this slot are free once its D2H finished
                        self.s_in.wait_event(done_out[slot])
                    dv = self._stage(("v", slot), value, n)
                    dl = self._stage(("l", slot), loc, n)
                    da = self._stage(("a", slot), attn, n)
                    dg = self._stage(("g", slot), grad_out, n)
                    dv.copy_(value[lo:hi], non_blocking=True)
                    dl.copy_(loc[lo:hi], non_blocking=True)
                    da.copy_(attn[lo:hi], non_blocking=True)
                    dg.copy_(grad_out[lo:hi], non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(self.s_in)
                h2d += sum(t[lo:hi].numel() * t.element_size() for t in (value, loc, attn, grad_out))
                with torch.cuda.stream(self.s_run):
                    self.s_run.wait_event(ready)
                    step = min(n, im2col_step)
                    while n % step:
                        step -= 1
                    o = ms_deform_attn_forward(dv, shapes_dev, start_dev, dl, da, step)
                    gv, gl, ga = ms_deform_attn_backward(dv, shapes_dev, start_dev, dl, da, dg, step)
                    computed = torch.cuda.Event()
                    computed.record(self.s_run)
                cur = (lo, hi, slot, o, gv, gl, ga, computed)
            if prev is not None:
                lo, hi, slot, o, gv, gl, ga, computed = prev
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(computed)
                    out[lo:hi].copy_(o, non_blocking=True)
                    grad_value[lo:hi].copy_(gv, non_blocking=True)
                    grad_loc[lo:hi].copy_(gl, non_blocking=True)
                    grad_attn[lo:hi].copy_(ga, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.s_out)
                    done_out[slot] = ev
                    for t in (o, gv, gl, ga):  # allocator: these blocks are in use on s_out until the copies finish
                        t.record_stream(self.s_out)
                d2h += sum(t.numel() * t.element_size() for t in (o, gv, gl, ga))
            prev = cur
        self.s_out.synchronize()
        return h2d, d2h


def msda_forward_backward_host(value, shapes_dev, start_dev, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn,
                               im2col_step=128, pipeline=None, chunk_frames=8):
    """One-shot convenience wrapper around HostPipeline (allocates staging buffers on every call when no pipeline
    is passed)."""
    pipe = pipeline or HostPipeline(shapes_dev.device, chunk_frames)
    return pipe.forward_backward(value, shapes_dev, start_dev, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn,
                                 im2col_step)
