#!/usr/bin/env python
"""GPU: what the fused prologue buys at module level, S4 shape, N frames, fp32, CUDA events, median of 50.
  core      -- MSDeformAttnFunction on precomputed locations / weights (what bench.py times)
  unfused   -- softmax + location arithmetic in PyTorch, then the core (the reference's composition, ms_deform_attn.py:102-120)
  fused     -- MSDeformAttnPrologueFunction on the Linear outputs
  module_*  -- the whole MSDeformAttn module (four Linear layers included), with the all-False padding mask the pixel decoder
               passes (msdeformattn.py:68,86) and without a mask."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200 as pkg
from bench import S4_SHAPES, M, D, L, P, S, LQ

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
g = torch.Generator(device="cpu").manual_seed(5)
value = (torch.rand(N, S, M, D, generator=g) * 0.01).to(dev).requires_grad_(True)
off = (torch.randn(N, LQ, M, L, P, 2, generator=g) * 2).to(dev).requires_grad_(True)
logits = torch.randn(N, LQ, M, L * P, generator=g).to(dev).requires_grad_(True)
ref = torch.rand(N, LQ, L, 2, generator=g).to(dev)
gout = torch.rand(N, LQ, M * D, generator=g).to(dev)
norm = torch.tensor([[w, h] for h, w in S4_SHAPES], dtype=torch.float32, device=dev).view(1, 1, 1, L, 1, 2)


def timed(fn, iters=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    us = [a.elapsed_time(b) * 1e3 for a, b in evs]
    return statistics.median(us)


def zero():
    value.grad = off.grad = logits.grad = None


loc_c = (ref[:, :, None, :, None, :] + off.detach() / norm).contiguous().requires_grad_(True)
attn_c = torch.softmax(logits.detach(), -1).view(N, LQ, M, L, P).contiguous().requires_grad_(True)


def core():
    value.grad = loc_c.grad = attn_c.grad = None
    pkg.MSDeformAttnFunction.apply(value, shapes, start, loc_c, attn_c, 128).backward(gout)


def unfused():
    zero()
    a = torch.softmax(logits, -1).view(N, LQ, M, L, P)
    loc = ref[:, :, None, :, None, :] + off / norm
    pkg.MSDeformAttnFunction.apply(value, shapes, start, loc.contiguous(), a.contiguous(), 128).backward(gout)


def fused():
    zero()
    pkg.MSDeformAttnPrologueFunction.apply(value, shapes, start, ref, off, logits, 128).backward(gout)


res = {"frames": N, "core_us": timed(core), "unfused_prologue_plus_core_us": timed(unfused), "fused_us": timed(fused)}
mod = pkg.MSDeformAttn(d_model=256, n_levels=3, n_heads=8, n_points=4).to(dev)
with torch.no_grad():
    mod.sampling_offsets.weight.normal_(0, 0.02)
    mod.attention_weights.weight.normal_(0, 0.1)
query = torch.randn(N, LQ, 256, generator=g).to(dev).requires_grad_(True)
src = torch.randn(N, S, 256, generator=g).to(dev).requires_grad_(True)
mask = torch.zeros(N, S, dtype=torch.bool, device=dev)
gy = torch.rand(N, LQ, 256, generator=g).to(dev)
for fuse in (False, True):
    for mk, mname in ((mask, "allfalse_mask"), (None, "no_mask")):
        mod.fuse_prologue = fuse

        def step():
            query.grad = src.grad = None
            mod.zero_grad(set_to_none=True)
            mod(query, ref, src, shapes, start, mk).backward(gy)
        res["module_%s_%s_us" % ("fused" if fuse else "unfused", mname)] = timed(step)
print(json.dumps(res, indent=1))
