#!/usr/bin/env python
"""GPU: a few forward + backward calls of MSDeformAttnPrologueFunction (S4 shape, N frames, all-False padding mask) and nothing
else -- the target for `ncu -k regex:'slab32|sorted32|softmax_backward|zero_masked'`."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200 as pkg
from bench import S4_SHAPES, M, D, L, P, S, LQ

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
g = torch.Generator().manual_seed(5)
value = (torch.rand(N, S, M, D, generator=g) * 0.01).to(dev).requires_grad_(True)
off = (torch.randn(N, LQ, M, L, P, 2, generator=g) * 2).to(dev).requires_grad_(True)
logits = torch.randn(N, LQ, M, L * P, generator=g).to(dev).requires_grad_(True)
ref = torch.rand(N, LQ, L, 2, generator=g).to(dev)
gout = torch.rand(N, LQ, M * D, generator=g).to(dev)
mask = torch.zeros(N, S, dtype=torch.bool, device=dev)
for _ in range(4):
    value.grad = off.grad = logits.grad = None
    pkg.MSDeformAttnPrologueFunction.apply(value, shapes, start, ref, off, logits, 128, mask).backward(gout)
torch.cuda.synchronize()
print(pkg.last_variant("forward"), pkg.last_variant("backward"))
