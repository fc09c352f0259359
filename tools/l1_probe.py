#!/usr/bin/env python
"""GPU experiment: time the sorted32 backward on a pyramid whose slabs are small (three 14 x 14 levels, 588 queries), so that the
kernel's own shared memory is ~83 KB and MSDA_SORTED_SMEM_EXTRA (set by the caller) decides the L1 size: how does the walk depend on it?
(The switch -- extra dynamic shared memory added to the launch -- existed only in the experimental build of round 2, kept as
profiles/experiments/r07_*.patch; with the shipped library this script times the small pyramid as is.  Results: DESIGN.md 5.7, r07p.)
usage: python tools/l1_probe.py [frames]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200 as pkg

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
shapes_l = [(14, 14), (14, 14), (14, 14)]
M, D, P, L = 8, 32, 4, 3
S = sum(h * w for h, w in shapes_l)
LQ = S
shapes = torch.tensor(shapes_l, dtype=torch.long, device=dev)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
sets = []
for i in range(4):
    g = torch.Generator().manual_seed(5 + i)
    v = (torch.rand(frames, S, M, D, generator=g) * 0.01).to(dev)
    l = torch.rand(frames, LQ, M, L, P, 2, generator=g).to(dev)
    a = torch.rand(frames, LQ, M, L, P, generator=g) + 1e-5
    a = (a / a.sum(-1, keepdim=True).sum(-2, keepdim=True)).to(dev)
    go = torch.rand(frames, LQ, M * D, generator=g).to(dev)
    sets.append((v, l, a, go))
pkg.select_variant("backward", "sorted32")
for i in range(10):
    pkg.ms_deform_attn_backward(sets[i % 4][0], shapes, start, sets[i % 4][1], sets[i % 4][2], sets[i % 4][3], 128)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100):
    pkg.ms_deform_attn_backward(sets[i % 4][0], shapes, start, sets[i % 4][1], sets[i % 4][2], sets[i % 4][3], 128)
e1.record()
torch.cuda.synchronize()
print("3x(14x14) N=%d bwd[%s] %.1f us  env=%s" % (frames, pkg.last_variant("backward"), e0.elapsed_time(e1) * 10,
      {k: v for k, v in os.environ.items() if k.startswith("MSDA_")}), flush=True)
