#!/usr/bin/env python
"""GPU: two rounds of {forward, split32 backward, sorted32 backward} on the S4 shape and nothing else -- the target for
`ncu -k regex:'slab32|scatter32|sorted32'` (the second round is the warm one)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200 as pkg
from bench import S4_SHAPES, synth_inputs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
sets = [synth_inputs(torch, N, 3 + i, torch.float32, torch.float32, dev) for i in range(2)]
for v, l, a, g in sets:
    pkg.select_variant("backward", "split32")
    pkg.ms_deform_attn_forward(v, shapes, start, l, a, 128)
    pkg.ms_deform_attn_backward(v, shapes, start, l, a, g, 128)
    pkg.select_variant("backward", "sorted32")
    pkg.ms_deform_attn_backward(v, shapes, start, l, a, g, 128)
torch.cuda.synchronize()
print(pkg.last_variant("forward"), pkg.last_variant("backward"))
