#!/usr/bin/env python
"""GPU: torch.profiler over two training steps of the COMBO-R50 pixel decoder (40 frames, fp32): which ATen ops -- and which Python lines --
the time outside the library's kernels belongs to (copies, LayerNorm backward, bias-gradient reductions)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200  # noqa
from combo_avs_b200.pixel_decoder import MSDeformAttnPixelDecoder
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
r50 = {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)}
torch.manual_seed(7)
top = MSDeformAttnPixelDecoder(r50, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=6,
                               conv_dim=256, mask_dim=256, norm="GN", transformer_in_features=["res3", "res4", "res5"], common_stride=4).to(dev)
opt = torch.optim.AdamW(top.parameters(), lr=1e-4, fused=True)
g = torch.Generator().manual_seed(1)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
feats = {k: torch.randn(frames, c, 224 // s, 224 // s, generator=g).to(dev) for k, (c, s) in r50.items()}


def step():
    opt.zero_grad(set_to_none=True)
    mask_feat, low, multi = top(feats)
    (mask_feat.mean() + low.mean() + sum(m.mean() for m in multi)).backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=50, max_shapes_column_width=60))
print(prof.key_averages(group_by_stack_n=6).table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=40, max_src_column_width=110))
