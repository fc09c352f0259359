#!/usr/bin/env python
"""GPU: two training steps of the COMBO-R50 pixel decoder (40 frames, fp32) and nothing else -- the target for an ncu launch list
(`ncu --metrics gpu__time_duration.sum`), to see what the step is made of."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200  # noqa
from combo_avs_b200.pixel_decoder import MSDeformAttnPixelDecoder
dev = torch.device("cuda:0")
r50 = {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)}
torch.manual_seed(7)
top = MSDeformAttnPixelDecoder(r50, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=6,
                               conv_dim=256, mask_dim=256, norm="GN", transformer_in_features=["res3", "res4", "res5"], common_stride=4).to(dev)
with torch.no_grad():
    for layer in top.transformer.encoder.layers:
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
        layer.self_attn.attention_weights.weight.normal_(0, 0.1)
opt = torch.optim.AdamW(top.parameters(), lr=1e-4, fused=True)
g = torch.Generator().manual_seed(1)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
feats = {k: torch.randn(frames, c, 224 // s, 224 // s, generator=g).to(dev) for k, (c, s) in r50.items()}
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    opt.zero_grad(set_to_none=True)
    mask_feat, low, multi = top(feats)
    (mask_feat.mean() + low.mean() + sum(m.mean() for m in multi)).backward()
    opt.step()
torch.cuda.synchronize()
