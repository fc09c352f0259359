"""Restatement of the reference's pure-PyTorch CPU path -- TEST INFRASTRUCTURE ONLY.

The reference's only CPU implementation of the hot path is ms_deform_attn_core_pytorch
(ops/functions/ms_deform_attn_func.py:53-72): per level, F.grid_sample(bilinear, zeros,
align_corners=False) of the level's feature map at 2*loc-1, then the attention-weighted sum over
(level, point).  /root/reference does not exist on the GPU box, so this module restates that
algorithm (same torch ops, so the same arithmetic and the same threading behaviour) for
 * bench.py's cpu_baseline / --impl reference legs (the timed CPU baseline), and
 * tests that pin it against tests/golden/*.npz.
It is never imported by the product package.
"""
import torch
import torch.nn.functional as F


def msda_core_pytorch(value, spatial_shapes, sampling_locations, attention_weights):
    """value[N,S,M,D], spatial_shapes[L,2] (H,W), sampling_locations[N,Lq,M,L,P,2] (x,y in [0,1]),
    attention_weights[N,Lq,M,L,P] -> [N, Lq, M*D].  Differentiable in all three float inputs."""
    N, S, M, D = value.shape
    Lq, L, P = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    hw = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    per_level = torch.split(value, [h * w for h, w in hw], dim=1)
    grid = sampling_locations * 2.0 - 1.0  # grid_sample's [-1, 1] convention
    sampled = []
    for lvl, (h, w) in enumerate(hw):
        # feature map of this level as an image batch over (frame, head): [N*M, D, h, w]
        fmap = per_level[lvl].permute(0, 2, 3, 1).reshape(N * M, D, h, w)
        # sample positions of this level: [N*M, Lq, P, 2]
        pos = grid[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        sampled.append(F.grid_sample(fmap, pos, mode="bilinear", padding_mode="zeros", align_corners=False))
    stacked = torch.stack(sampled, dim=3).reshape(N * M, D, Lq, L * P)  # [N*M, D, Lq, L*P]
    wts = attention_weights.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L * P)
    out = (stacked * wts).sum(dim=-1)  # [N*M, D, Lq]
    return out.reshape(N, M * D, Lq).permute(0, 2, 1).contiguous()
