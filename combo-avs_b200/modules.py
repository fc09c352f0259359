"""MSDeformAttn nn.Module -- drop-in for the reference module
(/root/reference/models/modeling/pixel_decoder/ops/modules/ms_deform_attn.py:34-129).

Same constructor, attributes, sub-module names (state_dict compatible: sampling_offsets / attention_weights /
value_proj / output_proj), initialisation scheme and forward signature.  One deliberate difference: the reference
wraps the op call in a bare `except:` that silently reroutes ANY failure to the slow pure-PyTorch path
(ms_deform_attn.py:119-125); this module lets errors propagate -- there is no CPU fallback.
"""
import math
import os
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from .linear import linear as tc_linear
from .functions import MSDeformAttnFunction, MSDeformAttnPrologueFunction, _host_geometry, fused_prologue_supported


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    # Fold softmax and the sampling-location arithmetic into the kernels when the call qualifies (CUDA float32, 2-d reference
    # points without gradient, S4-like shapes); same numbers, less memory traffic.  False reproduces the reference's composition.
    fuse_prologue = True
    # Run the four projections on the tcgen05 tensor cores at float32 accuracy (3xTF32, combo_avs_b200.linear) instead of as
    # strict-fp32 cuBLAS GEMMs when the call qualifies (CUDA float32, K % 4 == 0, N % 16 == 0).  False: plain nn.Linear.
    tensor_core_linear = os.environ.get("MSDA_TC_LINEAR", "1") != "0"

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        """Multi-scale deformable attention (ms_deform_attn.py:35-66).
        d_model: hidden size; n_levels: feature levels; n_heads: attention heads; n_points: samples per head per level."""
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head a power of 2 "
                          "which is more efficient in our CUDA implementation.")
        self.im2col_step = 128
        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # ms_deform_attn.py:68-84: zero offset weights; bias = (k+1) * unit direction of head m (max-norm), same for
        # every level; uniform attention (zero weight and bias); Xavier value / output projections with zero bias.
        constant_(self.sampling_offsets.weight.data, 0.0)
        theta = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        direction = torch.stack([theta.cos(), theta.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]
        steps = torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        grid = direction.view(self.n_heads, 1, 1, 2) * steps  # [M, 1, P, 2]
        grid = grid.expand(self.n_heads, self.n_levels, self.n_points, 2).contiguous()
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid.view(-1))
        constant_(self.attention_weights.weight.data, 0.0)
        constant_(self.attention_weights.bias.data, 0.0)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.0)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.0)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index, input_padding_mask=None):
        """query [N, Lq, C]; reference_points [N, Lq, L, 2] in [0,1] (or [N, Lq, L, 4] boxes); input_flatten [N, sum H_l*W_l, C];
        input_spatial_shapes [L, 2]; input_level_start_index [L]; input_padding_mask [N, sum H_l*W_l] (True = padding).
        Returns [N, Lq, C].  (ms_deform_attn.py:86-129)"""
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        # the reference's check (ms_deform_attn.py:100) reads a device scalar on every call -- a host synchronisation per layer, and
        # illegal under CUDA-graph capture; on CUDA the same comparison runs on the cached host mirror of the geometry tensors
        if input_spatial_shapes.is_cuda:
            assert sum(h * w for h, w in _host_geometry(input_spatial_shapes, input_level_start_index)[2]) == Len_in
        else:
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in

        proj = (lambda layer, x: tc_linear(x, layer.weight, layer.bias)) if self.tensor_core_linear else (lambda layer, x: layer(x))
        value = proj(self.value_proj, input_flatten).view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        sampling_offsets = proj(self.sampling_offsets, query).view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        attention_weights = proj(self.attention_weights, query).view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        if self.fuse_prologue and fused_prologue_supported(value, input_spatial_shapes, input_level_start_index, reference_points, sampling_offsets) \
                and (input_padding_mask is None or input_padding_mask.dtype == torch.bool):
            # softmax, location arithmetic and the padding mask happen inside the kernels (include/msda_b200.h: msda_fused_*)
            output = MSDeformAttnPrologueFunction.apply(value, input_spatial_shapes, input_level_start_index, reference_points.contiguous(),
                                                        sampling_offsets, attention_weights, self.im2col_step, input_padding_mask)
            return proj(self.output_proj, output)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None, None], float(0))
        attention_weights = F.softmax(attention_weights, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            sampling_locations = (reference_points[:, :, None, :, None, :2]
                                  + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5)
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        # no try/except here: a failing kernel must be loud (the reference's bare except hides it, ms_deform_attn.py:123)
        output = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index, sampling_locations.contiguous(),
                                            attention_weights.contiguous(), self.im2col_step)
        return proj(self.output_proj, output)
