"""The module that owns the hot path: COMBO's multi-scale deformable-attention pixel decoder, without detectron2.

Equivalent of MSDeformAttnPixelDecoder (/root/reference/models/modeling/pixel_decoder/msdeformattn.py:169-368): 1x1 input
projections + GroupNorm for the encoder's feature levels, sine position encoding, the deformable encoder (encoder.py, i.e. the
sm_100a kernels), then an FPN-style top-down pass over the higher-resolution backbone features and a 1x1 mask-feature head.
Same sub-module / state-dict names (input_proj.N.0/1, transformer.*, adapter_K, layer_K, mask_features), same initialisation
schemes, same forward_features contract; the detectron2 pieces it needs are written out here:
    Conv2d with an attached norm and activation   -> ConvNormAct (detectron2.layers.Conv2d semantics: conv, then norm, then act)
    get_norm("GN", C)                             -> GroupNorm(32, C); "" -> no norm, conv bias on
    fvcore c2_xavier_fill                         -> kaiming_uniform_(a=1) on the weight, zero bias
    ShapeSpec                                     -> (channels, stride) pairs
    @configurable / from_config                   -> plain keyword arguments
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .encoder import MSDeformAttnTransformerEncoderOnly


class ConvNormAct(nn.Conv2d):
    """nn.Conv2d followed by an optional norm module (attribute `norm`, so its parameters are `<name>.norm.*`) and activation."""

    def __init__(self, *args, norm=None, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = super().forward(x)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def _caffe2_xavier(conv):
    nn.init.kaiming_uniform_(conv.weight, a=1)
    if conv.bias is not None:
        nn.init.constant_(conv.bias, 0)


def _make_norm(norm, channels):
    if norm in (None, ""):
        return None
    if norm == "GN":
        return nn.GroupNorm(32, channels)
    if callable(norm):
        return norm(channels)
    raise ValueError("unsupported norm %r (GN, '' or a callable)" % (norm,))


class PositionEmbeddingSine(nn.Module):
    """Sine / cosine position encoding of an unpadded feature map (transformer_decoder/position_encoding.py:12-52 with mask=None):
    cumulative row / column index, optionally normalised to (0, scale], divided by temperature^(2*(k//2)/F), sin on even and
    cos on odd feature indices, y features first.  Returns [N, 2F, H, W]."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and not normalize:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats, self.temperature, self.normalize = num_pos_feats, temperature, normalize
        self.scale = 2 * math.pi if scale is None else scale

    def forward(self, x, mask=None):
        n, _, h, w = x.shape
        if mask is None:
            ys = torch.arange(1, h + 1, dtype=torch.float32, device=x.device).view(1, h, 1).expand(n, h, w)
            xs = torch.arange(1, w + 1, dtype=torch.float32, device=x.device).view(1, 1, w).expand(n, h, w)
        else:
            keep = ~mask
            ys, xs = keep.cumsum(1, dtype=torch.float32), keep.cumsum(2, dtype=torch.float32)
        if self.normalize:
            ys = ys / (ys[:, -1:, :] + 1e-6) * self.scale
            xs = xs / (xs[:, :, -1:] + 1e-6) * self.scale
        k = torch.arange(self.num_pos_feats, dtype=torch.float32, device=x.device)
        freq = self.temperature ** (2 * torch.div(k, 2, rounding_mode="floor") / self.num_pos_feats)

        def encode(coord):
            angle = coord[..., None] / freq
            return torch.stack((angle[..., 0::2].sin(), angle[..., 1::2].cos()), dim=4).flatten(3)

        return torch.cat((encode(ys), encode(xs)), dim=3).permute(0, 3, 1, 2)


class MSDeformAttnPixelDecoder(nn.Module):
    """input_shape: {name: (channels, stride)} for every backbone feature the decoder sees (COMBO: res2..res5);
    transformer_in_features: the ones fed to the deformable encoder (COMBO: res3, res4, res5).  See the module docstring."""

    def __init__(self, input_shape, *, transformer_dropout, transformer_nheads, transformer_dim_feedforward, transformer_enc_layers,
                 conv_dim, mask_dim, norm=None, transformer_in_features, common_stride):
        super().__init__()
        by_stride = sorted(((k, tuple(v)) for k, v in input_shape.items()), key=lambda kv: kv[1][1])
        self.in_features = [k for k, _ in by_stride]
        self.feature_channels = [v[0] for _, v in by_stride]
        self.feature_strides = [v[1] for _, v in by_stride]
        enc = [(k, v) for k, v in by_stride if k in transformer_in_features]
        self.transformer_in_features = [k for k, _ in enc]
        self.transformer_feature_strides = [v[1] for _, v in enc]
        self.transformer_num_feature_levels = len(enc)
        enc_channels = [v[0] for _, v in enc]
        proj_channels = enc_channels[::-1] if len(enc) > 1 else enc_channels[-1:]   # low resolution first (res5 -> res3)
        self.input_proj = nn.ModuleList(nn.Sequential(nn.Conv2d(c, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim)) for c in proj_channels)
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)
        self.transformer = MSDeformAttnTransformerEncoderOnly(
            d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads, dim_feedforward=transformer_dim_feedforward,
            num_encoder_layers=transformer_enc_layers, num_feature_levels=self.transformer_num_feature_levels)
        self.pe_layer = PositionEmbeddingSine(conv_dim // 2, normalize=True)
        self.mask_dim = mask_dim
        self.mask_features = ConvNormAct(conv_dim, mask_dim, kernel_size=1, stride=1, padding=0)
        _caffe2_xavier(self.mask_features)
        self.maskformer_num_feature_levels = 3
        self.common_stride = common_stride
        self.num_fpn_levels = int(math.log2(min(self.transformer_feature_strides)) - math.log2(common_stride))
        laterals, outputs = [], []
        for idx, channels in enumerate(self.feature_channels[: self.num_fpn_levels]):
            # msdeformattn.py:271: use_bias = norm == ""  (norm=None builds the convolutions without bias AND without norm)
            lateral = ConvNormAct(channels, conv_dim, kernel_size=1, bias=norm == "", norm=_make_norm(norm, conv_dim))
            output = ConvNormAct(conv_dim, conv_dim, kernel_size=3, stride=1, padding=1, bias=norm == "",
                                 norm=_make_norm(norm, conv_dim), activation=F.relu)
            _caffe2_xavier(lateral)
            _caffe2_xavier(output)
            self.add_module("adapter_%d" % (idx + 1), lateral)
            self.add_module("layer_%d" % (idx + 1), output)
            laterals.append(lateral)
            outputs.append(output)
        self.lateral_convs, self.output_convs = laterals[::-1], outputs[::-1]   # top-down order

    def forward_features(self, features):
        """features: {name: [N, C, H, W]}.  Returns (mask_features [N, mask_dim, H/4, W/4], the lowest-resolution encoder map,
        the first three maps of the top-down pyramid), as msdeformattn.py:315-368 -- always in float32."""
        with torch.autocast("cuda", enabled=False):
            maps = [features[f].float() for f in self.transformer_in_features[::-1]]
            srcs = [proj(x) for proj, x in zip(self.input_proj, maps)]
            memory, spatial_shapes, level_start_index = self.transformer(srcs, [self.pe_layer(x) for x in maps])
            batch = memory.shape[0]
            sizes = [int(x.shape[2]) * int(x.shape[3]) for x in srcs]
            out = [z.transpose(1, 2).reshape(batch, -1, x.shape[2], x.shape[3]) for z, x in zip(torch.split(memory, sizes, dim=1), srcs)]
            for lateral, output, name in zip(self.lateral_convs, self.output_convs, self.in_features[: self.num_fpn_levels][::-1]):
                skip = lateral(features[name].float())
                out.append(output(skip + F.interpolate(out[-1], size=skip.shape[-2:], mode="bilinear", align_corners=False)))
            return self.mask_features(out[-1]), out[0], out[: self.maskformer_num_feature_levels]

    forward = forward_features
