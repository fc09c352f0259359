"""PVTv2-B5 feature pyramid for the harness (BASELINE configs 4 / 5: COMBO-PVTv2-B5 on MS3 / AVSS) -- reference
models/modeling/backbone/pvtv2.py:236-409 (`build_pvtv2_b5_backbone`: embed dims 64/128/320/512, depths 3/6/40/3, heads 1/2/5/8,
spatial-reduction ratios 8/4/2/1, MLP ratio 4, qkv bias, LayerNorm eps 1e-6), restated from the architecture:

  stage i = overlapping patch embedding (7x7 stride 4 for the first stage, 3x3 stride 2 after) + LayerNorm, then `depth` pre-norm blocks
            x += Attn(LN(x));  x += Mlp(LN(x))            (stochastic depth is a training regulariser, not modelled: rate 0)
  Attn     = multi-head attention whose keys / values come from the token map reduced by a strided convolution (sr x sr, stride sr) + LN
  Mlp      = Linear -> 3x3 depth-wise convolution on the token map -> GELU -> Linear
  output   = LayerNorm of the stage's tokens, reshaped to [B, C_i, H/2^(i+2), W/2^(i+2)]   ("res2" .. "res5", strides 4 .. 32)

Harness code: random initialisation (no checkpoint ships), dense library kernels (cuDNN / cuBLAS / SDPA), no parity claim."""
import torch
import torch.nn.functional as F
from torch import nn

PVTV2_B5_FEATURES = {"res2": (64, 4), "res3": (128, 8), "res4": (320, 16), "res5": (512, 32)}


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.dwconv = nn.Conv2d(hidden, hidden, 3, 1, 1, groups=hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x, hw):
        b, n, _ = x.shape
        x = self.fc1(x)
        x = self.dwconv(x.transpose(1, 2).reshape(b, -1, *hw)).flatten(2).transpose(1, 2)
        return self.fc2(F.gelu(x))


class _SRAttention(nn.Module):
    def __init__(self, dim, heads, sr):
        super().__init__()
        self.heads = heads
        self.q = nn.Linear(dim, dim)
        self.kv = nn.Linear(dim, 2 * dim)
        self.proj = nn.Linear(dim, dim)
        self.sr = nn.Conv2d(dim, dim, sr, sr) if sr > 1 else None
        self.norm = nn.LayerNorm(dim, eps=1e-6) if sr > 1 else None

    def forward(self, x, hw):
        b, n, c = x.shape
        q = self.q(x).reshape(b, n, self.heads, c // self.heads).transpose(1, 2)
        src = x
        if self.sr is not None:
            src = self.norm(self.sr(x.transpose(1, 2).reshape(b, c, *hw)).flatten(2).transpose(1, 2))
        k, v = self.kv(src).reshape(b, -1, 2, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        out = F.scaled_dot_product_attention(q, k, v)
        return self.proj(out.transpose(1, 2).reshape(b, n, c))


class _Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio, sr):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _SRAttention(dim, heads, sr)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, dim * mlp_ratio)

    def forward(self, x, hw):
        x = x + self.attn(self.norm1(x), hw)
        return x + self.mlp(self.norm2(x), hw)


class PVTv2B5Features(nn.Module):
    """images [B, 3, H, W] -> {"res2": [B, 64, H/4, W/4], "res3": [B, 128, H/8, W/8], "res4": [B, 320, H/16, W/16], "res5": [B, 512, H/32, W/32]}"""

    def __init__(self, embed_dims=(64, 128, 320, 512), depths=(3, 6, 40, 3), heads=(1, 2, 5, 8), sr_ratios=(8, 4, 2, 1), mlp_ratio=4):
        super().__init__()
        self.patch_embed, self.embed_norm, self.blocks, self.norm = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for i, dim in enumerate(embed_dims):
            k, s = (7, 4) if i == 0 else (3, 2)
            self.patch_embed.append(nn.Conv2d(3 if i == 0 else embed_dims[i - 1], dim, k, s, k // 2))
            self.embed_norm.append(nn.LayerNorm(dim, eps=1e-6))
            self.blocks.append(nn.ModuleList(_Block(dim, heads[i], mlp_ratio, sr_ratios[i]) for _ in range(depths[i])))
            self.norm.append(nn.LayerNorm(dim, eps=1e-6))
        self.apply(self._init)

    @staticmethod
    def _init(m):  # truncated-normal Linear weights, fan-out normal convolutions, unit LayerNorm (the reference's `_init_weights`)
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv2d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
            nn.init.normal_(m.weight, 0.0, (2.0 / fan_out) ** 0.5)
            nn.init.zeros_(m.bias)

    def forward(self, x):
        out = {}
        for i, name in enumerate(PVTV2_B5_FEATURES):
            x = self.patch_embed[i](x)
            hw = x.shape[-2:]
            x = self.embed_norm[i](x.flatten(2).transpose(1, 2))
            for blk in self.blocks[i]:
                x = blk(x, hw)
            x = self.norm[i](x).transpose(1, 2).reshape(x.shape[0], -1, *hw).contiguous()
            out[name] = x
        return out
