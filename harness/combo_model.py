"""COMBO-R50 (S4 / MS3 settings) without detectron2: reference models/maskformer_model.py:274-391 (forward),
modeling/meta_arch/mask_former_head.py:141-159 (head with late AVFuse), modeling/fusion_module (AVFuse "MHA-B"),
modeling/transformer_decoder/transformer_decoder.py:402-519 (9-layer masked decoder), modeling/audio_backbone/torchvggish/vggish.py,
utils/misc.py:112-131 (channel gate).  The pixel decoder -- the caller of the hot path -- is the product's own."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from combo_avs_b200.pixel_decoder import MSDeformAttnPixelDecoder, PositionEmbeddingSine

R50_FEATURES = {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)}


class FrozenBatchNorm2d(nn.Module):
    """detectron2's default ResNet norm ("FrozenBN"): fixed statistics and affine, no gradients."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def forward(self, x):
        scale = self.weight * (self.running_var + 1e-5).rsqrt()
        shift = self.bias - self.running_mean * scale
        return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


class ResNet50Features(nn.Module):
    """ResNet-50 trunk returning res2..res5 (strides 4..32; build_resnet_backbone with STRIDE_IN_1X1 False, FrozenBN, FREEZE_AT 0:
    configs/avs_s4/R50-AVSS4-SemanticSegmentation.yaml).  torchvision's definition, random init."""

    def __init__(self):
        super().__init__()
        import torchvision
        net = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)
        self.stem = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool)
        self.res2, self.res3, self.res4, self.res5 = net.layer1, net.layer2, net.layer3, net.layer4

    def forward(self, x):
        out = {}
        x = self.stem(x)
        for name in ("res2", "res3", "res4", "res5"):
            x = getattr(self, name)(x)
            out[name] = x
        return out


class ChannelGate(nn.Module):
    """channel_weighted_block (utils/misc.py:112-131): squeeze-and-excitation weights for the pre-mask branch."""

    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, dim // 16)
        self.fc2 = nn.Linear(dim // 16, dim)

    def forward(self, x):
        y = x.mean((2, 3))
        return torch.sigmoid(self.fc2(F.relu(self.fc1(y)))).view(x.shape[0], -1, 1, 1)


class VGGish(nn.Module):
    """torchvggish VGG (vggish.py:9-27, 95-110): log-mel patch [B, 1, 96, 64] -> 128-d embedding.  Frozen in COMBO
    (FREEZE_AUDIO_EXTRACTOR True) and called under no_grad (maskformer_model.py:327-328)."""

    def __init__(self):
        super().__init__()
        layers, c = [], 1
        for v in (64, "M", 128, "M", 256, 256, "M", 512, 512, "M"):
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(c, v, 3, padding=1), nn.ReLU(inplace=True)]
                c = v
        self.features = nn.Sequential(*layers)
        self.embeddings = nn.Sequential(nn.Linear(512 * 4 * 6, 4096), nn.ReLU(True), nn.Linear(4096, 4096), nn.ReLU(True),
                                        nn.Linear(4096, 128), nn.ReLU(True))

    def forward(self, x):
        x = self.features(x).permute(0, 2, 3, 1).contiguous()
        return self.embeddings(x.view(x.shape[0], -1))


class BiMultiHeadAttention(nn.Module):
    """fuse_helper.py:102-237: visual tokens attend to the single audio token and vice versa (8 heads, embed 256)."""

    def __init__(self, v_dim, a_dim, embed_dim, num_heads, dropout=0.1):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.dropout = dropout
        self.v_proj, self.a_proj = nn.Linear(v_dim, embed_dim), nn.Linear(a_dim, embed_dim)
        self.values_v_proj, self.values_a_proj = nn.Linear(v_dim, embed_dim), nn.Linear(a_dim, embed_dim)
        self.out_v_proj, self.out_a_proj = nn.Linear(embed_dim, v_dim), nn.Linear(embed_dim, a_dim)
        for m in (self.v_proj, self.a_proj, self.values_v_proj, self.values_a_proj, self.out_v_proj, self.out_a_proj):
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)

    def _heads(self, t, bsz):
        return t.view(bsz, -1, self.num_heads, self.head_dim).transpose(1, 2).reshape(bsz * self.num_heads, -1, self.head_dim)

    def forward(self, v, a, pos_v=None, pos_a=None):
        bsz, tgt_len, _ = v.shape
        q = self._heads(self.v_proj(v if pos_v is None else v + pos_v) * self.scale, bsz)
        k = self._heads(self.a_proj(a if pos_a is None else a + pos_a), bsz)
        val_v, val_a = self._heads(self.values_v_proj(v), bsz), self._heads(self.values_a_proj(a), bsz)
        attn = torch.bmm(q, k.transpose(1, 2)).clamp(-50000, 50000)            # [B*h, HW, 1]
        attn_t = attn.transpose(1, 2)
        attn_a = (attn_t - attn_t.max(-1, keepdim=True)[0]).clamp(-50000, 50000).softmax(-1)
        attn_v = attn.softmax(1)                                                # (the reference normalises over the tokens: :204)
        out_v = torch.bmm(F.dropout(attn_v, self.dropout, self.training), val_a)
        out_a = torch.bmm(F.dropout(attn_a, self.dropout, self.training), val_v)
        out_v = out_v.view(bsz, self.num_heads, tgt_len, self.head_dim).transpose(1, 2).reshape(bsz, tgt_len, self.embed_dim)
        out_a = out_a.view(bsz, self.num_heads, -1, self.head_dim).transpose(1, 2).reshape(bsz, -1, self.embed_dim)
        return self.out_v_proj(out_v), self.out_a_proj(out_a)


class AVFuse(nn.Module):
    """AVFuse "MHA-B" on the mask features (AVFuse.py:10-125 + BiAttentionBlock fuse_helper.py:240-333): pre-LayerNorm, the
    bi-directional attention, layer-scale residuals (init 1e-4), audio position embedding and sine image positions."""

    def __init__(self, vision_dim=256, audio_dim=128, num_heads=8):
        super().__init__()
        self.audio_pos = nn.Embedding(1, audio_dim)
        self.pe_layer = PositionEmbeddingSine(vision_dim // 2, normalize=True)
        self.level_embed = nn.Embedding(1, vision_dim)
        self.norm_v, self.norm_a = nn.LayerNorm(vision_dim), nn.LayerNorm(audio_dim)
        self.attn = BiMultiHeadAttention(vision_dim, audio_dim, vision_dim, num_heads, dropout=0.1)
        self.gamma_v = nn.Parameter(1e-4 * torch.ones(vision_dim))
        self.gamma_a = nn.Parameter(1e-4 * torch.ones(audio_dim))

    def forward(self, feat, audio):
        bs, c, h, w = feat.shape
        pos_v = self.pe_layer(feat).flatten(2).permute(0, 2, 1)
        pos_a = self.audio_pos.weight.unsqueeze(0).expand(bs, -1, -1)
        v = (feat.flatten(2) + self.level_embed.weight[0][None, :, None]).permute(0, 2, 1)
        v, a = self.norm_v(v), self.norm_a(audio)
        dv, da = self.attn(v, a, pos_v, pos_a)
        v = v + self.gamma_v * dv
        a = a + self.gamma_a * da
        return v.transpose(1, 2).reshape(bs, c, h, w), a


class AudioMLP(nn.Module):
    """audio_mlp (modeling/misc/audio_transformation.py): 128 -> 4096 -> 4096 -> 256."""

    def __init__(self, in_dim=128, middle_dim=4096, out_dim=256):
        super().__init__()
        self.embeddings = nn.Sequential(nn.Linear(in_dim, middle_dim), nn.ReLU(True), nn.Linear(middle_dim, middle_dim), nn.ReLU(True),
                                        nn.Linear(middle_dim, out_dim))

    def forward(self, x):
        return self.embeddings(x)


class _AttnLayer(nn.Module):
    """SelfAttentionLayer / CrossAttentionLayer, post-norm form (transformer_decoder.py:20-140)."""

    def __init__(self, d_model, nhead):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, nhead, dropout=0.0)
        self.norm = nn.LayerNorm(d_model)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, tgt, memory=None, memory_mask=None, pos=None, query_pos=None):
        q = tgt + query_pos
        if memory is None:
            out = self.attn(q, q, tgt)[0]
        else:
            out = self.attn(q, memory + pos, memory, attn_mask=memory_mask)[0]
        return self.norm(tgt + out)


class _FFNLayer(nn.Module):
    def __init__(self, d_model, dim_feedforward):
        super().__init__()
        self.linear1, self.linear2 = nn.Linear(d_model, dim_feedforward), nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, tgt):
        return self.norm(tgt + self.linear2(F.relu(self.linear1(tgt))))


class MultiScaleMaskedTransformerDecoder(nn.Module):
    """transformer_decoder.py:244-519 with COMBO's settings: 100 queries ("add" audio fusion into the query features), 9 layers of
    masked cross-attention (round-robin over the 3 scales) -> self-attention -> FFN (2048), prediction heads after every layer
    (deep supervision), mask logits kept per layer for the frame-to-frame cosine loss."""

    def __init__(self, in_channels=256, num_classes=2, hidden_dim=256, num_queries=100, nheads=8, dim_feedforward=2048, dec_layers=9,
                 mask_dim=256, use_cosine_loss=True):
        super().__init__()
        self.pe_layer = PositionEmbeddingSine(hidden_dim // 2, normalize=True)
        self.num_heads, self.num_layers, self.num_queries = nheads, dec_layers, num_queries
        self.cross = nn.ModuleList(_AttnLayer(hidden_dim, nheads) for _ in range(dec_layers))
        self.self_attn = nn.ModuleList(_AttnLayer(hidden_dim, nheads) for _ in range(dec_layers))
        self.ffn = nn.ModuleList(_FFNLayer(hidden_dim, dim_feedforward) for _ in range(dec_layers))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.num_feature_levels = 3
        self.level_embed = nn.Embedding(3, hidden_dim)
        self.input_proj = nn.ModuleList(nn.Sequential() if in_channels == hidden_dim else nn.Conv2d(in_channels, hidden_dim, 1) for _ in range(3))
        self.class_embed = nn.Linear(hidden_dim, num_classes + 1)
        self.mask_embed = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.ReLU(True), nn.Linear(hidden_dim, hidden_dim), nn.ReLU(True),
                                        nn.Linear(hidden_dim, mask_dim))
        self.use_cosine_loss = use_cosine_loss

    def heads(self, output, mask_features, size):
        dec = self.decoder_norm(output).transpose(0, 1)
        outputs_class = self.class_embed(dec)
        outputs_mask = torch.einsum("bqc,bchw->bqhw", self.mask_embed(dec), mask_features)
        attn_mask = F.interpolate(outputs_mask, size=size, mode="bilinear", align_corners=False)
        attn_mask = (attn_mask.sigmoid().flatten(2).unsqueeze(1).repeat(1, self.num_heads, 1, 1).flatten(0, 1) < 0.5).detach()
        return outputs_class, outputs_mask, attn_mask

    def forward(self, x, audio_features, mask_features):
        bt = mask_features.shape[0]
        src, pos, sizes = [], [], []
        for i in range(3):
            sizes.append(x[i].shape[-2:])
            pos.append(self.pe_layer(x[i]).flatten(2).permute(2, 0, 1))
            src.append((self.input_proj[i](x[i]).flatten(2) + self.level_embed.weight[i][None, :, None]).permute(2, 0, 1))
        query_embed = self.query_embed.weight.unsqueeze(1).repeat(1, bt, 1)
        output = self.query_feat.weight.unsqueeze(1).repeat(1, bt, 1)
        output = output + audio_features.repeat(1, self.num_queries, 1).reshape(self.num_queries, -1, audio_features.shape[-1])  # "add"
        classes, masks, middles = [], [], []
        c, m, attn_mask = self.heads(output, mask_features, sizes[0])
        classes.append(c)
        masks.append(m)
        if self.use_cosine_loss:
            middles.append(m.reshape(bt, self.num_queries, -1))
        for i in range(self.num_layers):
            lvl = i % 3
            attn_mask = attn_mask & ~attn_mask.all(-1, keepdim=True)  # a query whose mask is empty attends everywhere (:453)
            output = self.cross[i](output, src[lvl], memory_mask=attn_mask, pos=pos[lvl], query_pos=query_embed)
            output = self.self_attn[i](output, query_pos=query_embed)
            output = self.ffn[i](output)
            c, m, attn_mask = self.heads(output, mask_features, sizes[(i + 1) % 3])
            classes.append(c)
            masks.append(m)
            if self.use_cosine_loss and i != self.num_layers - 1:
                middles.append(m.reshape(bt, self.num_queries, -1))
        return {"pred_logits": classes[-1], "pred_masks": masks[-1],
                "aux_outputs": [{"pred_logits": a, "pred_masks": b} for a, b in zip(classes[:-1], masks[:-1])],
                "middles_attn_mask": middles}


class ComboR50(nn.Module):
    """MaskFormer (maskformer_model.py:274-391) with COMBO-R50's configuration: image backbone + pre-mask backbone with channel gates
    (:335-352), frozen VGGish (:327-330), MaskFormerHead = product pixel decoder -> late AVFuse on the mask features -> audio MLP ->
    masked transformer decoder (mask_former_head.py:141-159)."""

    def __init__(self, num_classes=2, num_frames=5, backbone="r50"):
        """backbone: "r50" (configs/avs_s4, avs_ms3 R50 settings) or "pvtv2_b5" (BASELINE configs 4 / 5: harness/pvtv2.py; the pre-mask
        branch uses the same backbone type, PRE_SAM_DIM [64, 128, 320, 512]).  num_classes 2 (S4 / MS3: background + object), 71 (AVSS)."""
        super().__init__()
        self.num_frames = num_frames
        if backbone == "pvtv2_b5":
            from .pvtv2 import PVTV2_B5_FEATURES, PVTv2B5Features
            features, make = PVTV2_B5_FEATURES, PVTv2B5Features
        else:
            features, make = R50_FEATURES, ResNet50Features
        self.backbone = make()
        self.pre_sam_backbone = make()
        self.scale_factor_module = nn.ModuleList(ChannelGate(c) for c, _ in features.values())
        self.audio_backbone = VGGish()
        for p in self.audio_backbone.parameters():
            p.requires_grad = False
        self.pixel_decoder = MSDeformAttnPixelDecoder(features, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                                                      transformer_enc_layers=6, conv_dim=256, mask_dim=256, norm="GN",
                                                      transformer_in_features=["res3", "res4", "res5"], common_stride=4)
        self.fusion_module = AVFuse(256, 128, 8)
        self.audio_transformation = AudioMLP(128, 4096, 256)
        self.predictor = MultiScaleMaskedTransformerDecoder(num_classes=num_classes)
        self.register_buffer("pixel_mean", torch.tensor([123.675, 116.280, 103.530]).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor([58.395, 57.120, 57.375]).view(-1, 1, 1), False)

    def forward(self, images, pre_masks, audio_log_mel):
        """images / pre_masks [B*T, 3, H, W] (0..255), audio_log_mel [B*T, 1, 96, 64] -> decoder outputs (dict)."""
        with torch.no_grad():
            audio = self.audio_backbone(audio_log_mel).unsqueeze(1)                          # [B*T, 1, 128]
        feats = self.backbone((images - self.pixel_mean) / self.pixel_std)
        pre = self.pre_sam_backbone((pre_masks - self.pixel_mean) / self.pixel_std)
        feats = {k: feats[k] + gate(pre[k]) * pre[k] for (k, gate) in zip(feats.keys(), self.scale_factor_module)}
        mask_features, _, multi_scale = self.pixel_decoder.forward_features(feats)
        fused, audio = self.fusion_module(mask_features, audio)
        audio = self.audio_transformation(audio)
        return self.predictor(multi_scale, audio, fused)


def build_optimizer(model, base_lr=1e-4, weight_decay=0.05, backbone_multiplier=0.1):
    """train_net.py:147-226 (AdamW, lr x0.1 on the backbones, no decay on norms / embeddings) -- with the parameters GROUPED by
    hyper-parameters (4 groups, one fused multi-tensor AdamW launch each) instead of detectron2's one group per parameter: same update
    rule, a few hundred fewer kernel launches per step."""
    norm_types = (nn.BatchNorm2d, nn.GroupNorm, nn.LayerNorm)
    groups = {}
    seen = set()
    for module_name, module in model.named_modules():
        for _, p in module.named_parameters(recurse=False):
            if not p.requires_grad or p in seen:
                continue
            seen.add(p)
            lr = base_lr * (backbone_multiplier if "backbone" in module_name else 1.0)
            wd = 0.0 if isinstance(module, norm_types + (nn.Embedding,)) else weight_decay
            groups.setdefault((lr, wd), []).append(p)
    return torch.optim.AdamW([{"params": ps, "lr": lr, "weight_decay": wd} for (lr, wd), ps in groups.items()], lr=base_lr, fused=torch.cuda.is_available())


def clip_full_model_gradients(model, max_norm=0.01):
    """FullModelGradientClippingOptimizer (train_net.py:205-209): one global L2 norm over every gradient, multi-tensor (foreach) form."""
    return torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], max_norm, foreach=True)


def synthetic_clips(n_clips, num_frames=5, size=224, device="cpu", seed=0, supervised_frames=1, num_classes=2):
    """AVSBench-shaped synthetic batch (maskformer_model.py:316-323): uint8-range frames and SAM pre-masks, VGGish log-mel patches, and
    ground truth for the supervised frames only -- S4 supervises the FIRST frame of each clip (register_avss4_sem.py:18,49;
    criterion.py:241-254), MS3 all five.  Each supervised frame carries a background (label 0) and an object (label 1) mask."""
    g = torch.Generator().manual_seed(seed)
    n = n_clips * num_frames
    images = (torch.rand(n, 3, size, size, generator=g) * 255).to(device)
    pre_masks = (torch.rand(n, 1, size, size, generator=g) > 0.7).float().expand(-1, 3, -1, -1).contiguous().mul(255).to(device)
    audio = torch.randn(n, 1, 96, 64, generator=g).to(device)
    targets = []
    for _ in range(n_clips * supervised_frames):
        cy, cx, r = (torch.rand(3, generator=g) * torch.tensor([size, size, size / 3.0]) + torch.tensor([0.0, 0.0, size / 8.0])).tolist()
        yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
        obj = ((yy - cy) ** 2 + (xx - cx) ** 2 < r * r)
        label = 1 if num_classes <= 2 else int(torch.randint(1, num_classes, (1,), generator=g))
        targets.append({"labels": torch.tensor([0, label], device=device), "masks": torch.stack([~obj, obj]).to(device)})
    return images, pre_masks, audio, targets
