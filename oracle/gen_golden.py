#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF -- run in the build container only.

This script imports the reference's own, unmodified ms_deform_attn_core_pytorch and MSDeformAttn from
/root/reference (ops/functions/ms_deform_attn_func.py:53-72, ops/modules/ms_deform_attn.py:34-129),
with the compiled-extension import stubbed (the guard at ms_deform_attn_func.py:21-29 only needs the
module name to resolve), runs them on CPU in float64 and stores inputs, outputs and autograd gradients.
/root/reference does not exist on the GPU box, hence the committed fixtures.

Cases follow the reference's only test artefact, ops/test.py:
  testpy_fwd_double / testpy_fwd_float : seed 3, N,M,D=1,2,2, Lq,L,P=2,2,2, shapes (6,4),(3,2)  (test.py:24-71);
                                         the RNG draw order of test.py is reproduced, so these are the very
                                         tensors test.py feeds the CUDA op.
  testpy_grad_D{30,32,64,71}           : the gradcheck shapes of test.py:74-96 (values + analytic grads)
plus what test.py does not cover: the S4 level pyramid with D=32, P=4, out-of-range and
border-straddling sample locations, N>1, and the MSDeformAttn module end to end.

Usage: python oracle/gen_golden.py   (writes tests/golden/, prints a manifest)
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/models/modeling/pixel_decoder"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch
        from ops.modules import MSDeformAttn
    return ms_deform_attn_core_pytorch, MSDeformAttn


def import_reference_encoder():
    """The reference's encoder classes (pixel_decoder/msdeformattn.py:23-166).  The file imports detectron2 / fvcore at the
    top for the pixel-decoder class further down; those names are stubbed (the encoder classes only use torch), and the two
    sibling packages are registered as bare namespace packages so that their __init__ files (which pull in the whole model)
    do not run."""
    import importlib
    root = "/root/reference/models/modeling"
    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    pkg("refmodeling", root)
    pkg("refmodeling.pixel_decoder", root + "/pixel_decoder")
    pkg("refmodeling.transformer_decoder", root + "/transformer_decoder")
    pkg("refmodeling.pixel_decoder.ops", root + "/pixel_decoder/ops")
    import_reference()  # registers the extension stub
    import collections

    class Conv2d(torch.nn.Conv2d):  # detectron2.layers.Conv2d: conv, then the attached norm, then the activation
        def __init__(self, *args, **kwargs):
            norm, activation = kwargs.pop("norm", None), kwargs.pop("activation", None)
            super().__init__(*args, **kwargs)
            self.norm, self.activation = norm, activation

        def forward(self, x):
            x = torch.nn.functional.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
            if self.norm is not None:
                x = self.norm(x)
            return self.activation(x) if self.activation is not None else x

    def get_norm(norm, channels):  # detectron2.layers.get_norm for the two settings COMBO uses
        return None if norm in (None, "") else {"GN": lambda c: torch.nn.GroupNorm(32, c)}[norm](channels)

    def c2_xavier_fill(module):  # fvcore.nn.weight_init.c2_xavier_fill
        torch.nn.init.kaiming_uniform_(module.weight, a=1)
        if module.bias is not None:
            torch.nn.init.constant_(module.bias, 0)

    ShapeSpec = collections.namedtuple("ShapeSpec", ["channels", "height", "width", "stride"], defaults=(None, None, None, None))
    for name, attrs in (("fvcore", {}), ("fvcore.nn", {}), ("fvcore.nn.weight_init", {"c2_xavier_fill": c2_xavier_fill}),
                        ("detectron2", {}), ("detectron2.config", {"configurable": lambda f: f}),
                        ("detectron2.layers", {"Conv2d": Conv2d, "ShapeSpec": ShapeSpec, "get_norm": get_norm}),
                        ("detectron2.modeling", {"SEM_SEG_HEADS_REGISTRY": types.SimpleNamespace(register=lambda: (lambda c: c))})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
    sys.modules["fvcore.nn"].weight_init = sys.modules["fvcore.nn.weight_init"]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("refmodeling.pixel_decoder.msdeformattn")
    return mod.MSDeformAttnTransformerEncoderOnly


def level_start(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def run_case(core, value, shapes, loc, attn, grad_out=None):
    """fp64 forward + autograd backward through the reference's torch path."""
    v = value.double().clone().requires_grad_(True)
    l = loc.double().clone().requires_grad_(True)
    a = attn.double().clone().requires_grad_(True)
    out = core(v, shapes, l, a)
    if grad_out is None:
        grad_out = torch.ones_like(out)
    out.backward(grad_out.double())
    return out.detach(), v.grad, l.grad, a.grad


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def draw(N, S, M, D, Lq, L, P):
    """Same draw order and distributions as test.py:36-39."""
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    attn = torch.rand(N, Lq, M, L, P) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value, loc, attn


def main():
    os.makedirs(OUT, exist_ok=True)
    core, MSDeformAttn = import_reference()
    torch.set_num_threads(1)

    # ---- the reference's own test.py, reproduced draw by draw -------------------------------------------
    N, M, D = 1, 2, 2
    Lq, L, P = 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    for tag in ("double", "float"):  # test.py runs the double check first, then the float check
        value, loc, attn = draw(N, S, M, D, Lq, L, P)
        out, gv, gl, ga = run_case(core, value, shapes, loc, attn)
        out32 = core(value, shapes, loc, attn)  # what test.py:62 computes in float
        save("testpy_fwd_" + tag, value=value, shapes=shapes, start=level_start(shapes), loc=loc, attn=attn,
             out=out, out_f32path=out32, grad_value=gv, grad_loc=gl, grad_attn=ga, im2col_step=2)
    for ch in (30, 32, 64, 71):  # test.py:95 (the 1025/2048/3096 widths are generated procedurally in tests)
        value, loc, attn = draw(N, S, M, ch, Lq, L, P)
        out, gv, gl, ga = run_case(core, value, shapes, loc, attn)
        save("testpy_grad_D%d" % ch, value=value, shapes=shapes, start=level_start(shapes), loc=loc, attn=attn,
             out=out, grad_value=gv, grad_loc=gl, grad_attn=ga, im2col_step=2)

    # ---- S4 pyramid (configs/avs_s4: 224x224, strides 32/16/8), D=32, P=4 ----------------------------------
    shapes = torch.as_tensor([(7, 7), (14, 14), (28, 28)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    g = torch.Generator().manual_seed(1234)
    N, M, D, Lq, L, P = 1, 2, 32, 96, 3, 4
    value = torch.rand(N, S, M, D, generator=g) * 0.01
    attn = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    grad_out = torch.rand(N, Lq, M * D, generator=g)
    # (a) locations as in test.py
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
    out, gv, gl, ga = run_case(core, value, shapes, loc, attn, grad_out)
    save("s4pyr_unit", value=value, shapes=shapes, start=level_start(shapes), loc=loc, attn=attn, grad_out=grad_out,
         out=out, grad_value=gv, grad_loc=gl, grad_attn=ga, im2col_step=128)
    # (b) edge set: locations in [-0.3, 1.3] -> out-of-range samples and border-straddling footprints
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.6 - 0.3
    out, gv, gl, ga = run_case(core, value, shapes, loc, attn, grad_out)
    save("s4pyr_edge", value=value, shapes=shapes, start=level_start(shapes), loc=loc, attn=attn, grad_out=grad_out,
         out=out, grad_value=gv, grad_loc=gl, grad_attn=ga, im2col_step=128)

    # ---- non-square levels, odd sizes, N=3, P=3, L=4 --------------------------------------------------------
    shapes = torch.as_tensor([(5, 9), (3, 4), (2, 2), (1, 1)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    N, M, D, Lq, L, P = 3, 3, 8, 17, 4, 3
    value = torch.randn(N, S, M, D, generator=g)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.4 - 0.2
    attn = torch.randn(N, Lq, M, L, P, generator=g)  # unnormalised, signed: the op does not require a softmax
    grad_out = torch.randn(N, Lq, M * D, generator=g)
    out, gv, gl, ga = run_case(core, value, shapes, loc, attn, grad_out)
    save("ragged_levels", value=value, shapes=shapes, start=level_start(shapes), loc=loc, attn=attn, grad_out=grad_out,
         out=out, grad_value=gv, grad_loc=gl, grad_attn=ga, im2col_step=3)

    # ---- MSDeformAttn module (the caller contract, ops/modules/ms_deform_attn.py:86-129) -----------------------
    torch.manual_seed(7)
    shapes = torch.as_tensor([(4, 4), (8, 8), (16, 16)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    N, C = 2, 32
    mod = MSDeformAttn(d_model=C, n_levels=3, n_heads=2, n_points=4).double()
    with torch.no_grad():  # move off the zero-init so offsets / weights depend on the query
        mod.sampling_offsets.weight.normal_(0, 0.05)
        mod.attention_weights.weight.normal_(0, 0.5)
        mod.attention_weights.bias.normal_(0, 0.5)
    query = torch.randn(N, S, C).double()
    src = torch.randn(N, S, C).double()
    ref2 = torch.rand(N, S, 3, 2).double()
    ref4 = torch.cat([ref2, torch.rand(N, S, 3, 2).double() * 0.3 + 0.05], -1)
    mask = torch.rand(N, S) < 0.1
    state = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    start = level_start(shapes)
    y2 = mod(query, ref2, src, shapes, start, None)        # CPU -> bare-except torch path (ms_deform_attn.py:123-125)
    y2m = mod(query, ref2, src, shapes, start, mask)
    y4 = mod(query, ref4, src, shapes, start, None)
    save("module_small", query=query, src=src, ref2=ref2, ref4=ref4, mask=mask, shapes=shapes, start=start,
         y2=y2.detach(), y2_masked=y2m.detach(), y4=y4.detach(),
         **{"sd_" + k.replace(".", "__"): v for k, v in state.items()})
    # ---- the module at the production head width (d_model 256, 8 heads -> D = 32) with gradients: pins the fused prologue ------
    torch.manual_seed(11)
    shapes = torch.as_tensor([(4, 4), (6, 6), (10, 10)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    N, C = 1, 256
    mod = MSDeformAttn(d_model=C, n_levels=3, n_heads=8, n_points=4)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
        mod.attention_weights.bias.normal_(0, 0.5)
    query = torch.randn(N, S, C)          # float32-representable inputs and weights, evaluated by the reference in float64
    src = torch.randn(N, S, C)
    ref2 = torch.rand(N, S, 3, 2)
    mask = torch.rand(N, S) < 0.1
    wout = torch.randn(N, S, C)
    state = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    mod = mod.double()
    start = level_start(shapes)
    qd, sd = query.double().requires_grad_(True), src.double().requires_grad_(True)
    y = mod(qd, ref2.double(), sd, shapes, start, mask)
    (y * wout.double()).sum().backward()
    save("module_d256", query=query, src=src, ref2=ref2, mask=mask, shapes=shapes, start=start, wout=wout, y=y.detach(),
         grad_query=qd.grad, grad_src=sd.grad,
         grad_offsets_bias=mod.sampling_offsets.bias.grad, grad_attn_bias=mod.attention_weights.bias.grad,
         grad_value_bias=mod.value_proj.bias.grad, grad_offsets_weight_rowsum=mod.sampling_offsets.weight.grad.sum(1),
         grad_attn_weight_rowsum=mod.attention_weights.weight.grad.sum(1),
         **{"sd_" + k.replace(".", "__"): v for k, v in state.items()})
    # ---- the caller: one encoder layer stack of the pixel decoder (msdeformattn.py:23-166), reference classes, float64 -------
    Enc = import_reference_encoder()
    torch.manual_seed(13)
    enc = Enc(d_model=256, nhead=8, num_encoder_layers=2, dim_feedforward=32, dropout=0.0, activation="relu",
              num_feature_levels=3, enc_n_points=4)
    with torch.no_grad():
        for layer in enc.encoder.layers:  # the attention weights of module_d256 (stored there), small random FFN / norms
            layer.self_attn.load_state_dict(state)
            layer.norm1.weight.normal_(1, 0.1); layer.norm1.bias.normal_(0, 0.1)
            layer.norm2.weight.normal_(1, 0.1); layer.norm2.bias.normal_(0, 0.1)
    extra = {k: v.detach().clone() for k, v in enc.state_dict().items() if ".self_attn." not in k}
    srcs = [torch.randn(1, 256, int(h), int(w)) for h, w in shapes.tolist()]
    poss = [torch.randn(1, 256, int(h), int(w)) * 0.5 for h, w in shapes.tolist()]
    wmem = torch.randn(1, S, 256)
    enc = enc.double()
    srcs_d = [x.double().requires_grad_(True) for x in srcs]
    memory, shp_out, start_out = enc(srcs_d, [p.double() for p in poss])
    (memory * wmem.double()).sum().backward()
    save("encoder_d256", memory=memory.detach(), wmem=wmem, shapes_out=shp_out, start_out=start_out,
         grad_level_embed=enc.level_embed.grad, grad_norm2_bias_l1=enc.encoder.layers[1].norm2.bias.grad,
         grad_linear1_bias_l0=enc.encoder.layers[0].linear1.bias.grad,
         **{"src%d" % i: x for i, x in enumerate(srcs)}, **{"pos%d" % i: x for i, x in enumerate(poss)},
         **{"grad_src%d" % i: x.grad for i, x in enumerate(srcs_d)},
         **{"sd_" + k.replace(".", "__"): v for k, v in extra.items()})
    # ---- the pixel decoder (msdeformattn.py:169-368): the reference class, with detectron2's Conv2d / get_norm / c2_xavier_fill
    # / ShapeSpec stubbed by their documented semantics (import_reference_encoder registers the stubs) -----------------------------
    Dec = sys.modules["refmodeling.pixel_decoder.msdeformattn"].MSDeformAttnPixelDecoder
    Spec = sys.modules["detectron2.layers"].ShapeSpec
    torch.manual_seed(17)
    in_shape = {"res2": Spec(channels=16, stride=4), "res3": Spec(channels=24, stride=8), "res4": Spec(channels=32, stride=16),
                "res5": Spec(channels=48, stride=32)}
    dec = Dec(in_shape, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=32, transformer_enc_layers=2,
              conv_dim=256, mask_dim=64, norm="GN", transformer_in_features=["res3", "res4", "res5"], common_stride=4)
    enc_state = {k: v for k, v in enc.float().state_dict().items()}           # the 2-layer encoder of the case above
    dec.transformer.load_state_dict(enc_state)
    fa, fb = torch.randn(256, 4) * 0.2, torch.randn(4, 256 * 9) * 0.2        # the 3x3 output conv as a rank-4 product (small fixture)
    with torch.no_grad():
        dec.layer_1.weight.copy_((fa @ fb).view(256, 256, 3, 3))
        for gn in [m for m in dec.modules() if isinstance(m, torch.nn.GroupNorm)]:
            gn.weight.normal_(1, 0.1); gn.bias.normal_(0, 0.1)
    small = {k: v.detach().clone() for k, v in dec.state_dict().items() if not k.startswith("transformer.") and k != "layer_1.weight"}
    feats = {"res2": torch.randn(1, 16, 20, 24), "res3": torch.randn(1, 24, 10, 12), "res4": torch.randn(1, 32, 5, 6), "res5": torch.randn(1, 48, 3, 3)}
    dec = dec.double()
    fd = {k: v.double().requires_grad_(True) for k, v in feats.items()}
    class KeepDouble:  # forward_features calls .float() on every input (msdeformattn.py:321,346); keep the oracle run in float64
        def __init__(self, t):
            self.t = t

        def float(self):
            return self.t
    mask_feat, low, multi = dec.forward_features({k: KeepDouble(v) for k, v in fd.items()})
    wm = torch.randn(mask_feat.shape)
    ((mask_feat * wm.double()).sum() + low.sum() + sum(m.square().sum() for m in multi)).backward()
    save("pixel_decoder_d256", mask_features=mask_feat.detach(), low=low.detach(), wm=wm, layer1_fa=fa, layer1_fb=fb,
         **{"multi%d" % i: m.detach() for i, m in enumerate(multi)}, **{"feat_" + k: v for k, v in feats.items()},
         **{"grad_" + k: v.grad for k, v in fd.items()}, **{"sd_" + k.replace(".", "__"): v for k, v in small.items()})
    fresh = MSDeformAttn(d_model=256, n_levels=3, n_heads=8, n_points=4)
    save("module_init_256", offsets_bias=fresh.sampling_offsets.bias.detach(),
         offsets_weight_absmax=fresh.sampling_offsets.weight.detach().abs().max(),
         attn_weight_absmax=fresh.attention_weights.weight.detach().abs().max(),
         attn_bias_absmax=fresh.attention_weights.bias.detach().abs().max(),
         value_bias_absmax=fresh.value_proj.bias.detach().abs().max(),
         output_bias_absmax=fresh.output_proj.bias.detach().abs().max(),
         value_weight_bound=np.sqrt(6.0 / (256 + 256)),
         n_params=sum(p.numel() for p in fresh.parameters()))


if __name__ == "__main__":
    main()
