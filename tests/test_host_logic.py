"""CPU: the host-side mirror of the reference interface (module, autograd function wrapper, error behaviour)."""
import numpy as np
import pytest
import torch

from conftest import load_golden


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import combo_avs_b200
    return combo_avs_b200


def test_module_matches_reference_init(pkg):
    g = load_golden("module_init_256")
    torch.manual_seed(0)
    m = pkg.MSDeformAttn(d_model=256, n_levels=3, n_heads=8, n_points=4)
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"]) == 205600
    assert np.allclose(m.sampling_offsets.bias.detach().numpy(), g["offsets_bias"], atol=1e-6)  # ms_deform_attn.py:68-78
    assert m.sampling_offsets.weight.abs().max() == 0
    assert m.attention_weights.weight.abs().max() == 0 and m.attention_weights.bias.abs().max() == 0
    assert m.value_proj.bias.abs().max() == 0 and m.output_proj.bias.abs().max() == 0
    bound = float(g["value_weight_bound"])
    assert 0.9 * bound < m.value_proj.weight.abs().max() <= bound + 1e-7  # xavier_uniform_
    assert (m.im2col_step, m.d_model, m.n_levels, m.n_heads, m.n_points) == (128, 256, 3, 8, 4)


def test_state_dict_is_reference_compatible(pkg):
    g = load_golden("module_small")
    m = pkg.MSDeformAttn(d_model=32, n_levels=3, n_heads=2, n_points=4).double()
    ref_sd = {k[3:].replace("__", "."): torch.from_numpy(v) for k, v in g.items() if k.startswith("sd_")}
    assert set(ref_sd) == set(m.state_dict())
    m.load_state_dict(ref_sd, strict=True)


def test_constructor_errors_and_warning(pkg):
    with pytest.raises(ValueError):
        pkg.MSDeformAttn(d_model=30, n_heads=8)
    with pytest.warns(UserWarning):
        pkg.MSDeformAttn(d_model=24, n_heads=2)  # head width 12: not a power of two


def test_no_cpu_fallback(pkg):
    """The reference module's bare except reroutes CPU tensors to the PyTorch path (ms_deform_attn.py:123-125);
    the reference *extension* raises 'Not implemented on the CPU' (ms_deform_attn.h:43). We keep the latter."""
    shapes = torch.tensor([[6, 4], [3, 2]])
    start = torch.tensor([0, 24])
    args = (torch.zeros(1, 30, 2, 2), shapes, start, torch.zeros(1, 2, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2, 2), 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        pkg.MSDeformAttnFunction.apply(*args)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        pkg.ms_deform_attn_forward(*args)
    m = pkg.MSDeformAttn(d_model=4, n_levels=2, n_heads=2, n_points=2)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 4), torch.zeros(1, 2, 2, 2), torch.zeros(1, 30, 4), shapes, start)
    with pytest.raises(ValueError, match="Last dim of reference_points"):
        m(torch.zeros(1, 2, 4), torch.zeros(1, 2, 2, 3), torch.zeros(1, 30, 4), shapes, start)


def test_contiguity_is_checked_like_the_reference(pkg):
    shapes = torch.tensor([[6, 4], [3, 2]])
    start = torch.tensor([0, 24])
    value = torch.zeros(1, 30, 2, 4)[..., ::2]
    with pytest.raises(RuntimeError, match="value tensor has to be contiguous"):
        pkg.ms_deform_attn_forward(value, shapes, start, torch.zeros(1, 2, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2, 2), 2)


def test_missing_library_is_loud(pkg, monkeypatch, tmp_path):
    from combo_avs_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libmsda_b200.so"))
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.load()


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_reference_python_files_bind_to_the_shim(pkg):
    """The reference's own ms_deform_attn_func.py, unmodified, resolves `import MultiScaleDeformableAttention` to the shim and
    reaches libmsda_b200 through it (here on CPU it must hit the extension's 'Not implemented on the CPU' error, not a
    PyTorch fallback)."""
    import importlib
    import os
    import sys
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "combo-avs_b200", "shim")
    ref = "/root/reference/models/modeling/pixel_decoder"
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "MultiScaleDeformableAttention" or k == "ops" or k.startswith("ops.")}
    sys.path[:0] = [shim, ref]
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = importlib.import_module("ops.functions.ms_deform_attn_func")
        assert mod.MSDA.ms_deform_attn_forward is pkg.ms_deform_attn_forward
        shapes, start = torch.tensor([[6, 4], [3, 2]]), torch.tensor([0, 24])
        with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
            mod.MSDeformAttnFunction.apply(torch.zeros(1, 30, 2, 2), shapes, start, torch.zeros(1, 2, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2, 2), 2)
    finally:
        sys.path.remove(shim)
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "MultiScaleDeformableAttention" or k == "ops" or k.startswith("ops.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)


def test_encoder_classes_mirror_reference_state_dict_and_reference_points(pkg):
    """combo_avs_b200.encoder: parameter names equal the reference encoder's (golden file written by the reference classes),
    and get_reference_points equals pixel centres / size (msdeformattn.py:143-157) for unit valid ratios."""
    from combo_avs_b200.encoder import MSDeformAttnTransformerEncoder, MSDeformAttnTransformerEncoderOnly
    g, gm = load_golden("encoder_d256"), load_golden("module_d256")
    enc = MSDeformAttnTransformerEncoderOnly(d_model=256, nhead=8, num_encoder_layers=2, dim_feedforward=32, dropout=0.0,
                                             num_feature_levels=3, enc_n_points=4)
    want = {k[3:].replace("__", ".") for k in g if k.startswith("sd_")}
    want |= {"encoder.layers.%d.self_attn.%s" % (i, k[3:].replace("__", ".")) for i in range(2) for k in gm if k.startswith("sd_")}
    assert set(enc.state_dict().keys()) == want
    shapes = [(2, 3), (4, 5)]
    ref = MSDeformAttnTransformerEncoder.get_reference_points(shapes, torch.ones(2, 2, 2), "cpu")
    assert ref.shape == (2, 26, 2, 2)
    assert torch.allclose(ref[0, 0, 0], torch.tensor([0.5 / 3, 0.5 / 2])) and torch.allclose(ref[1, 6 + 7, 1], torch.tensor([2.5 / 5, 1.5 / 4]))
    assert torch.equal(ref[:, :, 0], ref[:, :, 1])
    t = MSDeformAttnTransformerEncoder.get_reference_points(torch.tensor(shapes), torch.ones(2, 2, 2), "cpu")
    assert torch.equal(t, ref)


def _pixel_decoder(pkg_):
    from combo_avs_b200.pixel_decoder import MSDeformAttnPixelDecoder
    shape = {"res2": (16, 4), "res3": (24, 8), "res4": (32, 16), "res5": (48, 32)}
    return MSDeformAttnPixelDecoder(shape, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=32,
                                    transformer_enc_layers=2, conv_dim=256, mask_dim=64, norm="GN",
                                    transformer_in_features=["res3", "res4", "res5"], common_stride=4)


def test_pixel_decoder_mirrors_reference_names_and_position_encoding(pkg):
    """State-dict names equal the reference decoder's (golden written by the reference class); the sine position encoding
    equals the closed form of position_encoding.py:33-52 for an unpadded map."""
    from combo_avs_b200.pixel_decoder import PositionEmbeddingSine
    g, ge, gm = load_golden("pixel_decoder_d256"), load_golden("encoder_d256"), load_golden("module_d256")
    dec = _pixel_decoder(pkg)
    want = {k[3:].replace("__", ".") for k in g if k.startswith("sd_")} | {"layer_1.weight"}
    want |= {"transformer." + k[3:].replace("__", ".") for k in ge if k.startswith("sd_")}
    want |= {"transformer.encoder.layers.%d.self_attn.%s" % (i, k[3:].replace("__", ".")) for i in range(2) for k in gm if k.startswith("sd_")}
    assert set(dec.state_dict().keys()) == want
    assert dec.num_fpn_levels == 1 and dec.in_features == ["res2", "res3", "res4", "res5"]
    pe = PositionEmbeddingSine(4, normalize=True)(torch.zeros(1, 3, 2, 3))
    assert pe.shape == (1, 8, 2, 3)
    import math
    y = 2 * math.pi * 1 / (2 + 1e-6)                      # first row, normalised
    assert abs(float(pe[0, 0, 0, 0]) - math.sin(y)) < 1e-6 and abs(float(pe[0, 1, 0, 0]) - math.cos(y)) < 1e-6
    assert abs(float(pe[0, 2, 0, 0]) - math.sin(y / 10000 ** 0.5)) < 1e-6
