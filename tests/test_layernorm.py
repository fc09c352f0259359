"""Row f2, the elementwise epilogue of the encoder layer: y = LayerNorm(a + b) (msda_add_layernorm_forward / msda_layernorm_backward) and
the bias-gradient column sum (msda_colsum), csrc/msda_layernorm.cu -- reference models/modeling/pixel_decoder/msdeformattn.py:119-134
(`src = norm1(src + dropout1(src2))`, `src = norm2(src + dropout3(src2))`).  The oracle is torch's float64 layer_norm of the same float32
inputs; the bar is float32 rounding (1e-5 relative, max norm; the row-reduced parameter gradients 2e-5)."""
import pytest
import torch
import torch.nn.functional as F


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_layernorm_entry_points_validate_arguments_without_a_gpu():
    from combo_avs_b200 import _lib
    L = _lib.load()
    null = None
    assert L.msda_layernorm_supported(256) == 1 and L.msda_layernorm_supported(200) == 0 and L.msda_layernorm_supported(1024) == 0
    assert L.msda_add_layernorm_forward(null, null, null, null, null, null, null, 4, 200, 1e-5, 0, null) == -2   # unsupported width
    assert L.msda_add_layernorm_forward(null, null, null, null, null, null, null, 0, 256, 1e-5, 0, null) == 0    # empty problem
    assert L.msda_add_layernorm_forward(null, null, null, null, null, null, null, 4, 256, 1e-5, 0, null) == -1   # NULL tensors
    assert L.msda_colsum_supported(256) == 1 and L.msda_colsum_supported(6) == 0
    assert L.msda_colsum(null, null, null, 0, 4, 6, 0, null) == -2


def test_add_layer_norm_falls_back_to_torch_on_cpu():
    from combo_avs_b200.layernorm import add_layer_norm, column_sum
    torch.manual_seed(0)
    a, b = torch.randn(3, 7, 256), torch.randn(3, 7, 256)
    w, bt = torch.randn(256), torch.randn(256)
    assert torch.equal(add_layer_norm(a, b, w, bt, 1e-5), F.layer_norm(a + b, (256,), w, bt, 1e-5))
    assert torch.equal(add_layer_norm(a, None, w, bt, 1e-5), F.layer_norm(a, (256,), w, bt, 1e-5))
    assert torch.equal(column_sum(a.reshape(-1, 256)), a.reshape(-1, 256).sum(0))


@pytest.mark.gpu
@pytest.mark.parametrize("rows,C,with_b", [((1,), 128, True), ((3, 7), 256, True), ((2, 1029), 256, True), ((5, 1029), 256, False),
                                           ((4, 333), 384, True), ((6000,), 512, True)])
def test_add_layer_norm_matches_float64(rows, C, with_b):
    from combo_avs_b200.layernorm import add_layer_norm, supported
    g = torch.Generator().manual_seed(C + len(rows))
    a = (torch.randn(*rows, C, generator=g) * 2 + 0.5).cuda().requires_grad_(True)
    b = torch.randn(*rows, C, generator=g).cuda().requires_grad_(True) if with_b else None
    w = (torch.rand(C, generator=g) + 0.5).cuda().requires_grad_(True)
    bt = torch.randn(C, generator=g).cuda().requires_grad_(True)
    gy = torch.randn(*rows, C, generator=g).cuda()
    assert supported(a, b, w, bt)
    y = add_layer_norm(a, b, w, bt, 1e-5)
    y.backward(gy)
    ad, wd, btd = (t.detach().double().requires_grad_(True) for t in (a, w, bt))
    bd = b.detach().double().requires_grad_(True) if with_b else None
    yr = F.layer_norm(ad + bd if with_b else ad, (C,), wd, btd, 1e-5)
    yr.backward(gy.double())
    assert rel(y, yr) < 1e-5
    assert rel(a.grad, ad.grad) < 1e-5
    if with_b:
        assert torch.equal(a.grad, b.grad)
    assert rel(w.grad, wd.grad) < 2e-5 and rel(bt.grad, btd.grad) < 2e-5
    # against PyTorch's own float32 kernels: the same function to rounding
    yt = F.layer_norm((a + b) if with_b else a, (C,), w, bt, 1e-5)
    assert rel(y, yt) < 2e-6


@pytest.mark.gpu
def test_add_layer_norm_is_deterministic_and_graph_capturable():
    from combo_avs_b200.layernorm import add_layer_norm
    torch.manual_seed(1)
    a = torch.randn(8, 1029, 256, device="cuda", requires_grad=True)
    b = torch.randn(8, 1029, 256, device="cuda")
    w = torch.randn(256, device="cuda", requires_grad=True)
    bt = torch.randn(256, device="cuda", requires_grad=True)
    gy = torch.randn(8, 1029, 256, device="cuda")
    grads = []
    for _ in range(2):
        for t in (a, w, bt):
            t.grad = None
        add_layer_norm(a, b, w, bt).backward(gy)
        grads.append((a.grad.clone(), w.grad.clone(), bt.grad.clone()))
    assert all(torch.equal(x, y) for x, y in zip(*grads))      # fixed-order reductions: bit-identical from run to run
    static_a, static_out = a.detach().clone(), None
    g = torch.cuda.CUDAGraph()
    with torch.no_grad():
        add_layer_norm(static_a, b, w, bt)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            static_out = add_layer_norm(static_a, b, w, bt)
        static_a.copy_(a.detach() * 3)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(static_out, add_layer_norm(a.detach() * 3, b, w, bt))


@pytest.mark.gpu
@pytest.mark.parametrize("M,N", [(1, 4), (7, 96), (41160, 192), (41160, 256), (5145, 1024), (1000, 2048)])
def test_column_sum_matches_float64(M, N):
    from combo_avs_b200.layernorm import column_sum
    g = torch.randn(M, N, generator=torch.Generator().manual_seed(M + N)).cuda()
    got = column_sum(g)
    ref = g.double().sum(0)
    assert float((got.double() - ref).abs().max()) < 2e-5 * float(g.double().abs().sum(0).max())
    assert torch.equal(got, column_sum(g))
