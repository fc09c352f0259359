"""Row f2: the tcgen05 / TMEM / TMA Linear kernel (msda_linear_forward, csrc/msda_linear.cu) that replaces the strict-fp32 cuBLAS GEMMs
of the deformable encoder layer (reference msdeformattn.py:119-134, ops/modules/ms_deform_attn.py:105-128).  The oracle is a float64
matrix product of the same float32 inputs; the bar is float32-level accuracy (cuBLAS fp32 itself is 2e-7 .. 1e-6 on these shapes)."""
import ctypes

import numpy as np
import pytest
import torch


def test_linear_entry_point_validates_arguments_without_a_gpu():
    from combo_avs_b200 import _lib
    L = _lib.load()
    null = None
    assert L.msda_linear_forward(null, null, null, null, -1, 16, 32, 0, 0, null) == -2      # bad dimension
    assert L.msda_linear_forward(null, null, null, null, 0, 16, 32, 0, 0, null) == 0        # empty problem
    assert L.msda_linear_forward(null, null, null, null, 4, 16, 32, 0, 0, null) == -1       # NULL tensors
    assert b"NULL" in L.msda_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,relu,bias", [
    (128, 256, 32, False, True), (1, 16, 4, False, True), (129, 48, 100, True, True), (1029, 192, 256, False, True),
    (5145, 96, 256, False, False), (2058, 1024, 256, True, True), (2058, 256, 1024, False, True),
    (256, 1024, 5144, False, False),      # weight-gradient shape: few tiles, long K -> split over the SMs, partial tiles added
])
def test_linear_matches_float64(M, N, K, relu, bias):
    from combo_avs_b200 import linear as L
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    y = L.linear_forward(x, w, b, relu=relu)
    ref = x.double() @ w.double().t() + (b.double() if bias else 0)
    if relu:
        ref = ref.clamp_min(0)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err < 5e-6, err
    plain = L.linear_forward(x, w, b, relu=relu, tf32_only=True)   # the unsplit path is TF32-accurate only: the split is what buys fp32
    assert 1e-5 < float((plain.double() - ref).abs().max() / ref.abs().max()) < 5e-3 or K <= 4


@pytest.mark.gpu
@pytest.mark.parametrize("relu", [False, True])
def test_linear_autograd_matches_float64(relu):
    from combo_avs_b200 import linear as L
    torch.manual_seed(0)
    x = torch.randn(4, 1029, 256, device="cuda", requires_grad=True)
    w = (torch.randn(1024, 256, device="cuda") / 16).requires_grad_(True)
    b = torch.randn(1024, device="cuda", requires_grad=True)
    gy = torch.randn(4, 1029, 1024, device="cuda")
    assert L.supported(x, w)
    y = L.linear(x, w, b, relu=relu)
    y.backward(gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.linear(xd, wd, bd)
    if relu:  # the same active set as the kernel's output (a pre-activation within rounding of zero may land on either side)
        yr = yr * (y.detach() > 0)
    yr.backward(gy.double())
    for got, ref in ((y, yr), (x.grad, xd.grad), (w.grad, wd.grad), (b.grad, bd.grad)):
        assert float((got.detach().double() - ref.detach()).abs().max() / ref.detach().abs().max()) < 1e-5


@pytest.mark.gpu
def test_linear_falls_back_when_it_does_not_apply():
    from combo_avs_b200 import linear as L
    x = torch.randn(7, 30, device="cuda")           # K % 4 != 0
    w = torch.randn(16, 30, device="cuda")
    assert not L.supported(x, w)
    assert torch.equal(L.linear(x, w), torch.nn.functional.linear(x, w))
    xd = torch.randn(8, 32, device="cuda", dtype=torch.float64)
    assert not L.supported(xd, torch.randn(16, 32, device="cuda", dtype=torch.float64))
    with pytest.raises(ValueError):
        L.linear_forward(x, w)


def test_linear_wgrad_entry_point_validates_arguments_without_a_gpu():
    from combo_avs_b200 import _lib
    L = _lib.load()
    null = None
    assert L.msda_linear_wgrad(null, null, null, -1, 16, 32, 0, null) == -2       # bad dimension
    assert L.msda_linear_wgrad(null, null, null, 4, 16, 32, 0, null) == -1        # NULL result
    assert L.msda_linear_wgrad_supported(0, 16, 32) == 0
    assert L.msda_linear_wgrad_supported(64, 18, 32) == 0                         # N % 4
    assert L.msda_linear_wgrad_supported(64, 16, 48) == 0                         # K % 32


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(64, 32, 32), (159, 32, 256), (1029, 100, 96), (5145, 1024, 256), (5145, 256, 1024), (5145, 96, 256)])
def test_linear_wgrad_matches_float64(M, N, K):
    """grad_weight[N, K] = grad_y[M, N]^T . x[M, K] with both operands read in place (MN-major tcgen05 operands, 32 x 32 TMA boxes,
    rows past M zero-filled by TMA), the contraction split over the SMs."""
    from combo_avs_b200 import linear as L
    g = torch.Generator().manual_seed(M + N + K)
    gy = torch.randn(M, N, generator=g).cuda()
    x = torch.randn(M, K, generator=g).cuda()
    gw = L.linear_wgrad(gy, x)
    ref = gy.double().t() @ x.double()
    err = float((gw.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err
