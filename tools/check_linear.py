#!/usr/bin/env python
"""GPU: the tcgen05 Linear kernel (msda_linear_forward) against a float64 product, and timed against cuBLAS fp32 / TF32 on the encoder's
GEMM shapes.  usage: python tools/check_linear.py [tf32]   (tf32: skip the 3xTF32 split -- bring-up)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200  # noqa: F401
from combo_avs_b200 import linear as L

tf32_only = len(sys.argv) > 1 and sys.argv[1] == "tf32"
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
ok = True


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (M, N, K, relu) in [(128, 256, 32, False), (128, 16, 32, False), (200, 96, 64, False), (1029, 192, 256, True), (5145, 256, 1024, False),
                        (41160, 256, 256, False), (41160, 1024, 256, True), (41160, 256, 1024, False), (41160, 192, 256, False), (41160, 96, 256, False),
                        (1024, 256, 41160, False), (256, 1024, 41160, False), (256, 256, 5144, False), (96, 256, 41160, False), (300, 48, 100, False)]:
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    y = L.linear_forward(x, w, b, relu=relu, tf32_only=tf32_only)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    cub = torch.nn.functional.linear(x, w, b)
    err_cublas = float((cub.double().clamp_min(0) if relu else cub.double()).sub(ref).abs().max() / ref.abs().max())
    tol = 2e-3 if tf32_only else 1e-5
    good = err < tol
    ok &= good
    line = "M=%6d N=%4d K=%4d relu=%d  rel err %.2e (cuBLAS fp32 %.2e) %s" % (M, N, K, relu, err, err_cublas, "ok" if good else "MISMATCH")
    if M * N * K >= 10 ** 9:
        t_ours = timeit(lambda: L.linear_forward(x, w, b, relu=relu, tf32_only=tf32_only))
        t_cublas = timeit(lambda: torch.nn.functional.linear(x, w, b))
        torch.backends.cuda.matmul.allow_tf32 = True
        t_tf32 = timeit(lambda: torch.nn.functional.linear(x, w, b))
        torch.backends.cuda.matmul.allow_tf32 = False
        fl = 2.0 * M * N * K
        line += "   ours %.1f us (%.0f TFLOP/s)  cuBLAS fp32 %.1f us  cuBLAS TF32 %.1f us" % (t_ours, fl / t_ours / 1e6, t_cublas, t_tf32)
    print(line, flush=True)
# the weight gradient with both operands read in place (MN-major): grad_w[N, K] = g[M, N]^T @ x[M, K]
for (M, N, K) in [(64, 32, 32), (159, 32, 256), (1029, 64, 32), (1029, 100, 96), (41160, 1024, 256), (41160, 256, 1024), (41160, 192, 256), (41160, 96, 256)]:
    gen = torch.Generator().manual_seed(M + N + K)
    gy = torch.randn(M, N, generator=gen).to(dev)
    x = torch.randn(M, K, generator=gen).to(dev)
    gw = L.linear_wgrad(gy, x)
    torch.cuda.synchronize()
    ref = gy.double().t() @ x.double()
    err = float((gw.double() - ref).abs().max() / ref.abs().max())
    err_cublas = float(((gy.t() @ x).double() - ref).abs().max() / ref.abs().max())
    good = err < 2e-5
    ok &= good
    line = "wgrad M=%6d N=%4d K=%4d  rel err %.2e (cuBLAS fp32 %.2e) %s" % (M, N, K, err, err_cublas, "ok" if good else "MISMATCH")
    if M > 40000:
        line += "   ours %.1f us  cuBLAS fp32 %.1f us" % (timeit(lambda: L.linear_wgrad(gy, x)), timeit(lambda: gy.t() @ x))
    print(line, flush=True)

# the autograd function against F.linear (forward, grad_x, grad_weight, grad_bias) on an encoder-layer shape
torch.manual_seed(0)
for relu in (False, True):
    x = torch.randn(8, 1029, 256, device=dev, requires_grad=True)
    w = (torch.randn(1024, 256, device=dev) / 16).requires_grad_(True)
    b = torch.randn(1024, device=dev, requires_grad=True)
    gy = torch.randn(8, 1029, 1024, device=dev)
    y = L.linear(x, w, b, relu=relu)
    y.backward(gy)
    got = [y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone()]
    x.grad = w.grad = b.grad = None
    xd, wd, bd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    yr = torch.nn.functional.linear(xd, wd, bd)
    if relu:  # the SAME active set as the kernel's output: a pre-activation within rounding of zero may legitimately land on either side
        yr = yr * (y.detach() > 0)
    yr.backward(gy.double())
    for name, a, r in zip(("y", "grad_x", "grad_w", "grad_b"), got, (yr.detach(), xd.grad, wd.grad, bd.grad)):
        err = float((a.double() - r).abs().max() / r.abs().max())
        good = err < 1e-5
        ok &= good
        print("autograd relu=%d %-7s rel err %.2e %s" % (relu, name, err, "ok" if good else "MISMATCH"), flush=True)
sys.exit(0 if ok else 1)
