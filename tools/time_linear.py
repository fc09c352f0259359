#!/usr/bin/env python
"""GPU: time msda_linear_forward on one shape. usage: python tools/time_linear.py M N K [tf32]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200  # noqa
from combo_avs_b200 import linear as L
M, N, K = (int(v) for v in sys.argv[1:4])
tf32 = len(sys.argv) > 4
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
for _ in range(3): L.linear_forward(x, w, b, tf32_only=tf32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): L.linear_forward(x, w, b, tf32_only=tf32)
e1.record(); torch.cuda.synchronize()
print("M=%d N=%d K=%d %s %.1f us env=%s" % (M, N, K, "tf32" if tf32 else "3xtf32", e0.elapsed_time(e1) / 20 * 1e3, {k: v for k, v in os.environ.items() if k.startswith("MSDA_")}))
