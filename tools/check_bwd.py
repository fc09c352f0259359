#!/usr/bin/env python
"""GPU: compare one backward variant against another on the S4 shape (several N, in-range and edge locations) and time it.
usage: python tools/check_bwd.py <variant> [baseline_variant=sorted32] [frames ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200 as pkg
from bench import S4_SHAPES, synth_inputs

var = sys.argv[1]
base = sys.argv[2] if len(sys.argv) > 2 else "sorted32"
frames_list = [int(x) for x in sys.argv[3:]] or [1, 5, 7, 40]
dev = torch.device("cuda:0")
shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


ok = True
for frames in frames_list:
    for edge in (False, True):
        v, l, a, g = synth_inputs(torch, frames, 11 + frames, torch.float32, torch.float32, dev)
        if edge:
            l = (l * 1.6 - 0.3).contiguous()
        pkg.select_variant("backward", base)
        ref = pkg.ms_deform_attn_backward(v, shapes, start, l, a, g, 128)
        pkg.select_variant("backward", var)
        for rep in range(3):
            got = pkg.ms_deform_attn_backward(v, shapes, start, l, a, g, 128)
            torch.cuda.synchronize()
            errs = [rel(x, y) for x, y in zip(got, ref)]
            # grad_loc / grad_attn of one sample do not depend on summation order: they must agree to rounding of the 8-lane sums
            tol = float(os.environ.get("CHECK_TOL", "2e-6"))
            good = errs[0] < tol and errs[1] < tol and errs[2] < tol
            ok &= good
            print("N=%d edge=%d rep=%d %s vs %s: gv %.2e gl %.2e ga %.2e %s" % (frames, edge, rep, pkg.last_variant("backward"), base, *errs, "ok" if good else "MISMATCH"), flush=True)
sets = [synth_inputs(torch, 40, 3 + i, torch.float32, torch.float32, dev) for i in range(4)]
for name in (base, var):
    pkg.select_variant("backward", name)
    for i in range(10):
        pkg.ms_deform_attn_backward(sets[i % 4][0], shapes, start, sets[i % 4][1], sets[i % 4][2], sets[i % 4][3], 128)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100):
        pkg.ms_deform_attn_backward(sets[i % 4][0], shapes, start, sets[i % 4][1], sets[i % 4][2], sets[i % 4][3], 128)
    e1.record()
    torch.cuda.synchronize()
    print("N=40 bwd[%s] %.1f us  env=%s" % (name, e0.elapsed_time(e1) * 10, {k: v for k, v in os.environ.items() if k.startswith("MSDA_")}), flush=True)
sys.exit(0 if ok else 1)
