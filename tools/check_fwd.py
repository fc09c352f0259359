#!/usr/bin/env python
"""GPU: compare one forward variant against another on the S4 shape (several N, in-range and edge locations) and time it.
usage: python tools/check_fwd.py [variant=slab32] [baseline_variant=warp32] [frames ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200 as pkg
from bench import S4_SHAPES, synth_inputs

var = sys.argv[1] if len(sys.argv) > 1 else "slab32"
base = sys.argv[2] if len(sys.argv) > 2 else "warp32"
frames_list = [int(x) for x in sys.argv[3:]] or [1, 2, 5, 7, 40]
dev = torch.device("cuda:0")
shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
ok = True
for frames in frames_list:
    for edge in (False, True):
        v, l, a, g = synth_inputs(torch, frames, 11 + frames, torch.float32, torch.float32, dev)
        if edge:
            l = (l * 1.6 - 0.3).contiguous()
        pkg.select_variant("forward", base)
        ref = pkg.ms_deform_attn_forward(v, shapes, start, l, a, 128)
        pkg.select_variant("forward", var)
        for rep in range(3):
            got = pkg.ms_deform_attn_forward(v, shapes, start, l, a, 128)
            torch.cuda.synchronize()
            err = float((got.double() - ref.double()).abs().max() / ref.double().abs().max())
            good = err < 2e-6
            ok &= good
            print("N=%d edge=%d rep=%d %s vs %s: %.2e %s" % (frames, edge, rep, pkg.last_variant("forward"), base, err, "ok" if good else "MISMATCH"), flush=True)
sets = [synth_inputs(torch, 40, 3 + i, torch.float32, torch.float32, dev) for i in range(4)]
for name in (base, var):
    pkg.select_variant("forward", name)
    for i in range(10):
        pkg.ms_deform_attn_forward(sets[i % 4][0], shapes, start, sets[i % 4][1], sets[i % 4][2], 128)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100):
        pkg.ms_deform_attn_forward(sets[i % 4][0], shapes, start, sets[i % 4][1], sets[i % 4][2], 128)
    e1.record()
    torch.cuda.synchronize()
    print("N=40 fwd[%s] %.1f us  env=%s" % (name, e0.elapsed_time(e1) * 10, {k: v for k, v in os.environ.items() if k.startswith("MSDA_")}), flush=True)
sys.exit(0 if ok else 1)
