#!/usr/bin/env python
"""Quick GPU timing of selected forward/backward variants on the S4 shape (no reference comparison).
usage: python tools/time_variants.py [frames] [fwd_variant] [bwd_variant]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200 as pkg
from bench import S4_SHAPES, algorithmic_bytes, load_peaks, synth_inputs

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
fv = sys.argv[2] if len(sys.argv) > 2 else "auto"
bv = sys.argv[3] if len(sys.argv) > 3 else "auto"
dev = torch.device("cuda:0")
shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
NS = int(os.environ.get('MSDA_SETS', '4'))
sets = [synth_inputs(torch, frames, 3 + i, torch.float32, torch.float32, dev) for i in range(NS)]
pkg.select_variant("forward", fv)
pkg.select_variant("backward", bv)

def timeit(fn, reps=100, warm=10):
    for i in range(warm):
        fn(*sets[i % NS])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(*sets[i % NS])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def timeit_graph(fn, reps=20):
    """GPU time without the host in the loop: NS launches captured into one CUDA graph, replayed `reps` times (small N: the Python
    call + tensor-map encode of an eager launch take longer than the kernel itself)."""
    fn(*sets[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(NS):
            fn(*sets[i])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * NS) * 1e3

if os.environ.get("MSDA_GRAPH") == "1":
    timeit = lambda fn, reps=100, warm=10: timeit_graph(fn)
f = timeit(lambda v, l, a, g: pkg.ms_deform_attn_forward(v, shapes, start, l, a, 128))
b = timeit(lambda v, l, a, g: pkg.ms_deform_attn_backward(v, shapes, start, l, a, g, 128))
peak, _ = load_peaks()
bf, bb = algorithmic_bytes(frames, 4, 4)
print("frames %d  fwd[%s] %.1f us (%.1f%%)  bwd[%s] %.1f us (%.1f%%)  total %.1f us (%.1f%% of HBM roofline)  env=%s" % (
    frames, pkg.last_variant("forward"), f, 100 * bf / f / 1e3 / peak, pkg.last_variant("backward"), b, 100 * bb / b / 1e3 / peak,
    f + b, 100 * (bf + bb) / (f + b) / 1e3 / peak, {k: v for k, v in os.environ.items() if k.startswith("MSDA_")}))
