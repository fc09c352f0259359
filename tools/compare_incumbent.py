#!/usr/bin/env python
"""GPU: time and cross-check against the REFERENCE's own CUDA extension (oracle/_ref, built by oracle/build_ref.py).

Prints one JSON object: per-variant fwd / bwd microseconds of libmsda_b200 and of the reference kernels on the same
tensors (S4 shape), and the max-norm relative differences between the two implementations' outputs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import combo_avs_b200 as pkg  # noqa: E402
from bench import S4_SHAPES, algorithmic_bytes, load_peaks, synth_inputs  # noqa: E402
import build_ref  # noqa: E402


def timeit(fn, sets, reps=50, warm=5):
    for i in range(warm):
        fn(*sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(*sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    dev = torch.device("cuda:0")
    shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    sets = [synth_inputs(torch, frames, 3 + i, torch.float32, torch.float32, dev) for i in range(4)]
    peak, _ = load_peaks()
    bf, bb = algorithmic_bytes(frames, 4, 4)
    res = {"frames": frames, "ours": {}, "reference_cuda": None}
    for v in [x for x in pkg.variant_names("forward") if x != "auto"]:
        pkg.select_variant("forward", v)
        us = timeit(lambda val, loc, att, g: pkg.ms_deform_attn_forward(val, shapes, start, loc, att, 128), sets)
        res["ours"]["fwd_" + v] = {"us": us, "frac_hbm": bf / us / 1e3 / peak}
    for v in [x for x in pkg.variant_names("backward") if x != "auto"]:
        pkg.select_variant("backward", v)
        us = timeit(lambda val, loc, att, g: pkg.ms_deform_attn_backward(val, shapes, start, loc, att, g, 128), sets)
        res["ours"]["bwd_" + v] = {"us": us, "frac_hbm": bb / us / 1e3 / peak}
    pkg.select_variant("forward", "auto")
    pkg.select_variant("backward", "auto")
    ref = build_ref.load_ref()
    if ref is not None:
        fus = timeit(lambda val, loc, att, g: ref.ms_deform_attn_forward(val, shapes, start, loc, att, 128), sets)
        bus = timeit(lambda val, loc, att, g: ref.ms_deform_attn_backward(val, shapes, start, loc, att, g.view(frames, -1, 256), 128), sets, reps=20)
        val, loc, att, g = sets[0]
        ro = ref.ms_deform_attn_forward(val, shapes, start, loc, att, 128)
        rgv, rgl, rga = ref.ms_deform_attn_backward(val, shapes, start, loc, att, g, 128)
        diffs = {}
        for fv in [x for x in pkg.variant_names("forward") if x != "auto"]:
            pkg.select_variant("forward", fv)
            diffs["fwd_" + fv] = rel(pkg.ms_deform_attn_forward(val, shapes, start, loc, att, 128), ro)
        for bv in [x for x in pkg.variant_names("backward") if x != "auto"]:
            pkg.select_variant("backward", bv)
            gv, gl, ga = pkg.ms_deform_attn_backward(val, shapes, start, loc, att, g, 128)
            diffs["bwd_" + bv] = {"grad_value": rel(gv, rgv), "grad_loc": rel(gl, rgl), "grad_attn": rel(ga, rga)}
        res["reference_cuda"] = {"fwd_us": fus, "bwd_us": bus, "fwd_frac_hbm": bf / fus / 1e3 / peak, "bwd_frac_hbm": bb / bus / 1e3 / peak,
                                 "max_rel_diff_vs_ours": diffs}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
