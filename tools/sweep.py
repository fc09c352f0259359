#!/usr/bin/env python
"""GPU: the standalone sweep of SURVEY.md section 8(d), config 2 -- N in {5,10,20,40,80} x location sets x dtypes.

Location sets (labelled separately, as the survey asks):
  uniform  -- the reference test's inputs (ops/test.py:36-39): loc ~ U[0,1)  (no locality: the worst case for the sort)
  model    -- what the pixel decoder produces at initialisation: reference points on the query's own pixel centre
              (msdeformattn.py:144-157) + the module's bias offsets (k+1)*dir_m / (W,H) (ms_deform_attn.py:68-78)
              + Gaussian noise of 0.5 pixel
  edge     -- loc ~ U[-0.3, 1.3): a third of the samples touch or leave the border
Timing: 20 warm-ups, 100 timed iterations with one CUDA event pair each -> median and min; "warm" repeats one input set
(L2 not flushed between iterations), "cold" rotates 4 input sets (every launch starts cold in the 126 MB L2).
Prints one JSON object."""
import json, math, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combo_avs_b200 as pkg
from bench import S4_SHAPES, M, D, L, P, S, LQ, algorithmic_bytes, load_peaks

dev = torch.device("cuda:0")
shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def reference_points():
    """get_reference_points (msdeformattn.py:144-157) with valid_ratios = 1: pixel centres of every level, as queries."""
    pts = []
    for h, w in S4_SHAPES:
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h) / h, torch.linspace(0.5, w - 0.5, w) / w, indexing="ij")
        pts.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), -1))
    return torch.cat(pts, 0)  # [Lq, 2] (x, y)


def make(kind, n, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    value = torch.rand(n, S, M, D, generator=g) * 0.01
    attn = torch.rand(n, LQ, M, L, P, generator=g) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    gout = torch.rand(n, LQ, M * D, generator=g)
    if kind == "uniform":
        loc = torch.rand(n, LQ, M, L, P, 2, generator=g)
    elif kind == "edge":
        loc = torch.rand(n, LQ, M, L, P, 2, generator=g) * 1.6 - 0.3
    else:
        ref = reference_points()  # the same normalised point is used on every level
        theta = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
        d = torch.stack((theta.cos(), theta.sin()), -1)
        d = d / d.abs().max(-1, keepdim=True)[0]
        off = d.view(M, 1, 1, 2) * torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, P, 1)   # [M,1,P,2] pixels
        off = off + 0.5 * torch.randn(n, LQ, M, L, P, 2, generator=g)
        norm = torch.tensor([[w, h] for h, w in S4_SHAPES], dtype=torch.float32).view(1, 1, 1, L, 1, 2)
        loc = ref.view(1, LQ, 1, 1, 1, 2) + off / norm
    ldt = torch.float32
    return [value.to(dtype).to(dev), loc.to(ldt).contiguous().to(dev), attn.to(ldt).to(dev), gout.to(dtype).to(dev)]


def times(fn, sets, iters=100, warm=20):
    for i in range(warm):
        fn(*sets[i % len(sets)])
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(*sets[i % len(sets)])
        b.record()
    torch.cuda.synchronize()
    us = [a.elapsed_time(b) * 1e3 for a, b in evs]
    return {"median_us": statistics.median(us), "min_us": min(us)}


def main():
    peak, src = load_peaks()
    frames = [int(x) for x in sys.argv[1:]] or [5, 10, 20, 40, 80]
    res = {"peak_gbs": peak, "peak_source": src, "note": __doc__.split("Prints")[0].strip().split("\n", 1)[1], "rows": []}
    fwd = lambda v, l, a, g: pkg.ms_deform_attn_forward(v, shapes, start, l, a, 128)
    bwd = lambda v, l, a, g: pkg.ms_deform_attn_backward(v, shapes, start, l, a, g, 128)
    for dtype, dname in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        for kind in ("uniform", "model", "edge"):
            for n in frames:
                if dname == "bf16" and (kind != "uniform" or n not in (10, 40)):
                    continue
                sets = [make(kind, n, 7 + i, dtype) for i in range(4)]
                bf, bb = algorithmic_bytes(n, 4 if dname == "fp32" else 2, 4)
                row = {"dtype": dname, "locations": kind, "frames": n, "fwd_bytes": bf, "bwd_bytes": bb}
                for label, ss in (("cold", sets), ("warm", sets[:1])):
                    f, b = times(fwd, ss), times(bwd, ss)
                    row[label] = {"fwd": f, "bwd": b, "fwd_bwd_median_us": f["median_us"] + b["median_us"],
                                  "frac_hbm": (bf + bb) / (f["median_us"] + b["median_us"]) / 1e3 / peak}
                row["variants"] = [pkg.last_variant("forward"), pkg.last_variant("backward")]
                res["rows"].append(row)
                print("%s %-7s N=%2d  cold fwd %.1f bwd %.1f (%.1f%%)   warm fwd %.1f bwd %.1f (%.1f%%)" % (
                    dname, kind, n, row["cold"]["fwd"]["median_us"], row["cold"]["bwd"]["median_us"], 100 * row["cold"]["frac_hbm"],
                    row["warm"]["fwd"]["median_us"], row["warm"]["bwd"]["median_us"], 100 * row["warm"]["frac_hbm"]), file=sys.stderr, flush=True)
                del sets
                torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
