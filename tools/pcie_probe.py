#!/usr/bin/env python
"""GPU: pinned-memory PCIe bandwidth of this box -- H2D alone, D2H alone, both at once -- for several copy sizes, and the
chunked duplex pattern of the host-buffer pipeline (4 copies per chunk per direction, no kernels)."""
import sys, time
import torch
dev = torch.device("cuda:0")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
big = 132 * 1000 * 1000
h_in = torch.empty(big, dtype=torch.uint8).pin_memory()
h_out = torch.empty(big, dtype=torch.uint8).pin_memory()
d_in = torch.empty(big, dtype=torch.uint8, device=dev)
d_out = torch.empty(big, dtype=torch.uint8, device=dev)


def run(h2d, d2h, piece, reps=3):
    """move the whole 132 MB buffer in pieces of `piece` bytes"""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for lo in range(0, big, piece):
            hi = min(big, lo + piece)
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
    torch.cuda.synchronize()
    return big * reps / (time.perf_counter() - t0) / 1e9


run(True, True, big, 2)
for mb in (1, 2, 4, 8, 16, 33, 132):
    piece = mb * 1000 * 1000
    print("pieces of %4d MB: H2D alone %.1f GB/s, D2H alone %.1f GB/s, both at once %.1f GB/s each direction" %
          (mb, run(True, False, piece), run(False, True, piece), run(True, True, piece)), flush=True)
