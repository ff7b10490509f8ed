#!/usr/bin/env python3
"""Micro-benchmark of the wavefront loop filter: time per macroblock step (wide, 1-row frames) and
per-row hand-off lag (narrow, tall frames).  Prints one line per shape.  GPU only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vp8oclenc_b200 import host as eng  # noqa: E402
from vp8oclenc_b200.hostlogic import make_segment_data  # noqa: E402


def probe(w, h, reps=20, luma_only=False):
    M = (w // 16) * (h // 16)
    r = np.random.default_rng(1)
    y = torch.from_numpy(r.integers(100, 140, size=(h, w)).astype(np.uint8)).cuda()
    u = torch.from_numpy(r.integers(100, 140, size=(h // 2, w // 2)).astype(np.uint8)).cuda()
    v = u.clone()
    seg = torch.zeros(M, dtype=torch.int32, device="cuda")
    mask = torch.full((M,), -1, dtype=torch.int32, device="cuda")
    sd = torch.from_numpy(make_segment_data(lf_level=(20, 20, 20, 20))).cuda()
    def run():
        if luma_only == 2:
            eng.loop_filter_frame(u, seg, mask, sd, w // 2, h // 2, 8)
        elif luma_only:
            eng.loop_filter_frame(y, seg, mask, sd, w, h, 16)
        else:
            eng.loop_filter_planes(y, u, v, seg, mask, sd, w, h)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1000.0 / reps


if __name__ == "__main__":
    if len(sys.argv) > 1:  # "WxH [mode]": a few launches of one shape (for ncu)
        w, h = map(int, sys.argv[1].split("x"))
        print(probe(w, h, reps=3, luma_only=int(sys.argv[2]) if len(sys.argv) > 2 else 0))
        sys.exit(0)
    for w, h in ((1920, 16), (1920, 32), (1920, 64), (1920, 128), (1920, 256), (1920, 544), (1920, 1088), (32, 1088), (3840, 2160)):
        t = probe(w, h)
        t2 = probe(w, h, luma_only=True)
        t3 = probe(w, h, luma_only=2)
        print("%5dx%-5d  mb %4dx%-4d  3 planes %8.1f us   luma only %8.1f us  one chroma %8.1f us" % (w, h, w // 16, h // 16, t, t2, t3))
