#!/usr/bin/env python3
"""times vp8b200_intra_frame on synthetic frames (CUDA events), prints time per frame and per wavefront stage. GPU only."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_y4m
from vp8oclenc_b200 import host as eng
for w, h in ((1920, 16), (1920, 32), (1920, 1088), (3840, 2160)):
    y, u, v = (torch.from_numpy(np.ascontiguousarray(p).reshape(-1)).cuda() for p in gen_y4m.Clip(w, max(h, 128)).frame(1))
    y, u, v = y[: w * h].contiguous(), u[: w * h // 4].contiguous(), v[: w * h // 4].contiguous()
    M = (w // 16) * (h // 16)
    out = [torch.zeros(n, dtype=t, device="cuda") for n, t in ((w * h, torch.uint8), (w * h // 4, torch.uint8), (w * h // 4, torch.uint8), (M * 400, torch.int16), (M * 16, torch.int32), (M, torch.int32), (M, torch.int32))]
    for _ in range(2):
        keep = eng.intra_frame(y, u, v, *out, w, h, (19, 24, 7, 10))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        keep = eng.intra_frame(y, u, v, *out, w, h, (19, 24, 7, 10))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    stages = w // 16 + 2 * (h // 16 - 1)
    print("%dx%d: %.3f ms per frame, %d stages, %.2f us per stage" % (w, h, ms, stages, 1000 * ms / stages))
