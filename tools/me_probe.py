#!/usr/bin/env python3
"""Times the two motion-search kernels of a given engine library at 1080p (for kernel variants).
    python tools/me_probe.py [path/to/libvp8b200.so ...]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 1920, 1088


def probe(path):
    L = ctypes.CDLL(path)
    r = np.random.default_rng(3)
    base = r.integers(0, 255, size=(H + 16, W + 16)).astype(np.uint8)
    cur = torch.from_numpy(np.ascontiguousarray(base[8:8 + H, 8:8 + W])).cuda()
    ref = torch.from_numpy(np.ascontiguousarray(base[6:6 + H, 9:9 + W])).cuda()
    nb = W * H // 64
    net = torch.from_numpy(r.integers(-3, 4, size=(nb, 2)).astype(np.int16)).cuda()
    out = torch.zeros((nb, 2), dtype=torch.int16, device="cuda")
    met = torch.zeros(nb, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for name, fn in (("2step", lambda: L.vp8b200_luma_search_2step(st, P(cur), P(ref), P(net), P(out), P(met), W, H)),
                     ("1step", lambda: L.vp8b200_luma_search_1step(st, P(cur), P(ref), P(net), P(out), (W // 16) * 2, W, H, 1))):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1000)
        res[name] = sorted(ts)[len(ts) // 2]
    return res


if __name__ == "__main__":
    paths = sys.argv[1:] or [os.path.join(ROOT, "vp8oclenc_b200", "lib", "libvp8b200.so")]
    for p in paths:
        print(os.path.basename(p), {k: round(v, 1) for k, v in probe(p).items()}, flush=True)
