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


def probe_fused(L, flush, st):
    """the fused predict/transform/SSIM kernel on a static 1080p frame pair with random vectors"""
    sys.path.insert(0, ROOT)
    from vp8oclenc_b200.hostlogic import make_segment_data
    r = np.random.default_rng(5)
    M = (W // 16) * (H // 16)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    mk = lambda h, w: torch.from_numpy(r.integers(0, 255, size=(h, w)).astype(np.uint8)).cuda()
    cur = [mk(H, W), mk(H // 2, W // 2), mk(H // 2, W // 2)]
    img = [mk(H, W), mk(H // 2, W // 2), mk(H // 2, W // 2)]
    rec = [torch.zeros_like(t) for t in cur]
    imgs = (ctypes.c_void_p * 9)(*([P(t) for t in img] + [None] * 6))
    ref_frame = torch.zeros(M, dtype=torch.int32, device="cuda")
    vec = torch.from_numpy(r.integers(-9, 10, size=(M, 4, 2)).astype(np.int16)).cuda()
    parts = torch.from_numpy(r.integers(0, 2, size=M).astype(np.int32)).cuda()
    MB = torch.zeros(M * 400, dtype=torch.int16, device="cuda")
    seg = torch.zeros(M, dtype=torch.int32, device="cuda")
    ssim = torch.full((M,), -2.0, dtype=torch.float32, device="cuda")
    sd = torch.from_numpy(make_segment_data((24, 24, 24, 24))).cuda()
    fn = lambda: L.vp8b200_mb_predict_transform_fused(st, P(cur[0]), P(cur[1]), P(cur[2]), imgs, P(ref_frame), P(vec), P(parts), P(MB),
                                                      P(seg), P(ssim), P(rec[0]), P(rec[1]), P(rec[2]), P(sd), ctypes.c_float(-1.0), W, H)
    for _ in range(3):
        ssim.fill_(-2.0)
        fn()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        ssim.fill_(-2.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1000)
    return {"fused": sorted(ts)[len(ts) // 2], "fused_x20": batch(fn)}


def batch(fn, n=20, reps=5):
    """average of n back-to-back launches (warm L2): finer than the ~2 us resolution of one event pair"""
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1000 / n)
    return sorted(ts)[len(ts) // 2]


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
    res.update(probe_fused(L, flush, st))
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
        res[name + "_x20"] = batch(fn)
    return res


if __name__ == "__main__":
    paths = sys.argv[1:] or [os.path.join(ROOT, "vp8oclenc_b200", "lib", "libvp8b200.so")]
    for p in paths:
        print(os.path.basename(p), {k: round(v, 2) for k, v in probe(p).items()}, flush=True)
