#!/usr/bin/env python3
"""A/B check of two engine libraries on luma_search_2step: identical outputs on random inputs at several
sizes (including edge-heavy small ones), then timing.  usage: ab_2step.py libA.so libB.so"""
import ctypes
import subprocess
import sys

import numpy as np
import torch


def run(lib_path, w, h, seed, reps=0):
    L = ctypes.CDLL(lib_path)
    r = np.random.default_rng(seed)
    base = r.integers(0, 255, size=(h + 16, w + 16)).astype(np.uint8)
    cur = torch.from_numpy(np.ascontiguousarray(base[8:8 + h, 8:8 + w])).cuda()
    ref = torch.from_numpy(np.ascontiguousarray(base[6:6 + h, 9:9 + w])).cuda()
    nb = w * h // 64
    net = torch.from_numpy(r.integers(-6, 7, size=(nb, 2)).astype(np.int16)).cuda()
    out = torch.zeros((nb, 2), dtype=torch.int16, device="cuda")
    met = torch.zeros(nb, dtype=torch.int32, device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.vp8b200_luma_search_2step(st, P(cur), P(ref), P(net), P(out), P(met), w, h)
    torch.cuda.synchronize()
    return out.cpu().numpy().tobytes() + met.cpu().numpy().tobytes()


if __name__ == "__main__":
    if len(sys.argv) == 5:  # child: lib w h seed -> hash on stdout
        import hashlib
        print(hashlib.md5(run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))).hexdigest())
        sys.exit(0)
    a, b = sys.argv[1], sys.argv[2]
    ok = True
    for w, h, seed in ((64, 48, 1), (176, 144, 2), (352, 288, 3), (200, 120, 4), (1920, 1088, 5), (3840, 2160, 6), (16, 16, 7), (24, 8, 8)):
        ha = subprocess.run([sys.executable, __file__, a, str(w), str(h), str(seed)], capture_output=True, text=True, timeout=120).stdout.strip()
        hb = subprocess.run([sys.executable, __file__, b, str(w), str(h), str(seed)], capture_output=True, text=True, timeout=120).stdout.strip()
        print(w, h, "same" if ha == hb and ha else "DIFFERENT %s %s" % (ha, hb), flush=True)
        ok = ok and ha == hb and bool(ha)
    sys.exit(0 if ok else 1)
