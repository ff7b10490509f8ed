#!/usr/bin/env python3
"""A/B of the two ways luma_search_2step stages its reference windows (north star: "macroblock rows staged into shared
memory with TMA"): word loads + funnel shifts (production) against TMA 2-D box copies out of a replicate-padded plane
(vp8b200_experiment_search_2step_tma).  Same vectors and metrics (checked), CUDA-event times with the L2 flushed
between launches.  GPU only.

    python tools/me_tma_ab.py [WxH ...]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_y4m  # noqa: E402
from vp8oclenc_b200 import host as eng  # noqa: E402


def full_pel_net(cur, ref, w, h):
    """the five luma_search_1step levels, as the engine runs them (nets ping-pong, stride 2*mb_width)"""
    nb = w * h // 64
    nets = [torch.zeros((nb, 2), dtype=torch.int16, device="cuda") for _ in range(2)]
    cp, rp = [cur], [ref]
    for k in range(4):
        ww, hh = w >> k, h >> k
        for pyr in (cp, rp):
            d = torch.empty((hh // 2) * (ww // 2), dtype=torch.uint8, device="cuda")
            eng.downsample_x2(pyr[-1], d, ww, hh)
            pyr.append(d)
    for k in range(4, -1, -1):
        src = 1 if (k & 1) else 0
        eng.luma_search_1step(cp[k], rp[k], nets[src], nets[src ^ 1], (w // 16) * 2, w >> k, h >> k, 1 << k)
    return nets[1]


def ab(w, h, reps=20):
    clip = gen_y4m.Clip(w, h)
    cur = torch.from_numpy(np.ascontiguousarray(clip.frame(7)[0]).reshape(-1)).cuda()
    ref = torch.from_numpy(np.ascontiguousarray(clip.frame(6)[0]).reshape(-1)).cuda()
    net = full_pel_net(cur, ref, w, h)
    nb = w * h // 64
    out = [torch.zeros((nb, 2), dtype=torch.int16, device="cuda") for _ in range(2)]
    met = [torch.zeros(nb, dtype=torch.int32, device="cuda") for _ in range(2)]
    padded = torch.empty((w + 32) * (h + 32), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    L = eng.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

    def base():
        eng.luma_search_2step(cur, ref, net, out[0], met[0], w, h)

    def tma(pad_only=0):
        rc = L.vp8b200_experiment_search_2step_tma(st, P(cur), P(ref), P(padded), P(net), P(out[1]), P(met[1]), w, h, pad_only)
        assert rc == 0, rc

    def timed(fn):
        for _ in range(3):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            flush.fill_(1)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return 1000.0 * sum(a.elapsed_time(b) for a, b in ev) / reps

    t_base, t_tma, t_pad = timed(base), timed(tma), timed(lambda: tma(1))
    same = bool(torch.equal(out[0], out[1]) and torch.equal(met[0], met[1]))
    return {"size": [w, h], "loads_us": t_base, "tma_total_us": t_tma, "pad_plane_us": t_pad, "tma_search_only_us": t_tma - t_pad,
            "identical": same, "nonzero_vectors": int((out[0] != 0).any(dim=1).sum())}


if __name__ == "__main__":
    sizes = sys.argv[1:] or ["1920x1088", "3840x2160"]
    for s in sizes:
        w, h = map(int, s.split("x"))
        print(json.dumps(ab(w, h)), flush=True)
