#!/usr/bin/env python3
"""debug aid: encodes the cif_ssim test clip through the shim under several env settings and reports the first
differing frame against the CPU reference"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, ROOT)
import _trace, gen_y4m
from vp8oclenc_b200 import segments
SHIM = os.path.join(ROOT, "vp8oclenc_b200", "lib")
w, h, frames, args = 352, 288, 12, ["-qmin", 10, "-qmax", 50, "-g", 30, "-altref-range", 3, "-partitions", 4, "-threads", 12, "-SSIM-target", "93"]
d = tempfile.mkdtemp()
y4m = os.path.join(d, "clip.y4m")
gen_y4m.write_y4m(y4m, w, h, frames)
print(_trace.run_host(_trace.REF_DIR, d, y4m, os.path.join(d, "ref.ivf"), args + ["-print-info"])[-900:])
_, ref = segments.read_ivf(os.path.join(d, "ref.ivf"))
envs = [{}, {"VP8B200_SYNC": "poll40"}, {"VP8B200_PIN_HOST": "1"}, {"VP8B200_SYNC": "poll40", "VP8B200_PIN_HOST": "1"},
        {"VP8B200_SYNC": "yield", "VP8B200_PIN_HOST": "1"}, {"VP8B200_SYNC": "poll40", "VP8B200_PIN_HOST": "1", "VP8B200_ELIDE": "track"},
        {"VP8B200_SYNC": "poll40", "VP8B200_PIN_HOST": "1", "VP8B200_ELIDE": "off"}]
for env in envs * 2:
    out = os.path.join(d, "o.ivf")
    try:
        _trace.run_host(SHIM, d, y4m, out, args, env_extra=env)
        _, fr = segments.read_ivf(out)
        bad = [i for i, (a, b) in enumerate(zip(ref, fr)) if a[1] != b[1]]
        print(env, "frames", len(fr), "differing", bad)
    except Exception as e:
        print(env, "FAILED", e)
