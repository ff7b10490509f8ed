#!/usr/bin/env python3
"""Where does the wall time of ONE unmodified-host encoder instance go?  Encodes a synthetic clip through the
CUDA OpenCL shim with VP8B200_STATS on and prints the per-entry-point split.  GPU only.

    python tools/e2e_profile.py [WxH] [frames] [perf]
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_y4m  # noqa: E402
from vp8oclenc_b200 import segments  # noqa: E402

ENC_ARGS = ["-qmin", "24", "-qmax", "24", "-g", "150", "-altref-range", "5", "-partitions", "8", "-threads", "12"]

if __name__ == "__main__":
    w, h = map(int, (sys.argv[1] if len(sys.argv) > 1 else "1920x1080").split("x"))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    with tempfile.TemporaryDirectory() as tmp:
        y4m = os.path.join(tmp, "clip.y4m")
        gen_y4m.write_y4m(y4m, w, h, n)
        stats = os.path.join(tmp, "stats.json")
        t0 = time.perf_counter()
        pr = segments.EncoderProcess(y4m, os.path.join(tmp, "clip.ivf"), ENC_ARGS, os.path.join(tmp, "run"),
                                     env_extra={"VP8B200_STATS": stats})
        st = pr.wait()
        t1 = time.perf_counter()
        s = json.load(open(stats))
        steady = (st[-1] - st[4]) / (len(st) - 5) * 1000.0
        print("frames %d  wall %.1f ms/frame overall, %.2f ms/frame steady (%.1f fps)" % (n, (t1 - t0) * 1000 / n, steady, 1000.0 / steady))
        for k, v in s.items():
            print("  %-16s %12.3f per frame" % (k, v / n))
        inside = sum(v for k, v in s.items() if k.startswith("ms_") and k != "ms_total")
        print("  host program outside the OpenCL calls: %.3f ms/frame" % ((s["ms_total"] - inside) / n))
