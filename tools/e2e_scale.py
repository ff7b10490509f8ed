#!/usr/bin/env python3
"""How does the unmodified-host path scale with encoder instances per GPU?  Runs P instances on cuda:0 for
several P, samples GPU utilisation meanwhile.  GPU only.

    python tools/e2e_scale.py [WxH] [frames] [P,P,...]
"""
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_y4m  # noqa: E402
from vp8oclenc_b200 import segments  # noqa: E402

ENC_ARGS = ["-qmin", "24", "-qmax", "24", "-g", "150", "-altref-range", "5", "-partitions", "8", "-threads", "12"]


class Util:
    def __init__(self):
        self.vals, self.stop_flag = [], False
        self.t = threading.Thread(target=self.run, daemon=True)
        self.t.start()

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "--query-gpu=utilization.gpu", "--format=csv,noheader,nounits", "-i", "0"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                self.vals.append(float(o.splitlines()[0]))
            except Exception:
                pass
            time.sleep(0.2)

    def stop(self):
        self.stop_flag = True
        self.t.join()
        v = sorted(self.vals)
        return v[len(v) // 2] if v else None


if __name__ == "__main__":
    w, h = map(int, (sys.argv[1] if len(sys.argv) > 1 else "1920x1080").split("x"))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    Ps = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2,4,8,12,16").split(",")]
    extra = dict(kv.split("=", 1) for kv in sys.argv[4:])
    with tempfile.TemporaryDirectory() as tmp:
        y4m = os.path.join(tmp, "clip.y4m")
        gen_y4m.write_y4m(y4m, w, h, n)
        for P in Ps:
            u = Util()
            procs = [segments.EncoderProcess(y4m, os.path.join(tmp, "o%d.ivf" % p), ENC_ARGS, os.path.join(tmp, "run%d" % p),
                                             env_extra=dict(extra, VP8B200_STATS=os.path.join(tmp, "stats%d.json" % p)))
                     for p in range(P)]
            stamps = [pr.wait() for pr in procs]
            util = u.stop()
            t0 = max(st[4] for st in stamps)
            t1 = max(st[-1] for st in stamps)
            count = sum(1 for st in stamps for x in st if x > t0)
            la = os.getloadavg()[0]
            import json
            st = [json.load(open(os.path.join(tmp, "stats%d.json" % p))) for p in range(P)]
            keys = [k for k in st[0] if k.startswith(("ms_", "cpu_"))]
            print("      per frame per instance: " + "  ".join("%s %.2f" % (k, sum(x[k] for x in st) / P / n) for k in keys), flush=True)
            print("P=%2d  %7.1f frames/s  (%.2f ms/frame aggregate)  gpu util median %s%%  loadavg %.1f" %
                  (P, count / (t1 - t0), 1000.0 * (t1 - t0) / count, util, la), flush=True)
