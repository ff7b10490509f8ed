#!/usr/bin/env python3
"""Timeline of a multi-instance end-to-end run under MPS: when does every instance finish its first frame, its
warm-up and its last frame?  Shows how much of bench.py's e2e window is start-up stagger.  GPU only.

    python tools/e2e_stagger.py [P] [frames] [repeats] [GATE=0] [VP8B200_...=...]
"""
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
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 46
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    extra = dict(kv.split("=", 1) for kv in sys.argv[4:])
    sync = extra.pop("SYNC", "yield")
    W = 5
    tmp = tempfile.mkdtemp(prefix="stagger_", dir="/dev/shm")
    paths = []
    for p in range(min(P, 8)):
        y = os.path.join(tmp, "c%d.y4m" % p)
        gen_y4m.write_y4m(y, 1920, 1080, n, start=p * n)
        paths.append(y)
    for rep in range(reps):
        gate = os.path.join(tmp, "gate%d" % rep)
        os.makedirs(gate)
        if extra.get("GATE", "1") != "0":
            extra["VP8B200_START_GATE"] = "%s:%d:3" % (gate, P)
        with segments.MpsDaemon(os.path.join(tmp, "mps%d" % rep)) as d:
            T0 = time.perf_counter()
            procs = [segments.EncoderProcess(paths[p % 8], os.path.join(tmp, "o%d.ivf" % p), ENC_ARGS,
                                             os.path.join(tmp, "run%d" % p),
                                             env_extra=dict(d.env(), VP8B200_SYNC=sync, VP8B200_STATS=os.path.join(tmp, "st%d.json" % p),
                                                            **{k: v for k, v in extra.items() if k != "GATE"})) for p in range(P)]
            stamps = [pr.wait() for pr in procs]
        import json
        st = [json.load(open(os.path.join(tmp, "st%d.json" % p))) for p in range(P)]
        keys = [k for k in st[0] if k.startswith(("ms_", "cpu_"))]
        print("      per frame per instance: " + "  ".join("%s %.2f" % (k, sum(x[k] for x in st) / P / n) for k in keys), flush=True)
        first = sorted(s[0] - T0 for s in stamps)
        warm = sorted(s[W] - T0 for s in stamps)
        last = sorted(s[-1] - T0 for s in stamps)
        t0, t1 = max(s[W] for s in stamps), max(s[-1] for s in stamps)
        cnt = sum(1 for s in stamps for x in s if x > t0)
        u0, u1 = t0, min(s[-1] for s in stamps)
        cnt2 = sum(1 for s in stamps for x in s if u0 < x <= u1)
        print("rep %d mps=%s: first frame %.2f..%.2f s, warm %.2f..%.2f, last %.2f..%.2f | bench formula %.0f fps (%d frames) | "
              "all-active window %.0f fps (%d frames in %.2f s) | whole run %.0f fps" %
              (rep, d.active or bool(d.env()), first[0], first[-1], warm[0], warm[-1], last[0], last[-1], cnt / (t1 - t0), cnt,
               cnt2 / max(u1 - u0, 1e-9), cnt2, u1 - u0, P * n / (last[-1])), flush=True)
