#!/usr/bin/env python3
"""where the start-up of one encoder instance goes (VP8B200_STARTUP marks of the shim), alone and with several
instances starting at once, with and without MPS.  GPU only.

    python tools/startup_probe.py [instances ...]
"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_y4m  # noqa: E402
from vp8oclenc_b200 import segments  # noqa: E402

ENC = ["-qmin", "24", "-qmax", "24", "-g", "150", "-altref-range", "5", "-partitions", "8", "-threads", "12"]


def run(n, mps_env, tmp, tag, extra=None):
    y4m = os.path.join(tmp, "clip.y4m")
    if not os.path.exists(y4m):
        gen_y4m.write_y4m(y4m, 1920, 1080, 4)
    env = {"VP8B200_STARTUP": "1"}
    env.update(mps_env or {})
    env.update(extra or {})
    t0 = time.perf_counter()
    procs = [segments.EncoderProcess(y4m, os.path.join(tmp, "%s_%d.ivf" % (tag, i)), ENC, os.path.join(tmp, "%s_run%d" % (tag, i)),
                                     device=0, env_extra=env) for i in range(n)]
    for p in procs:
        p.wait(timeout=600)
    wall = time.perf_counter() - t0
    print("== %s: %d instance(s), wall %.2f s" % (tag, n, wall))
    for ln in procs[-1].output:  # (stderr is merged into the captured output)
        if "startup" in ln:
            print("   ", ln.strip())
    print("    python side: popen -> key frame coded %.2f s, inter frames %.2f s, -> exit %.2f s" %
          (procs[-1].stamps[0] - procs[-1].t_start, procs[-1].stamps[-1] - procs[-1].stamps[0], procs[-1].t_end - procs[-1].stamps[-1]))


def main():
    counts = [int(a) for a in sys.argv[1:]] or [1, 4]
    tmp = tempfile.mkdtemp(prefix="vp8startup_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    run(1, None, tmp, "warm-up (page cache, first driver load)")
    for n in counts:
        run(n, None, tmp, "no MPS")
    for n in counts:
        run(n, None, tmp, "no MPS, eager module loading", {"CUDA_MODULE_LOADING": "EAGER"})
    with segments.MpsDaemon(os.path.join(tmp, "mps")) as d:
        for n in counts:
            run(n, d.env(), tmp, "MPS")


if __name__ == "__main__":
    main()
