#!/usr/bin/env python3
"""repro loop for an intermittent failure of the 2160p shim encode: N runs of the unmodified host on the CUDA shim
(reference profile) over the same 20-frame clip; prints exit codes and md5s, and the output tail of any odd run."""
import hashlib
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen_y4m  # noqa: E402
import _trace  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
w, h, frames = (int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "3840x2160x20").split("x"))
extra = dict(kv.split("=", 1) for kv in sys.argv[3:])
args = ["-qmin", 24, "-qmax", 24, "-g", 150, "-altref-range", 5, "-partitions", 8, "-threads", 12]
tmp = tempfile.mkdtemp(prefix="flaky_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
y4m = os.path.join(tmp, "clip.y4m")
gen_y4m.write_y4m(y4m, w, h, frames)
shim = os.path.join(ROOT, "vp8oclenc_b200", "lib")
seen = {}
for i in range(N):
    out = os.path.join(tmp, "o%d.ivf" % i)
    env = {"VP8B200_HOST_PROFILE": "reference"}
    env.update(extra)
    try:
        _trace.run_host(shim, os.path.join(tmp, "run%d" % i), y4m, out, args, env_extra=env)
        md5 = hashlib.md5(open(out, "rb").read()).hexdigest()
    except RuntimeError as err:
        md5 = "FAILED"
        print("run %d failed: %s" % (i, str(err)[-1500:]))
    seen[md5] = seen.get(md5, 0) + 1
    if os.path.exists(out):
        os.remove(out)
print(seen)
