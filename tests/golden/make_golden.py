#!/usr/bin/env python3
"""Makes tests/golden/reference_kernels.json: the sha256 of every output of the REFERENCE's own kernels (oracle/_ref =
/root/reference/src/{GPU,CPU}_kernels.cl compiled for the CPU) on the seeded inputs of tests/test_oracle_vs_ref.py,
and tests/golden/reference_intra.json: the same for the reference's own intra path (src/intra_part.h compiled in
place, oracle/_ref/libref_intra.so) on the inputs of tests/test_intra_oracle.py.
Needs oracle/_ref (i.e. /root/reference); run it from the repo root:

    python tests/golden/make_golden.py

tests/test_oracle_vs_golden.py then checks the oracle against these vectors without the reference."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
out = os.path.join(HERE, "reference_kernels.json")
env = dict(os.environ, VP8_GOLDEN="record:" + out)
env.pop("VP8_NO_REF", None)
rc = subprocess.call([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_oracle_vs_ref.py"), "-q", "-x", "-p", "no:cacheprovider"],
                     cwd=ROOT, env=env)
print("wrote", out if rc == 0 else "NOTHING USABLE (tests failed)")
out2 = os.path.join(HERE, "reference_intra.json")
env = dict(os.environ, VP8_GOLDEN_INTRA="record:" + out2)
env.pop("VP8_NO_REF", None)
rc2 = subprocess.call([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_intra_oracle.py"), "-q", "-x", "-p", "no:cacheprovider"],
                      cwd=ROOT, env=env)
print("wrote", out2 if rc2 == 0 else "NOTHING USABLE (tests failed)")

# whole encodes: md5 of the .ivf the reference encoder (its host + its kernels on the CPU) writes for the cases of
# tests/test_gpu_e2e.py and the first frames of the 1080p / 2160p BASELINE configurations
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import hashlib  # noqa: E402
import json  # noqa: E402
import tempfile  # noqa: E402
import gen_y4m  # noqa: E402
import _trace  # noqa: E402
from golden_cases import IVF_CASES  # noqa: E402
table = {}
tmp = tempfile.mkdtemp(prefix="golden_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
for name, (w, h, frames, args) in IVF_CASES.items():
    y4m = os.path.join(tmp, name + ".y4m")
    gen_y4m.write_y4m(y4m, w, h, frames)
    ivf = os.path.join(tmp, name + ".ivf")
    _trace.run_host(_trace.REF_DIR, os.path.join(tmp, name + ".run"), y4m, ivf, args, env_extra={"OMP_NUM_THREADS": str(os.cpu_count() or 1)})
    b = open(ivf, "rb").read()
    table[name] = {"md5": hashlib.md5(b).hexdigest(), "bytes": len(b), "frames": frames}
    os.remove(y4m)
    print(name, table[name])
out3 = os.path.join(HERE, "reference_ivf.json")
json.dump(table, open(out3, "w"), indent=1, sort_keys=True)
print("wrote", out3)
sys.exit(rc or rc2)
