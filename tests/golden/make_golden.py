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
sys.exit(rc or rc2)
