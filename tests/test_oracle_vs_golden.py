"""The oracle against COMMITTED golden vectors of the reference (tests/golden/reference_kernels.json: sha256 of every
output of the reference's own kernels on the seeded inputs of tests/test_oracle_vs_ref.py, made by
tests/golden/make_golden.py where /root/reference exists).  Runs without the reference: the cases of
test_oracle_vs_ref.py are executed with the reference switched off (VP8_NO_REF) and every oracle output has to hash to
the stored reference output."""
import json
import os
import subprocess
import sys

from _libs import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_kernels.json")


def test_oracle_reproduces_the_committed_reference_vectors():
    table = json.load(open(GOLDEN))
    assert len(table) >= 30 and sum(len(v) for v in table.values()) >= 100
    env = dict(os.environ, VP8_GOLDEN="check:" + GOLDEN, VP8_NO_REF="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_oracle_vs_ref.py"), "-q", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_oracle_intra_path_reproduces_the_committed_reference_vectors():
    golden = os.path.join(ROOT, "tests", "golden", "reference_intra.json")
    table = json.load(open(golden))
    assert len(table) >= 7 and sum(len(v) for v in table.values()) >= 49
    env = dict(os.environ, VP8_GOLDEN_INTRA="check:" + golden, VP8_NO_REF="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_intra_oracle.py"), "-q", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_a_wrong_oracle_output_is_caught(tmp_path):
    """the check is not vacuous: with one stored hash altered the run fails"""
    table = json.load(open(GOLDEN))
    key = sorted(k for k in table if k.startswith("test_downsample"))[0]
    table[key][0] = "0" * 64
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps(table))
    env = dict(os.environ, VP8_GOLDEN="check:" + str(bad), VP8_NO_REF="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_oracle_vs_ref.py"), "-q", "-p", "no:cacheprovider",
                        "-k", "test_downsample"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "failed" in r.stdout
