"""Byte-identical .ivf at the sizes and lengths BASELINE.json's configurations name: the
UNMODIFIED reference host on the CUDA shim -- in the mode bench.py runs it (VP8B200_HOST_PROFILE=reference: lazy
downloads, page-locked source planes, polling waits, all entropy work on the GPU) -- against the same host on the
reference's own kernels compiled for the CPU (oracle/_ref), all host cores.

  config 2   1920x1080, 300 frames, -g 150: the forced key frame at 150, sixty golden/altref rotations, long-run
             drift (a single wrong pixel anywhere would propagate into every later frame of its GOP)
  config 3   3840x2160, 20 frames
  config 5   7680x4320, -g 15, 17 frames (two key frames)

The CPU side is the slow one (about 5 frames/s at 1080p on 16 cores, 0.3 at 4320p): the three cases take a few
minutes together and are marked `slow` as well as `gpu`.
"""
import hashlib
import json
import os
import sys

import pytest

import _trace
from _libs import ROOT

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.slow,
              pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not _trace.have_host(), reason="reference host binary not built")]

SHIM_DIR = os.path.join(ROOT, "vp8oclenc_b200", "lib")
sys.path.insert(0, os.path.join(ROOT, "tools"))

CASES = {
    "config2_1080p_300_frames_g150": (1920, 1080, 300, ["-qmin", 24, "-qmax", 24, "-g", 150, "-altref-range", 5, "-partitions", 8, "-threads", 12]),
    "config3_2160p_20_frames": (3840, 2160, 20, ["-qmin", 24, "-qmax", 24, "-g", 150, "-altref-range", 5, "-partitions", 8, "-threads", 12]),
    "config5_4320p_g15_17_frames": (7680, 4320, 17, ["-qmin", 24, "-qmax", 24, "-g", 15, "-altref-range", 5, "-partitions", 8, "-threads", 12]),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_baseline_configuration_byte_identical(case, tmp_path):
    import gen_y4m
    from vp8oclenc_b200 import segments
    w, h, frames, args = CASES[case]
    d = str(tmp_path)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else d
    y4m = os.path.join(shm, "vp8b200_%s_%d.y4m" % (case, os.getpid()))
    try:
        gen_y4m.write_y4m(y4m, w, h, frames)
        stats = os.path.join(d, "stats.json")
        for side, lib_dir, out, env in (
                ("CUDA shim", SHIM_DIR, "b200.ivf", {"VP8B200_HOST_PROFILE": "reference", "VP8B200_STATS": stats}),
                ("reference kernels on the CPU", _trace.REF_DIR, "ref.ivf", {"OMP_NUM_THREADS": str(os.cpu_count() or 1)})):
            try:
                _trace.run_host(lib_dir, os.path.join(d, out + ".run"), y4m, os.path.join(d, out), args, env_extra=env)
            except RuntimeError as err:
                pytest.fail("%s: the host program on the %s failed: %s" % (case, side, err))
    finally:
        if os.path.exists(y4m):
            os.remove(y4m)
    a = open(os.path.join(d, "ref.ivf"), "rb").read()
    b = open(os.path.join(d, "b200.ivf"), "rb").read()
    assert len(a) > 32 + 12 * frames
    if a != b:
        _, fa = segments.read_ivf(os.path.join(d, "ref.ivf"))
        _, fb = segments.read_ivf(os.path.join(d, "b200.ivf"))
        bad = [i for i, (x, y) in enumerate(zip(fa, fb)) if x[1] != y[1]]
        pytest.fail("%s: .ivf differs; frames %d / %d, first differing frames %s" % (case, len(fa), len(fb), bad[:8]))
    _, fr = segments.read_ivf(os.path.join(d, "b200.ivf"))
    keys = [i for i, (_, p) in enumerate(fr) if p[0] & 1 == 0]
    gop = args[args.index("-g") + 1]
    assert keys == list(range(0, frames, gop)), keys  # the clip forces no key frame: only the GOP counter does
    st = json.load(open(stats))
    assert st["entropy_host_fallbacks"] == 0          # the entropy stage never left the GPU
    assert st["host_kernels"] == frames               # only num_div_denom (1056 divisions) runs on the host, once per frame
    print("%s: %d frames, %d bytes, md5 %s, %d key frames" % (case, frames, len(b), hashlib.md5(b).hexdigest(), len(keys)))
