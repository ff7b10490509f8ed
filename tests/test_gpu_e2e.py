"""End-to-end drop-in test: the UNMODIFIED reference host, once with the reference's own kernels
on the CPU (oracle/_ref/libOpenCL.so.1) and once with our CUDA shim
(vp8oclenc_b200/lib/libOpenCL.so.1), on the same synthetic clip and the same -h options.
The two .ivf files must be byte-identical, and so must everything the host read back over the
OpenCL boundary (vectors, parts, reference ids, coefficients, segment ids, reconstruction;
the float SSIM within 1e-5 as the north star allows, in practice it is bit-identical too).
"""
import os

import numpy as np
import pytest

import _trace
from _libs import ROOT

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not _trace.have_host(), reason="reference host binary not built")]

SHIM_DIR = os.path.join(ROOT, "vp8oclenc_b200", "lib")

CASES = {
    # name: (w, h, frames, host args)
    "qcif": (176, 144, 14, ["-qmin", 20, "-qmax", 44, "-g", 12, "-altref-range", 4, "-partitions", 2, "-threads", 2]),
    "cif": (352, 288, 20, ["-qmin", 24, "-qmax", 24, "-g", 60, "-altref-range", 5, "-partitions", 1, "-threads", 2]),
    "cif_ssim": (352, 288, 12, ["-qmin", 10, "-qmax", 50, "-g", 30, "-altref-range", 3, "-partitions", 4, "-threads", 12,
                                "-SSIM-target", "93"]),
    "odd": (200, 120, 8, ["-qmin", 30, "-qmax", 30, "-g", 8, "-altref-range", 2, "-partitions", 8, "-threads", 12]),
}


@pytest.mark.parametrize("case", ["qcif", "cif", "cif_ssim", "odd", "cif_60_baseline", "1080p_12", "2160p_6"])
def test_ivf_matches_the_committed_reference_md5(case, tmp_path):
    """the shim's output against GOLDEN md5s of the reference encoder's files (tests/golden/reference_ivf.json, made on
    the CPU by tests/golden/make_golden.py): no reference run at test time"""
    import hashlib
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    from golden_cases import IVF_CASES
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_ivf.json")))[case]
    w, h, frames, args = IVF_CASES[case]
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, w, h, frames)
    _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, "b200.ivf"), args, env_extra={"VP8B200_HOST_PROFILE": "reference"})
    b = open(os.path.join(d, "b200.ivf"), "rb").read()
    assert (len(b), hashlib.md5(b).hexdigest()) == (want["bytes"], want["md5"]), case


def first_divergence(tr_a, tr_b):
    fa, _ = _trace.split_frames(_trace.read_trace(tr_a))
    fb, _ = _trace.split_frames(_trace.read_trace(tr_b))
    for i, (a, b) in enumerate(zip(fa, fb)):
        for (ka, na), (kb, nb) in zip(a["order"], b["order"]):
            if (ka, na) != (kb, nb):
                return "frame %d: call sequence differs: %s %s vs %s %s" % (i, ka, na, kb, nb)
        for kind in ("r",):
            for name in a[kind]:
                for j, (pa, pb) in enumerate(zip(a[kind][name], b[kind].get(name, []))):
                    if pa != pb:
                        if name == "macroblock_SSIM_gpu":
                            xa, xb = np.frombuffer(pa, np.float32), np.frombuffer(pb, np.float32)
                            if np.allclose(xa, xb, atol=1e-5, rtol=0):
                                continue
                        x, y = np.frombuffer(pa, np.uint8), np.frombuffer(pb, np.uint8)
                        k = int(np.flatnonzero(x != y)[0]) if x.size == y.size else -1
                        return "frame %d: read #%d of %s differs at byte %d of %d" % (i, j, name, k, x.size)
    return None


@pytest.mark.parametrize("case", list(CASES))
def test_ivf_and_readbacks_identical(case, tmp_path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    w, h, frames, args = CASES[case]
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, w, h, frames)
    _trace.run_host(_trace.REF_DIR, d, y4m, os.path.join(d, "ref.ivf"), args, trace=os.path.join(d, "ref.trace"))
    _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, "b200.ivf"), args, trace=os.path.join(d, "b200.trace"))
    div = first_divergence(os.path.join(d, "ref.trace"), os.path.join(d, "b200.trace"))
    assert div is None, div
    a = open(os.path.join(d, "ref.ivf"), "rb").read()
    b = open(os.path.join(d, "b200.ivf"), "rb").read()
    assert len(a) > 32 + 12 * frames
    assert a == b, "the .ivf bitstreams differ (sizes %d / %d)" % (len(a), len(b))


def test_without_trace_same_bitstream(tmp_path):
    """the trace forces synchronisation; the asynchronous path must give the same bytes"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    w, h, frames, args = CASES["cif"]
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, w, h, frames)
    _trace.run_host(_trace.REF_DIR, d, y4m, os.path.join(d, "ref.ivf"), args)
    _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, "b200.ivf"), args)
    assert open(os.path.join(d, "ref.ivf"), "rb").read() == open(os.path.join(d, "b200.ivf"), "rb").read()


@pytest.mark.parametrize("env", [{"VP8B200_TOKEN_CAP": "64"}, {"VP8B200_GPU_TOKENS": "0"}, {"VP8B200_ELIDE": "off"},
                                 {"VP8B200_ELIDE": "track"}, {"VP8B200_ELIDE": "assume"},
                                 {"VP8B200_ELIDE": "lazy", "VP8B200_GPU_TOKENS": "0"},
                                 {"VP8B200_FUSED": "0"}, {"VP8B200_SYNC": "sleep20"}, {"VP8B200_ENTROPY_THREADS": "3", "VP8B200_GPU_BOOLCODER": "0"},
                                 {"VP8B200_GPU_BOOLCODER": "0"}, {"VP8B200_SYNC": "poll40", "VP8B200_PIN_HOST": "1"},
                                 {"VP8B200_SYNC": "poll5", "VP8B200_SYNC_SPIN_US": "0", "VP8B200_PIN_HOST": "1", "VP8B200_ELIDE": "track"},
                                 {"VP8B200_SYNC": "yield", "VP8B200_PIN_HOST": "1"}, {"VP8B200_HOST_PROFILE": "reference"},
                                 {"VP8B200_ELIDE": "lazy"}],
                         ids=["token-scratch-grows", "host-entropy", "no-elision", "eager-downloads", "elision-assumed",
                              "lazy-host-entropy", "kernel-per-kernel", "sleep-sync", "three-entropy-threads", "host-bool-coder",
                              "poll-sync-pinned-host", "poll-sync-no-spin-eager", "yield-pinned-host", "reference-host-profile",
                              "lazy-downloads"])
def test_shim_modes_same_bitstream(env, tmp_path):
    """every switch of the shim changes HOW the bytes are produced, never the bytes: decision streams that
    outgrow their scratch, the host-only entropy path, no transfer elision, no fused launches, polling sync.
    The clip and options force key frames by SSIM fallback, so the host's intra path (which stores into the
    write-protected mirrors) is exercised as well."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    w, h, frames, args = CASES["cif_ssim"]
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, w, h, frames)
    _trace.run_host(_trace.REF_DIR, d, y4m, os.path.join(d, "ref.ivf"), args)
    _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, "b200.ivf"), args, env_extra=env)
    assert open(os.path.join(d, "ref.ivf"), "rb").read() == open(os.path.join(d, "b200.ivf"), "rb").read()


def test_cif_60_frames_baseline_configuration_byte_identical(tmp_path):
    """BASELINE configs[0]: 352x288 CIF, 60 frames, fixed q, against the reference's own kernels on the CPU"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    args = ["-qmin", 24, "-qmax", 24, "-g", 60, "-altref-range", 5, "-partitions", 4, "-threads", 4]
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, 352, 288, 60)
    _trace.run_host(_trace.REF_DIR, d, y4m, os.path.join(d, "ref.ivf"), args)
    _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, "b200.ivf"), args)
    a = open(os.path.join(d, "ref.ivf"), "rb").read()
    assert len(a) > 32 + 12 * 60
    assert a == open(os.path.join(d, "b200.ivf"), "rb").read()


def test_1080p_bench_configuration_byte_identical(tmp_path):
    """BASELINE configs[1] at full size (1920x1080 padded to 1088 lines, LAST+GOLDEN+ALTREF, 8 partitions, the
    options bench.py uses): a key frame and seven inter frames through the shim give the reference's bytes.
    The golden/altref rotation at altref-range 5 puts all three references in play."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    args = ["-qmin", 24, "-qmax", 24, "-g", 150, "-altref-range", 5, "-partitions", 8, "-threads", 12]
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, 1920, 1080, 8)
    _trace.run_host(_trace.REF_DIR, d, y4m, os.path.join(d, "ref.ivf"), args)
    _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, "b200.ivf"), args)
    a = open(os.path.join(d, "ref.ivf"), "rb").read()
    assert len(a) > 32 + 12 * 8
    assert a == open(os.path.join(d, "b200.ivf"), "rb").read()


def test_start_gate_two_instances(tmp_path):
    """VP8B200_START_GATE: two concurrent instances wait for each other at their second inter frame and then
    produce the bytes of an ungated run; a gate nobody else arrives at only delays (here: is not armed, the clip
    ends before its frame)"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    from vp8oclenc_b200 import segments
    w, h, frames, args = CASES["qcif"]
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, w, h, frames)
    _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, "plain.ivf"), args)
    gate = os.path.join(d, "gate")
    os.makedirs(gate)
    procs = [segments.EncoderProcess(y4m, os.path.join(d, "g%d.ivf" % i), args, os.path.join(d, "run%d" % i),
                                     env_extra={"VP8B200_START_GATE": "%s:2:2" % gate}) for i in range(2)]
    for pr in procs:
        pr.wait(timeout=300)
    assert len(os.listdir(gate)) == 2
    plain = open(os.path.join(d, "plain.ivf"), "rb").read()
    for i in range(2):
        assert open(os.path.join(d, "g%d.ivf" % i), "rb").read() == plain
    # armed for a frame the clip never reaches: no effect
    _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, "late.ivf"), args, env_extra={"VP8B200_START_GATE": "%s:5:1000" % gate})
    assert open(os.path.join(d, "late.ivf"), "rb").read() == plain


def test_lazy_downloads_skip_what_the_host_never_reads(tmp_path):
    """default mode: the coefficient / reconstruction / loop-filtered-frame downloads of inter frames are parked on
    the device and never cross the bus (the host only forwards those pointers); same bitstream as with eager
    downloads, far fewer device-to-host bytes"""
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    w, h, frames, args = CASES["cif"]
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, w, h, frames)
    stats = {}
    for mode in ("track", "lazy"):
        _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, mode + ".ivf"), args,
                        env_extra={"VP8B200_ELIDE": mode, "VP8B200_STATS": os.path.join(d, mode + ".json")})
        stats[mode] = json.load(open(os.path.join(d, mode + ".json")))
    assert open(os.path.join(d, "track.ivf"), "rb").read() == open(os.path.join(d, "lazy.ivf"), "rb").read()
    assert stats["lazy"]["d2h_bytes"] < 0.4 * stats["track"]["d2h_bytes"], stats


def _write_scene_cut_clip(path, w, h, frames, cut):
    """a clip whose chroma jumps at frame `cut` (the reference's colour-difference scene detector fires,
    src/vp8enc.cpp:265-311) and whose luma is a different texture from there on"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    clip = gen_y4m.Clip(w, h)
    with open(path, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420\n" % (w, h))
        for i in range(frames):
            y, u, v = clip.frame(i if i < cut else i + 400)
            if i >= cut:
                y = np.ascontiguousarray(y[::-1, ::-1])
                u = (255 - u.astype(np.int32)).clip(0, 255).astype(np.uint8)
                v = ((v.astype(np.int32) + 90) % 256).astype(np.uint8)
            f.write(b"FRAME\n")
            f.write(np.ascontiguousarray(y).tobytes() + np.ascontiguousarray(u).tobytes() + np.ascontiguousarray(v).tobytes())


@pytest.mark.parametrize("name,args,cut", [
    ("scene-cut", ["-qmin", 20, "-qmax", 40, "-g", 50, "-altref-range", 4, "-partitions", 4, "-threads", 4], 7),
    ("golden-altref-rotation", ["-qmin", 28, "-qmax", 28, "-g", 40, "-altref-range", 2, "-partitions", 2, "-threads", 4], None),
    ("gop-2", ["-qmin", 30, "-qmax", 30, "-g", 2, "-partitions", 2, "-threads", 2], None),
])
def test_host_control_flow_variants(name, args, cut, tmp_path):
    """paths of the unmodified host that the synthetic clips above do not reach: a key frame forced by the
    colour-difference scene detector in the middle of a GOP, many golden/altref refreshes, and -g 2 (every
    other frame a key frame; -g 1 crashes the reference host on its own runtime too, so it is not a case)"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    w, h, frames = 352, 288, 26 if cut is None else 14
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    if cut is None:
        gen_y4m.write_y4m(y4m, w, h, frames)
    else:
        _write_scene_cut_clip(y4m, w, h, frames, cut)
    out_ref = _trace.run_host(_trace.REF_DIR, d, y4m, os.path.join(d, "ref.ivf"), args)
    _trace.run_host(SHIM_DIR, d, y4m, os.path.join(d, "b200.ivf"), args)
    a = open(os.path.join(d, "ref.ivf"), "rb").read()
    assert len(a) > 32 + 12 * frames
    assert a == open(os.path.join(d, "b200.ivf"), "rb").read()
    if cut is not None:
        assert "1 scene changes detected by color change" in out_ref, out_ref[-400:]


def test_loop_filter_on_gpu_option_is_refused_cleanly(tmp_path):
    """-loop-filter-on-gpu (src/init.h:180-183, 285-293; src/loop_filter.h:57-138) cannot work against any runtime:
    the reference's own GPU_kernels.cl has the three kernels commented out and its host hands them a buffer it never
    creates (with oracle/_ref the host dies with SIGSEGV before the first frame).  The shim refuses the -DLOOP_FILTER
    build; the unmodified host then writes the build log to clErrors.txt and ends (src/init.h:188-201) without
    writing a frame.  Without the option the same clip encodes (every other test of this file)."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    d = str(tmp_path)
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, 176, 144, 4)
    for name, txt in (("GPU_kernels.cl", _trace.GPU_STUB), ("CPU_kernels.cl", _trace.CPU_STUB)):
        with open(os.path.join(d, name), "w") as f:
            f.write(txt)
    env = dict(os.environ, LD_LIBRARY_PATH=SHIM_DIR + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    p = subprocess.run([_trace.HOST_BIN, "-i", y4m, "-o", os.path.join(d, "out.ivf"), "-qmin", "24", "-qmax", "24", "-g", "12",
                        "-loop-filter-on-gpu"], cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert p.returncode not in (777 % 256, -11, 139), (p.returncode, p.stdout[-300:])  # neither "encoded" nor a crash
    assert b"kernel build fail" in p.stdout
    assert b"-loop-filter-on-gpu" in p.stderr
    log = open(os.path.join(d, "clErrors.txt")).read()
    assert "normal_loop_filter_MBH" in log and "Run without -loop-filter-on-gpu" in log
    assert not os.path.exists(os.path.join(d, "out.ivf")) or os.path.getsize(os.path.join(d, "out.ivf")) == 0
