"""Whole-frame pin: oracle's vp8o_inter_frame / vp8o_loop_filter_planes against what the
UNMODIFIED reference host + the reference's own kernels (oracle/_ref) actually moved over the
OpenCL boundary while encoding a synthetic clip.

The trace gives, per frame, every buffer the host uploaded (current frame, previous filtered
reconstruction, segment data) and everything it read back (vectors, parts, reference ids,
coefficients, segment ids, SSIM, reconstruction).  We replay the uploads through the oracle and
demand identical read-backs; the next frame's upload of the previous reconstruction checks the
loop filter.  This validates the oracle's enqueue order, buffer ping-pong and reference
rotation against the real host, not against our reading of it.
"""
import ctypes
import os

import numpy as np
import pytest

import _trace
from _libs import P, oracle, ref

pytestmark = pytest.mark.skipif(ref() is None or not _trace.have_host(),
                                reason="oracle/_ref not built (needs /root/reference)")

W, H, FRAMES, GOP, ALTREF = 176, 144, 14, 12, 4


@pytest.fixture(scope="module")
def encode(tmp_path_factory):
    import sys
    sys.path.insert(0, os.path.join(_trace.ROOT, "tools"))
    import gen_y4m
    d = str(tmp_path_factory.mktemp("hostrun"))
    y4m = os.path.join(d, "clip.y4m")
    gen_y4m.write_y4m(y4m, W, H, FRAMES)
    out = {}
    for tag, extra in (("plain", []), ("ssim", ["-SSIM-target", "95"])):
        trace = os.path.join(d, tag + ".trace")
        _trace.run_host(_trace.REF_DIR, d, y4m, os.path.join(d, tag + ".ivf"),
                        ["-qmin", 20, "-qmax", 44, "-g", GOP, "-altref-range", ALTREF, "-partitions", 2, "-threads", 2] + extra,
                        trace=trace)
        out[tag] = (trace, os.path.join(d, tag + ".ivf"))
    return out


@pytest.mark.parametrize("tag,target", [("plain", -1.0), ("ssim", 0.95)])
def test_replay_inter_frames_through_oracle(encode, tag, target):
    o = oracle()
    events = _trace.read_trace(encode[tag][0])
    frames, _ = _trace.split_frames(events)
    assert len(frames) == FRAMES
    M = (W // 16) * (H // 16)
    ctx = ctypes.c_void_p(o.vp8o_ctx_create(W, H))
    host = _trace.HostState(GOP, ALTREF)
    n_inter = 0
    refs_used = set()
    for fi, fr in enumerate(frames):
        st = host.next_frame()
        # a frame ends up as a key frame when intra_transform() ran: it uploads the intra
        # reconstruction to cpu_frame_* (src/intra_part.h:1122-1124).  That also happens AFTER an
        # inter pass the host decided to throw away (src/vp8enc.cpp:443-453); the device-side
        # effects of that pass persist, so it is replayed (and checked) like any other.
        became_key = "cpu_frame_Y" in fr["w"]
        has_inter_pass = "macroblock_coeffs_gpu" in fr["r"]
        if became_key:
            host.cur_key = host.cur_golden = host.cur_altref = 1
            host.until_key, host.until_altref = GOP, ALTREF
            host.golden_no = host.altref_no = st["n"]
        if not has_inter_pass:
            continue
        n_inter += 1
        cur = [_trace.arr(fr["w"]["current_frame_" + p][0], np.uint8) for p in "YUV"]
        rec = [_trace.arr(fr["w"]["reconstructed_frame_" + p][0], np.uint8) for p in "YUV"]
        sd = _trace.arr(fr["w"]["segments_data_gpu"][0], np.int32, (4, 11))
        coef = np.zeros(M * 400, np.int16)
        vec = np.zeros(M * 8, np.int16)
        parts = np.zeros(M, np.int32)
        refid = np.zeros(M, np.int32)
        seg = np.zeros(M, np.int32)
        ssim = np.zeros(M, np.float32)
        o.vp8o_inter_frame(ctx, P(cur[0]), P(cur[1]), P(cur[2]), P(rec[0]), P(rec[1]), P(rec[2]), P(sd),
                           ctypes.c_float(target), st["prev_golden"], st["prev_altref"], st["altref_differs"],
                           P(coef), P(vec), P(parts), P(refid), P(seg), P(ssim))
        r = fr["r"]
        assert np.array_equal(vec, _trace.arr(r["macroblock_vectors_gpu"][0], np.int16)), "vectors, frame %d" % fi
        assert np.array_equal(parts, _trace.arr(r["macroblock_parts_gpu"][0], np.int32)), "parts, frame %d" % fi
        assert np.array_equal(refid, _trace.arr(r["macroblock_reference_frame_gpu"][0], np.int32)), "refs, frame %d" % fi
        assert np.array_equal(seg, _trace.arr(r["macroblock_segment_id_gpu"][0], np.int32)), "segment ids, frame %d" % fi
        assert np.array_equal(coef, _trace.arr(r["macroblock_coeffs_gpu"][0], np.int16)), "coefficients, frame %d" % fi
        assert np.array_equal(ssim.view(np.uint32), _trace.arr(r["macroblock_SSIM_gpu"][0], np.uint32)), "SSIM, frame %d" % fi
        for k, p in enumerate("YUV"):
            assert np.array_equal(rec[k], _trace.arr(r["reconstructed_frame_" + p][0], np.uint8)), "recon %s, frame %d" % (p, fi)
        refs_used |= set(refid.tolist())

        # loop filter: inputs are what the host handed back at unmap time, the result is what it
        # uploads as LAST at the start of the next frame
        if not became_key and fi + 1 < len(frames) and "reconstructed_frame_Y" in frames[fi + 1]["w"]:
            planes = [_trace.arr(fr["unmap"]["cpu_frame_" + p][0], np.uint8) for p in "YUV"]
            mb = _trace.arr(fr["unmap"]["macroblock_coeffs_cpu"][0], np.int16)
            parts_h = _trace.arr(fr["w"]["macroblock_parts_cpu"][0], np.int32)
            seg_h = _trace.arr(fr["w"]["macroblock_segment_id_cpu"][0], np.int32)
            sd_lf = _trace.arr(fr["w"]["segments_data_cpu"][-1], np.int32, (4, 11))
            nz = np.zeros(M, np.int32)
            o.vp8o_loop_filter_planes(P(planes[0]), P(planes[1]), P(planes[2]), P(mb), P(parts_h), P(seg_h), P(sd_lf),
                                      P(nz), W, H)
            nxt = frames[fi + 1]["w"]
            for k, p in enumerate("YUV"):
                assert np.array_equal(planes[k], _trace.arr(nxt["reconstructed_frame_" + p][0], np.uint8)), \
                    "loop-filtered %s, frame %d" % (p, fi)
            # the non-zero counts are read back after the next frame boundary marker
            nz_host = frames[fi + 1]["r"].get("macroblock_non_zero_coeffs_cpu")
            if nz_host:
                assert np.array_equal(nz, _trace.arr(nz_host[0], np.int32))
    o.vp8o_ctx_destroy(ctx)
    assert n_inter >= 3
    if tag == "plain":
        assert n_inter >= FRAMES - 3
        assert refs_used >= {0, 1} or refs_used >= {0, 2}, refs_used  # golden/altref really were searched


def test_reference_ivf_decodes(encode):
    cv2 = pytest.importorskip("cv2")
    cap = cv2.VideoCapture(encode["plain"][1])
    n = 0
    while True:
        ok, _ = cap.read()
        if not ok:
            break
        n += 1
    assert n == FRAMES
