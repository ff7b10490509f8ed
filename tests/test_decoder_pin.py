"""Independent pin of the oracle (VERDICT r01 "missing" #6): the .ivf the reference
encoder writes is decoded by libvpx/FFmpeg (through OpenCV, code none of us wrote) and compared, frame by frame,
with the encoder's OWN loop-filtered reconstruction as it crossed the OpenCL boundary.

This is the one check that does not rest on oracle/clc_compat.hpp: if the emulated OpenCL C mis-computed a
convert_*_sat, an image clamp or a filter tap, the encoder's reconstruction (made by the emulated kernels) and
the decoder's (made by libvpx from the bitstream alone) would part within a frame and never meet again, because
every inter frame predicts from the previous reconstruction.

  luma    bit-exact: OpenCV hands out the decoder's raw luma plane with CAP_PROP_CONVERT_RGB=0
  chroma  through the colour conversion: the encoder's reconstruction is written as a Y4M file and read back with
          the same OpenCV/swscale path as the .ivf; the two BGR frames must be identical (a +-2 change of one
          chroma sample changes them, see test_bgr_comparison_sees_chroma)

Where the reference's reconstruction is NOT what a conforming decoder computes (the wrapping lines of
`construct`; Q7: unclamped chaining in the loop filter) it drifts; test_quirk_drift_on_saturating_content shows
that this needs saturating content and records where it starts.  The CUDA path reproduces the reference there
(bit-exact .ivf parity is the contract), drift included.
"""
import os
import sys

import numpy as np
import pytest

import _trace

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.skipif(not _trace.have_host(), reason="oracle/_ref not built (needs /root/reference)")

sys.path.insert(0, os.path.join(_trace.ROOT, "tools"))


def decode(path, raw_luma):
    cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
    if raw_luma:
        cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f.copy())
    return out


def encoder_reconstructions(trace, w, h):
    """{frame index: (Y, U, V)}: frame i's loop-filtered reconstruction is what the host uploads as LAST at the start
    of frame i+1 (src/vp8enc.cpp:385-387); a frame followed by a key frame has no such upload"""
    frames, _ = _trace.split_frames(_trace.read_trace(trace))
    rec = {}
    for i in range(len(frames) - 1):
        wr = frames[i + 1]["w"]
        if "reconstructed_frame_Y" in wr:
            rec[i] = (_trace.arr(wr["reconstructed_frame_Y"][0], np.uint8).reshape(h, w),
                      _trace.arr(wr["reconstructed_frame_U"][0], np.uint8).reshape(h // 2, w // 2),
                      _trace.arr(wr["reconstructed_frame_V"][0], np.uint8).reshape(h // 2, w // 2))
    return rec, len(frames)


def write_y4m_planes(path, w, h, planes_list):
    with open(path, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420\n" % (w, h))
        for y, u, v in planes_list:
            f.write(b"FRAME\n" + y.tobytes() + u.tobytes() + v.tobytes())


def quirk_events_per_frame(trace, w, h, gop, altref_range, target):
    """replays the encode through the oracle (as tests/test_oracle_frame_vs_host.py does) and returns, per frame,
    (number of blocks whose predictor Q5 changed, number of loop-filter edges that chained an unclamped value, Q7,
    whether the frame ended up as a key frame)"""
    import ctypes
    from _libs import P, oracle
    o = oracle()
    o.vp8o_quirk_log_get.restype = ctypes.c_int
    frames, _ = _trace.split_frames(_trace.read_trace(trace))
    M = (w // 16) * (h // 16)
    ctx = ctypes.c_void_p(o.vp8o_ctx_create(w, h))
    host = _trace.HostState(gop, altref_range)
    out = []
    scratch = np.zeros(3, np.int32)
    for fr in frames:
        st = host.next_frame()
        became_key = "cpu_frame_Y" in fr["w"]
        if became_key:
            host.cur_key = host.cur_golden = host.cur_altref = 1
            host.until_key, host.until_altref = gop, altref_range
            host.golden_no = host.altref_no = st["n"]
        o.vp8o_quirk_log_reset()
        if "macroblock_coeffs_gpu" in fr["r"]:  # an inter pass ran (its device-side effects persist even if thrown away)
            cur = [_trace.arr(fr["w"]["current_frame_" + p][0], np.uint8) for p in "YUV"]
            rec = [_trace.arr(fr["w"]["reconstructed_frame_" + p][0], np.uint8) for p in "YUV"]
            sd = _trace.arr(fr["w"]["segments_data_gpu"][0], np.int32, (4, 11))
            coef, vec = np.zeros(M * 400, np.int16), np.zeros(M * 8, np.int16)
            parts, refid, seg = np.zeros(M, np.int32), np.zeros(M, np.int32), np.zeros(M, np.int32)
            ssim = np.zeros(M, np.float32)
            o.vp8o_inter_frame(ctx, P(cur[0]), P(cur[1]), P(cur[2]), P(rec[0]), P(rec[1]), P(rec[2]), P(sd),
                               ctypes.c_float(target), st["prev_golden"], st["prev_altref"], st["altref_differs"],
                               P(coef), P(vec), P(parts), P(refid), P(seg), P(ssim))
        q5 = 0 if became_key else o.vp8o_quirk_log_get(0, P(scratch), 1)  # (a discarded inter pass leaves no trace in the stream)
        q7 = 0
        if "cpu_frame_Y" in fr["unmap"] and "segments_data_cpu" in fr["w"]:
            planes = [_trace.arr(fr["unmap"]["cpu_frame_" + p][0], np.uint8) for p in "YUV"]
            mb = _trace.arr(fr["unmap"]["macroblock_coeffs_cpu"][0], np.int16)
            nz = np.zeros(M, np.int32)
            parts_h = _trace.arr(fr["w"]["macroblock_parts_cpu"][0], np.int32)
            seg_h = _trace.arr(fr["w"]["macroblock_segment_id_cpu"][0], np.int32)
            sd_lf = _trace.arr(fr["w"]["segments_data_cpu"][-1], np.int32, (4, 11))
            o.vp8o_quirk_log_reset()
            o.vp8o_loop_filter_planes(P(planes[0]), P(planes[1]), P(planes[2]), P(mb), P(parts_h), P(seg_h), P(sd_lf),
                                      P(nz), w, h)
            q7 = o.vp8o_quirk_log_get(1, P(scratch), 1)
        out.append((q5, q7, became_key))
    o.vp8o_ctx_destroy(ctx)
    return out


def check_against_decoder(ivf, trace, w, h, tmp, want_frames, gop, altref_range, target=-1.0, replaced=None):
    """-> (frames compared, frames that had to match exactly, frames where the reference legitimately drifted)"""
    rec, n = encoder_reconstructions(trace, w, h)
    assert n == want_frames
    luma = decode(ivf, True)
    bgr = decode(ivf, False)
    assert len(luma) == want_frames and len(bgr) == want_frames
    idx = sorted(rec)
    y4m = os.path.join(tmp, "recon.y4m")
    write_y4m_planes(y4m, w, h, [rec[i] for i in idx])
    ours = decode(y4m, False)
    assert len(ours) == len(idx)
    quirks = quirk_events_per_frame(trace, w, h, gop, altref_range, target)
    # a frame may differ from the decoder's only if, since the last key frame, the reference's own non-conforming
    # arithmetic (Q5 / Q7) changed a pixel: references carry the difference forward until the next key frame
    tainted, clean_since_key = [], True
    for i in range(want_frames):
        q5, q7, key = quirks[i]
        if key:
            clean_since_key = True
        if q5 or q7 or (replaced or {}).get(i, 0):
            clean_since_key = False
        tainted.append(not clean_since_key)
    exact, drifted = [], []
    for k, i in enumerate(idx):
        assert luma[i].shape == (h, w)
        same = np.array_equal(luma[i], rec[i][0]) and np.array_equal(bgr[i], ours[k])
        if not tainted[i]:
            assert np.array_equal(luma[i], rec[i][0]), "decoded luma != encoder reconstruction, frame %d" % i
            assert np.array_equal(bgr[i], ours[k]), "decoded frame != encoder reconstruction (colour path), frame %d" % i
            exact.append(i)
        elif not same:
            drifted.append(i)
    return idx, exact, drifted, quirks


# (w, h, frames, gop, altref range, SSIM target or None, frames that must be compared exactly at least)
CASES = {
    # "soft": the clip's luma is compressed to 72..181, so the six-tap overshoot never leaves 0..255, Q5 cannot fire
    # and EVERY frame has to match the independent decoder -- the strong form of the pin
    "cif_soft_gop12": (352, 288, 30, 12, 4, None, 27),
    "vga_soft_q24_8parts": (640, 368, 16, 150, 5, None, 15),
    "cif_gop12": (352, 288, 24, 12, 4, None, 4),
    "qcif_ssim_ladder": (176, 144, 14, 12, 4, 95, 2),
    "vga_q24_8parts": (640, 368, 12, 150, 5, None, 4),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_libvpx_decodes_to_the_encoders_reconstruction(case, tmp_path, record_property):
    import gen_y4m
    w, h, n, gop, altref, ssim, min_exact = CASES[case]
    args = ["-qmin", 24 if ssim is None and gop > 100 else 20, "-qmax", 24 if ssim is None and gop > 100 else 44, "-g", gop,
            "-altref-range", altref, "-partitions", 8 if gop > 100 else 2, "-threads", 12 if gop > 100 else 2]
    if ssim is not None:
        args += ["-SSIM-target", ssim]
    d = str(tmp_path)
    y4m, ivf, trace = os.path.join(d, "clip.y4m"), os.path.join(d, "out.ivf"), os.path.join(d, "out.trace")
    if "soft" in case:
        clip = gen_y4m.Clip(w, h)
        write_y4m_planes(y4m, w, h, [((64 + f[0] // 2).astype(np.uint8), f[1], f[2]) for f in (clip.frame(i) for i in range(n))])
    else:
        gen_y4m.write_y4m(y4m, w, h, n)
    out = _trace.run_host(_trace.REF_DIR, os.path.join(d, "run"), y4m, ivf, args + ["-print-info"], trace=trace)
    # macroblocks the HOST re-coded as intra (SSIM ladder fallback, src/vp8enc.cpp:227-262) are reconstructed by its
    # own C code (src/intra_part.h:855-1087), outside the path and outside the oracle: such frames count as tainted
    import re
    replaced = {int(m.group(1)): int(m.group(2)) for m in re.finditer(r"(\d+)>AvgSSIM=[^\n]*?repl:(\d+)", out)}
    idx, exact, drifted, quirks = check_against_decoder(ivf, trace, w, h, d, n, gop, altref,
                                                        -1.0 if ssim is None else ssim / 100.0, replaced)
    first_q = next((i for i, q in enumerate(quirks) if q[0] or q[1]), None)
    record_property("frames_compared", len(idx))
    record_property("frames_exact", len(exact))
    record_property("first_quirk_frame", first_q)
    record_property("frames_drifted", drifted)
    print("%s: %d frames compared, %d bit-exact against libvpx, first Q5/Q7 event at frame %s %s, drifted frames %s" %
          (case, len(idx), len(exact), first_q, quirks[first_q][:2] if first_q is not None else "", drifted))
    assert len(idx) >= n - 3       # every frame that is followed by an inter frame was compared
    assert len(exact) >= min_exact  # and most of them had to match the independent decoder bit for bit
    # a frame only ever differs from the decoder's after a logged quirk event (asserted inside); conversely the
    # drift, once there, ends at the next key frame
    for i in drifted:
        assert any(q[0] or q[1] for q in quirks[:i + 1]) or any(replaced.get(k, 0) for k in range(i + 1))


def test_bgr_comparison_sees_chroma(tmp_path):
    """the colour-path comparison is sensitive to a small chroma error (it is what pins U and V)"""
    import gen_y4m
    w, h = 176, 144
    y, u, v = gen_y4m.Clip(w, h).frame(0)
    u2 = u.copy()
    u2[10:12, 20:22] += 2
    a, b = os.path.join(str(tmp_path), "a.y4m"), os.path.join(str(tmp_path), "b.y4m")
    write_y4m_planes(a, w, h, [(y, u, v)])
    write_y4m_planes(b, w, h, [(y, u2, v)])
    fa, fb = decode(a, False)[0], decode(b, False)[0]
    assert not np.array_equal(fa, fb)
    assert np.array_equal(decode(a, True)[0], y)  # and the raw path is the luma plane, untouched


def test_quirk_drift_on_saturating_content(tmp_path):
    """Q5/Q7: on 0/255 content moving in half-pel steps the reference's reconstruction leaves the decoder's at the
    first inter frame (the six-tap overshoot wraps instead of saturating); the key frame is still exact.  This is
    reference behaviour the product reproduces bit for bit -- the test documents where it shows, so that a reader
    of a libvpx mismatch on such content knows it is the reference's, not the port's."""
    w, h, n = 176, 144, 6
    rng = np.random.default_rng(5)
    big = np.kron(rng.integers(0, 2, size=(h + 40, w + 40)), np.ones((2, 2), int)) * 255
    frames = []
    for i in range(n):
        p = big[i:i + 2 * h, 3 * i:3 * i + 2 * w]
        y = ((p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + 2) // 4).astype(np.uint8)
        c = np.full((h // 2, w // 2), 128, np.uint8)
        frames.append((y, c, c))
    d = str(tmp_path)
    y4m, ivf, trace = os.path.join(d, "harsh.y4m"), os.path.join(d, "h.ivf"), os.path.join(d, "h.trace")
    write_y4m_planes(y4m, w, h, frames)
    _trace.run_host(_trace.REF_DIR, os.path.join(d, "run"), y4m, ivf,
                    ["-qmin", 24, "-qmax", 24, "-g", 100, "-altref-range", 4, "-partitions", 2, "-threads", 2], trace=trace)
    rec, _ = encoder_reconstructions(trace, w, h)
    luma = decode(ivf, True)
    assert np.array_equal(luma[0], rec[0][0]), "the key frame has no quirk path: it must decode exactly"
    differing = [i for i in sorted(rec) if not np.array_equal(luma[i], rec[i][0])]
    assert differing and differing[0] == 1, differing  # drift starts with the first inter frame on this content
