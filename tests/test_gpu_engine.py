"""Frame-level parity: vp8b200_engine_* (the CUDA pipeline of one inter frame + loop filter)
against the oracle's vp8o_inter_frame / vp8o_loop_filter_planes over a sequence of frames with the
LAST / GOLDEN / ALTREF rotation of the reference host.  Bit-exact for every array, frame after
frame (errors would accumulate through the reference frames)."""
import ctypes
import os
import sys

import numpy as np
import pytest

import _trace
from _libs import P, ROOT, make_segment_data, oracle

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

sys.path.insert(0, os.path.join(ROOT, "tools"))


def run_sequence(w, h, nframes, altref_range, ssim_target, qi, host_api, fused=True, real_key=False):
    import gen_y4m
    from vp8oclenc_b200 import host as eng
    o = oracle()
    clip = gen_y4m.Clip(w, h)
    M = (w // 16) * (h // 16)
    os.environ["VP8B200_FUSED"] = "1" if fused else "0"  # read by vp8b200_engine_create
    e = eng.Engine(w, h)
    os.environ.pop("VP8B200_FUSED")
    ctx = ctypes.c_void_p(o.vp8o_ctx_create(w, h))
    state = _trace.HostState(10 ** 6, altref_range)
    sd = make_segment_data(qi)
    y0, u0, v0 = clip.frame(0)
    state.next_frame()  # frame 0 is the key frame
    if real_key:
        # coded for real: intra_transform() + loop filter with the key frame's segment data, oracle against engine
        from vp8oclenc_b200.hostlogic import intra_quants
        sdk = make_segment_data(qi, key=True)
        src = [np.ascontiguousarray(p).reshape(-1) for p in (y0, u0, v0)]
        rec = [np.zeros_like(p) for p in src]
        coef, modes = np.zeros(M * 400, np.int16), np.zeros(M * 16, np.int32)
        parts, seg, nz = np.zeros(M, np.int32), np.zeros(M, np.int32), np.zeros(M, np.int32)
        qk = np.asarray(intra_quants(sdk), np.int32)
        o.vp8o_intra_frame(w, h, P(src[0]), P(src[1]), P(src[2]), P(rec[0]), P(rec[1]), P(rec[2]), P(coef), P(modes), P(parts), P(seg), P(qk))
        o.vp8o_loop_filter_planes(P(rec[0]), P(rec[1]), P(rec[2]), P(coef), P(parts), P(seg), P(sdk), P(nz), w, h)
        e.key_frame(*[torch.from_numpy(p).cuda() for p in src], sdk)
        assert np.array_equal(e.read("intra_modes"), modes), "key frame: sub-block modes"
        assert np.array_equal(e.read("coeffs").reshape(M, 25, 16)[:, :24], coef.reshape(M, 25, 16)[:, :24]), "key frame: coefficients"
        e.loop_filter(sdk)
        assert np.array_equal(e.read("non_zero"), nz), "key frame: non-zero counts"
        for k, p in zip(("recon_y", "recon_u", "recon_v"), rec):
            assert np.array_equal(e.read(k), p), "key frame: loop-filtered %s" % k
    else:
        # its (pretend) reconstruction seeds LAST
        rec = [y0.copy().reshape(-1), u0.copy().reshape(-1), v0.copy().reshape(-1)]
        e.set_reconstruction(torch.from_numpy(rec[0]).cuda(), torch.from_numpy(rec[1]).cuda(), torch.from_numpy(rec[2]).cuda())
    refs_seen = set()
    for i in range(1, nframes):
        st = state.next_frame()
        y, u, v = [np.ascontiguousarray(p).reshape(-1) for p in clip.frame(i)]
        coef = np.zeros(M * 400, np.int16)
        vec = np.zeros(M * 8, np.int16)
        parts = np.zeros(M, np.int32)
        refid = np.zeros(M, np.int32)
        seg = np.zeros(M, np.int32)
        ssim = np.zeros(M, np.float32)
        nz = np.zeros(M, np.int32)
        o.vp8o_inter_frame(ctx, P(y), P(u), P(v), P(rec[0]), P(rec[1]), P(rec[2]), P(sd), ctypes.c_float(ssim_target),
                           st["prev_golden"], st["prev_altref"], st["altref_differs"], P(coef), P(vec), P(parts),
                           P(refid), P(seg), P(ssim))
        unfiltered = [r.copy() for r in rec]
        o.vp8o_loop_filter_planes(P(rec[0]), P(rec[1]), P(rec[2]), P(coef), P(parts), P(seg), P(sd), P(nz), w, h)
        if host_api:
            out = {k: torch.empty(n, dtype=dt).pin_memory() for k, (dt, n) in dict(
                coeffs=(torch.int16, M * 400), vectors=(torch.int16, M * 8), parts=(torch.int32, M),
                reference_frame=(torch.int32, M), segment_id=(torch.int32, M), ssim=(torch.float32, M),
                non_zero=(torch.int32, M), recon_y=(torch.uint8, w * h), recon_u=(torch.uint8, w * h // 4),
                recon_v=(torch.uint8, w * h // 4)).items()}
            e.encode_frame_host(torch.from_numpy(y).pin_memory(), torch.from_numpy(u).pin_memory(),
                                torch.from_numpy(v).pin_memory(), sd, ssim_target, st["prev_golden"], st["prev_altref"],
                                st["altref_differs"], out)
            got = {k: t.numpy() for k, t in out.items()}
        else:
            e.inter_frame(torch.from_numpy(y).cuda(), torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda(), sd,
                          ssim_target, st["prev_golden"], st["prev_altref"], st["altref_differs"])
            got = {k: e.read(k) for k in ("coeffs", "vectors", "parts", "reference_frame", "segment_id", "ssim")}
            for k, p in zip(("recon_y", "recon_u", "recon_v"), unfiltered):
                assert np.array_equal(e.read(k), p), "unfiltered %s, frame %d" % (k, i)
            e.loop_filter(sd)
            got.update({k: e.read(k) for k in ("non_zero", "recon_y", "recon_u", "recon_v")})
        assert np.array_equal(got["vectors"], vec), "vectors, frame %d" % i
        assert np.array_equal(got["parts"], parts), "parts, frame %d" % i
        assert np.array_equal(got["reference_frame"], refid), "reference ids, frame %d" % i
        assert np.array_equal(got["segment_id"], seg), "segment ids, frame %d" % i
        assert np.array_equal(got["coeffs"], coef), "coefficients, frame %d" % i
        assert np.array_equal(got["ssim"].view(np.uint32), ssim.view(np.uint32)), "SSIM, frame %d" % i
        assert np.array_equal(got["non_zero"], nz), "non-zero counts, frame %d" % i
        for k, p in zip(("recon_y", "recon_u", "recon_v"), rec):
            assert np.array_equal(got[k], p), "loop-filtered %s, frame %d" % (k, i)
        refs_seen |= set(refid.tolist())
    o.vp8o_ctx_destroy(ctx)
    e.close()
    return refs_seen


def test_engine_sequence_cif_three_references():
    refs = run_sequence(352, 288, 12, 4, -1.0, (24, 24, 24, 24), host_api=False)
    assert len(refs) >= 2, refs  # GOLDEN and/or ALTREF really were chosen for some macroblocks


def test_engine_sequence_ssim_ladder():
    run_sequence(352, 288, 6, 3, 0.93, (6, 20, 35, 50), host_api=False)


@pytest.mark.parametrize("target", [-1.0, 0.93])
def test_engine_kernel_per_kernel_sequence(target):
    """the unfused path (one launch per reference kernel, as the OpenCL shim replays it)"""
    run_sequence(352, 288, 5, 3, target, (6, 20, 35, 50), host_api=False, fused=False)


def test_engine_odd_macroblock_counts():
    """sizes whose macroblock count is not a multiple of the fused kernel's CTA size"""
    run_sequence(208, 176, 4, 3, 0.9, (10, 24, 40, 60), host_api=False)


def test_engine_key_frame_then_inter_frames():
    """a real key frame on the GPU (vp8b200_engine_key_frame: intra + loop filter with the key frame's filter table),
    then inter frames that reference it -- every array against the oracle"""
    refs = run_sequence(352, 288, 7, 3, -1.0, (24, 24, 24, 24), host_api=False, real_key=True)
    assert 0 in refs


def test_engine_key_frame_1080p():
    run_sequence(1920, 1088, 3, 5, -1.0, (24, 24, 24, 24), host_api=False, real_key=True)


def test_engine_host_buffers_api():
    run_sequence(176, 144, 8, 3, -1.0, (30, 30, 30, 30), host_api=True)


def test_engine_1080p_two_frames():
    run_sequence(1920, 1088, 3, 5, -1.0, (24, 24, 24, 24), host_api=True)


def test_engine_2160p_vs_oracle():
    """BASELINE config 3 (3840x2160): one inter frame + loop filter, bit-exact against the oracle"""
    run_sequence(3840, 2160, 2, 5, -1.0, (24, 24, 24, 24), host_api=False)


def test_engine_4320p_vs_oracle():
    """BASELINE config 4 (7680x4320), the largest size: one inter frame + loop filter, bit-exact against the oracle"""
    run_sequence(7680, 4320, 2, 5, -1.0, (24, 24, 24, 24), host_api=False)


def test_engine_4320p_fused_equals_kernel_per_kernel():
    """At the largest size, over several frames and with the SSIM ladder active: the two independent CUDA
    implementations of the frame (fused launches vs one launch per reference kernel) agree on every output,
    and a loop filter whose levels are all 0 is the identity (Q6)."""
    import gen_y4m
    from vp8oclenc_b200 import host as eng
    w, h, nframes = 7680, 4320, 3
    clip = gen_y4m.Clip(w, h)
    engines = []
    for fused in ("1", "0"):
        os.environ["VP8B200_FUSED"] = fused
        engines.append(eng.Engine(w, h))
    os.environ.pop("VP8B200_FUSED")
    state = _trace.HostState(10 ** 6, 5)
    sd = make_segment_data((6, 20, 35, 50))
    state.next_frame()
    key = [torch.from_numpy(np.ascontiguousarray(p).reshape(-1)).cuda() for p in clip.frame(0)]
    for e in engines:
        e.set_reconstruction(*key)
    for i in range(1, nframes):
        st = state.next_frame()
        cur = [torch.from_numpy(np.ascontiguousarray(p).reshape(-1)).cuda() for p in clip.frame(i)]
        outs = []
        for e in engines:
            e.inter_frame(cur[0], cur[1], cur[2], sd, 0.93, st["prev_golden"], st["prev_altref"], st["altref_differs"])
            e.loop_filter(sd)
            outs.append({k: e.read(k) for k in ("coeffs", "vectors", "parts", "reference_frame", "segment_id", "ssim",
                                                "non_zero", "recon_y", "recon_u", "recon_v")})
        for k in outs[0]:
            assert np.array_equal(outs[0][k].view(np.uint8), outs[1][k].view(np.uint8)), \
                "%s differs between the fused and the per-kernel path, frame %d" % (k, i)
        assert len(set(outs[0]["segment_id"].tolist())) > 1, "the SSIM ladder was not exercised"
    sd0 = make_segment_data((24, 24, 24, 24), lf_level=(0, 0, 0, 0))
    before = engines[0].read("recon_y").copy()
    engines[0].loop_filter(sd0)
    assert np.array_equal(engines[0].read("recon_y"), before), "loop filter with level 0 must be the identity"
    for e in engines:
        e.close()
