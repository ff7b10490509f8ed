"""GPU parity tests proper: every CUDA kernel of vp8oclenc_b200, called through the C ABI of
include/vp8b200.h, against the oracle (oracle/vp8_oracle.c, itself pinned against the
reference's own kernels) on the same seeded inputs.  Bit-exact for everything, the float SSIM
included (the 1e-5 tolerance of the north star is only needed against a real OpenCL device).
"""
import ctypes

import numpy as np
import pytest

import os

from _libs import P, ROOT, make_segment_data, oracle

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def rng(seed):
    return np.random.default_rng(seed)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def textured(r, h, w, smooth=4, noise=12):
    a = r.integers(0, 256, size=(h // smooth + 2, w // smooth + 2)).astype(np.float64)
    a = np.kron(a, np.ones((smooth, smooth)))[:h, :w]
    a = a + r.integers(-noise, noise + 1, size=(h, w))
    return np.clip(a, 0, 255).astype(np.uint8)


@pytest.fixture(scope="module")
def eng():
    from vp8oclenc_b200 import host as h
    info = h.device_info()
    assert info["sm_count"] > 0
    return h


def test_library_is_native(eng):
    assert b"sm_100a" in eng.lib().vp8b200_version()


@pytest.mark.parametrize("w,h", [(32, 16), (352, 288), (120, 68), (44, 36), (1920, 1088)])
def test_downsample(eng, w, h):
    o = oracle()
    src = rng(2).integers(0, 256, size=(h, w)).astype(np.uint8)
    a = np.zeros((h // 2, w // 2), np.uint8)
    o.vp8o_downsample_x2(P(src), P(a), w, h)
    d = torch.zeros((h // 2, w // 2), dtype=torch.uint8, device="cuda")
    eng.downsample_x2(dev(src), d, w, h)
    assert np.array_equal(a, host(d))


def test_reset_vectors(eng):
    n = 4 * 396
    nets = [torch.full((n, 2), 5, dtype=torch.int16, device="cuda") for _ in range(6)]
    mets = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(3)]
    eng.reset_vectors(*nets, *mets)
    assert all(int(t.abs().sum()) == 0 for t in nets)
    assert all(bool((t == 0x7fffffff).all()) for t in mets)


@pytest.mark.parametrize("rate", [16, 8, 4, 2, 1])
@pytest.mark.parametrize("w,h", [(64, 48), (120, 68), (88, 72), (352, 288), (22, 18), (11, 9), (45, 27)])
def test_luma_search_1step(eng, rate, w, h):
    o = oracle()
    for seed, lim in ((10, 6), (11, 40), (12, 0)):
        r = rng(seed * 100 + rate + w)
        prev = textured(r, h, w)
        sx, sy = int(r.integers(-3, 4)), int(r.integers(-3, 4))
        cur = np.clip(np.roll(prev, (sy, sx), axis=(0, 1)).astype(np.int32) + r.integers(-6, 7, size=(h, w)), 0, 255).astype(np.uint8)
        net_width = max(2, (w * rate) // 8)  # >= blocks per row, like 2*mb_width in the encoder
        net_h = max(2, ((h * rate) // 16) * 2) + 2
        src_net = (r.integers(-lim, lim + 1, size=(net_h * net_width, 2)) * rate).astype(np.int16)
        dst_a = np.full((net_h * net_width, 2), 77, np.int16)
        o.vp8o_luma_search_1step(P(cur), P(prev), P(src_net), P(dst_a), net_width, w, h, rate)
        d = dev(np.full((net_h * net_width, 2), 77, np.int16))
        eng.luma_search_1step(dev(cur), dev(prev), dev(src_net), d, net_width, w, h, rate)
        assert np.array_equal(dst_a, host(d)), (seed, rate, w, h)


def test_luma_search_1step_ushort_wrap(eng):
    """Q2: costs above 65535 wrap in the reference's unsigned short accumulator"""
    o = oracle()
    w, h, rate = 64, 48, 1
    r = rng(5)
    prev = (r.integers(0, 2, size=(h, w)) * 255).astype(np.uint8)
    cur = (255 - prev).astype(np.uint8)  # maximal residuals everywhere
    net_width, net_h = 8, 8
    src_net = np.zeros((net_h * net_width, 2), np.int16)
    dst_a = np.zeros((net_h * net_width, 2), np.int16)
    o.vp8o_luma_search_1step(P(cur), P(prev), P(src_net), P(dst_a), net_width, w, h, rate)
    d = dev(np.zeros((net_h * net_width, 2), np.int16))
    eng.luma_search_1step(dev(cur), dev(prev), dev(src_net), d, net_width, w, h, rate)
    assert np.array_equal(dst_a, host(d))


@pytest.mark.parametrize("w,h", [(64, 48), (96, 80), (352, 288)])
def test_luma_search_2step(eng, w, h):
    o = oracle()
    for seed in (20, 21, 22, 23):
        r = rng(seed + w)
        ref_img = textured(r, h, w, smooth=3)
        if seed == 22:
            ref_img = (r.integers(0, 2, size=(h, w)) * 255).astype(np.uint8)
        cur = np.clip(np.roll(ref_img, (1, -2), axis=(0, 1)).astype(np.int32) + r.integers(-5, 6, size=(h, w)), 0, 255).astype(np.uint8)
        if seed == 23:  # half-pel content: blend of two shifts
            cur = ((ref_img.astype(np.int32) + np.roll(ref_img, 1, axis=1)) // 2).astype(np.uint8)
        nb = w * h // 64
        net = r.integers(-9, 10, size=(nb, 2)).astype(np.int16)
        if seed == 23:
            net[:] = 0
        out_a = np.zeros((nb, 2), np.int16)
        m_a = np.zeros(nb, np.int32)
        o.vp8o_luma_search_2step(P(cur), P(ref_img), P(net), P(out_a), P(m_a), w, h)
        d_out = torch.zeros((nb, 2), dtype=torch.int16, device="cuda")
        d_m = torch.zeros(nb, dtype=torch.int32, device="cuda")
        eng.luma_search_2step(dev(cur), dev(ref_img), dev(net), d_out, d_m, w, h)
        assert np.array_equal(out_a, host(d_out)), seed
        assert np.array_equal(m_a, host(d_m)), seed


def test_select_reference_and_pack(eng):
    o = oracle()
    w, h = 96, 64
    M = (w // 16) * (h // 16)
    r = rng(30)
    for use_g, use_a in ((0, 0), (1, 0), (0, 1), (1, 1)):
        nets = [r.integers(-3, 4, size=(M * 4, 2)).astype(np.int16) for _ in range(3)]
        nets[0][:8] = 1
        mets = [r.integers(0, 50, size=M * 4).astype(np.int32) for _ in range(3)]
        ref_a = np.zeros(M, np.int32)
        vec_a = np.zeros((M, 8), np.int16)
        o.vp8o_select_reference(P(nets[0]), P(nets[1]), P(nets[2]), P(mets[0]), P(mets[1]), P(mets[2]), P(ref_a),
                                P(vec_a), w, h, use_g, use_a)
        d_ref = torch.zeros(M, dtype=torch.int32, device="cuda")
        d_vec = torch.zeros((M, 8), dtype=torch.int16, device="cuda")
        eng.select_reference(dev(nets[0]), dev(nets[1]), dev(nets[2]), dev(mets[0]), dev(mets[1]), dev(mets[2]), d_ref,
                             d_vec, w, h, use_g, use_a)
        assert np.array_equal(ref_a, host(d_ref)) and np.array_equal(vec_a, host(d_vec))
        parts_a = np.zeros(M, np.int32)
        s_a = np.zeros(M, np.float32)
        o.vp8o_pack_8x8_into_16x16(P(vec_a), P(parts_a), P(s_a), M)
        d_parts = torch.zeros(M, dtype=torch.int32, device="cuda")
        d_s = torch.zeros(M, dtype=torch.float32, device="cuda")
        eng.pack_8x8_into_16x16(d_vec, d_parts, d_s)
        assert np.array_equal(parts_a, host(d_parts)) and np.array_equal(s_a, host(d_s))


@pytest.mark.parametrize("plane", [0, 1, 2])
def test_prepare_predictors_and_residual(eng, plane):
    o = oracle()
    W, H = 96, 64
    w, h = (W, H) if plane == 0 else (W // 2, H // 2)
    M = (W // 16) * (H // 16)
    for seed in (40, 41, 42):
        r = rng(seed + plane)
        img = (r.integers(0, 2, size=(h, w)) * 255).astype(np.uint8) if seed == 41 else textured(r, h, w, 2)
        if seed == 42:
            img = np.zeros((h, w), np.uint8)
            img[:, ::2] = 255
        cur = r.integers(0, 256, size=(h, w)).astype(np.uint8)
        refs = r.integers(0, 3, size=M).astype(np.int32)
        vec = np.zeros((M, 4, 2), np.int16)
        for mb in range(M):
            for q in range(4):
                bx = (mb % (W // 16)) * 16 + (q % 2) * 8
                by = (mb // (W // 16)) * 16 + (q // 2) * 8
                vec[mb, q, 0] = r.integers(-4 * bx, 4 * (W - 8 - bx) + 1)
                vec[mb, q, 1] = r.integers(-4 * by, 4 * (H - 8 - by) + 1)
        for ref_id in range(3):
            pa = np.full((h, w), 9, np.uint8)
            ra = np.full((h, w), 9, np.int16)
            o.vp8o_prepare_predictors_and_residual(P(cur), P(img), P(pa), P(ra), P(refs), P(vec), w, h, plane, ref_id)
            d_p = dev(np.full((h, w), 9, np.uint8))
            d_r = dev(np.full((h, w), 9, np.int16))
            eng.prepare_predictors_and_residual(dev(cur), dev(img), d_p, d_r, dev(refs), dev(vec), w, h, plane, ref_id)
            assert np.array_equal(pa, host(d_p)), (seed, ref_id)
            assert np.array_equal(ra, host(d_r)), (seed, ref_id)


@pytest.mark.parametrize("qi", [0, 24, 60, 127])
def test_dct_wht_idct_chain(eng, qi):
    o = oracle()
    W, H = 64, 48
    M = (W // 16) * (H // 16)
    for ssim_mix, target in ((False, -1.0), (True, 0.9)):
        r = rng(50 + qi)
        res = [r.integers(-255, 256, size=(H, W)).astype(np.int16), r.integers(-80, 81, size=(H // 2, W // 2)).astype(np.int16),
               r.integers(-80, 81, size=(H // 2, W // 2)).astype(np.int16)]
        parts = r.integers(0, 2, size=M).astype(np.int32)
        ssim = np.full(M, -2.0, np.float32)
        if ssim_mix:
            ssim = r.choice(np.array([-2.0, 0.5, 0.97], np.float32), size=M)
        sd = make_segment_data((max(0, qi - 6), max(0, qi - 4), max(0, qi - 2), qi))
        coef_a = r.integers(-5, 6, size=(M, 400)).astype(np.int16)
        seg_a = r.integers(0, 4, size=M).astype(np.int32)
        pred = [r.integers(0, 256, size=x.shape).astype(np.uint8) for x in res]
        rec_a = [np.full(x.shape, 3, np.uint8) for x in res]
        d_coef, d_seg, d_parts, d_ssim, d_sd = dev(coef_a), dev(seg_a), dev(parts), dev(ssim), dev(sd)
        d_res = [dev(x) for x in res]
        d_pred = [dev(x) for x in pred]
        d_rec = [dev(x) for x in rec_a]
        for s in (3, 2, 1, 0):
            for p in range(3):
                w, h = (W, H) if p == 0 else (W // 2, H // 2)
                o.vp8o_dct4x4(P(res[p]), P(coef_a), P(seg_a), P(parts), P(ssim), w, h, P(sd), s, ctypes.c_float(target), p)
                eng.dct4x4(d_res[p], d_coef, d_seg, d_parts, d_ssim, w, h, d_sd, s, target, p)
            assert np.array_equal(coef_a, host(d_coef)) and np.array_equal(seg_a, host(d_seg))
            o.vp8o_wht4x4_iwht4x4(P(coef_a), P(seg_a), P(parts), P(sd), s, M)
            eng.wht4x4_iwht4x4(d_coef, d_seg, d_parts, d_sd, s)
            assert np.array_equal(coef_a, host(d_coef))
            for p in range(3):
                w, h = (W, H) if p == 0 else (W // 2, H // 2)
                o.vp8o_idct4x4(P(rec_a[p]), P(pred[p]), P(coef_a), P(seg_a), P(parts), w, h, P(sd), s, p)
                eng.idct4x4(d_rec[p], d_pred[p], d_coef, d_seg, d_parts, w, h, d_sd, s, p)
                assert np.array_equal(rec_a[p], host(d_rec[p]))


def test_ssim_bit_exact(eng):
    o = oracle()
    W, H = 352, 288
    M = (W // 16) * (H // 16)
    for seed in (60, 61, 62):
        r = rng(seed)
        a = textured(r, H, W, 2)
        noise = r.integers(-20, 21, size=(H, W)) if seed != 62 else r.integers(-2, 3, size=(H, W)) + 9
        b = np.clip(a.astype(np.int32) + noise, 0, 255).astype(np.uint8)
        seg = r.integers(0, 2, size=M).astype(np.int32)
        for mbs, w, h in ((16, W, H), (8, W // 2, H // 2)):
            fa = np.ascontiguousarray(a[:h, :w])
            fb = np.ascontiguousarray(b[:h, :w])
            ma = np.full(M, 5.0, np.float32)
            o.vp8o_count_SSIM(P(fa), P(fb), P(seg), P(ma), w, h, 1, mbs)
            d_m = dev(np.full(M, 5.0, np.float32))
            eng.count_SSIM(dev(fa), dev(fb), dev(seg), d_m, w, h, 1, mbs)
            assert np.array_equal(ma.view(np.uint32), host(d_m).view(np.uint32))
        m = [r.random(M).astype(np.float32) for _ in range(3)]
        ga = np.zeros(M, np.float32)
        o.vp8o_gather_SSIM(P(m[0]), P(m[1]), P(m[2]), P(ga), M)
        d_g = torch.zeros(M, dtype=torch.float32, device="cuda")
        eng.gather_SSIM(dev(m[0]), dev(m[1]), dev(m[2]), d_g)
        assert np.array_equal(ga.view(np.uint32), host(d_g).view(np.uint32))


def test_filter_mask(eng):
    o = oracle()
    W, H = 352, 288
    M = (W // 16) * (H // 16)
    r = rng(70)
    coef = (r.integers(-300, 301, size=(M, 400)) * (r.random((M, 400)) < 0.05)).astype(np.int16)
    coef[0] = 0
    coef[1] = 0
    coef[1, 24 * 16 + 3] = -7
    coef[2] = 0
    coef[2, 5 * 16] = 4
    coef[3] = -32768 // 2
    parts = r.integers(0, 3, size=M).astype(np.int32)
    parts[:3] = 0
    nz_a = np.zeros(M, np.int32)
    mk_a = np.zeros(M, np.int32)
    o.vp8o_prepare_filter_mask(P(coef), P(nz_a), P(parts), P(mk_a), W, H)
    d_nz = torch.zeros(M, dtype=torch.int32, device="cuda")
    d_mk = torch.zeros(M, dtype=torch.int32, device="cuda")
    eng.prepare_filter_mask(dev(coef), d_nz, dev(parts), d_mk, W, H)
    assert np.array_equal(nz_a, host(d_nz)) and np.array_equal(mk_a, host(d_mk))


def _lf_case(seed, mbw, mbh, n):
    r = rng(seed)
    w, h = mbw * n, mbh * n
    lv = 128 + np.cumsum(r.integers(-6, 7, size=(h // 4, w // 4)), axis=1) + np.cumsum(r.integers(-6, 7, size=(h // 4, w // 4)), axis=0)
    frame = (np.kron(lv, np.ones((4, 4), np.int64)) + r.integers(-2, 3, size=(h, w))).clip(0, 255).astype(np.uint8)
    if seed % 4 == 1:
        blocks = r.choice(np.array([0, 3, 252, 255], np.uint8), size=(h // 4, w // 4))
        frame = (np.kron(blocks, np.ones((4, 4), np.uint8)).astype(np.int32) + r.integers(-3, 4, size=(h, w))).clip(0, 255).astype(np.uint8)
    return r, w, h, frame


@pytest.mark.parametrize("mb_size", [16, 8])
@pytest.mark.parametrize("mbw,mbh", [(7, 5), (22, 18), (3, 70), (40, 3)])
def test_loop_filter(eng, mb_size, mbw, mbh):
    o = oracle()
    M = mbw * mbh
    for seed, levels, sharp in ((80, (10, 20, 40, 63), 0), (81, (63, 63, 63, 63), 0), (82, (30, 8, 0, 50), 3), (83, (5, 5, 5, 5), 7)):
        r, w, h, frame = _lf_case(seed, mbw, mbh, mb_size)
        seg = r.integers(0, 4, size=M).astype(np.int32)
        if seed == 82:
            seg[: M // 2] = r.choice(np.array([0, 1, 3], np.int32), size=M // 2)
        mask = r.choice(np.array([0, -1], np.int32), size=M)
        sd = make_segment_data(lf_level=levels, sharpness=sharp)
        fa = frame.copy()
        o.vp8o_loop_filter_frame(P(fa), P(seg), P(mask), P(sd), w, h, mb_size)
        d_f = dev(frame)
        eng.loop_filter_frame(d_f, dev(seg), dev(mask), dev(sd), w, h, mb_size)
        assert np.array_equal(fa, host(d_f)), (seed, mb_size, mbw, mbh)


def test_loop_filter_three_planes_1080p(eng):
    o = oracle()
    W, H = 1920, 1088
    M = (W // 16) * (H // 16)
    r, _, _, y = _lf_case(90, W // 16, H // 16, 16)
    _, _, _, u = _lf_case(91, W // 16, H // 16, 8)
    _, _, _, v = _lf_case(92, W // 16, H // 16, 8)
    seg = r.integers(0, 4, size=M).astype(np.int32)
    mask = r.choice(np.array([0, -1], np.int32), size=M)
    sd = make_segment_data(lf_level=(12, 20, 33, 50))
    ya, ua, va = y.copy(), u.copy(), v.copy()
    o.vp8o_loop_filter_frame(P(ya), P(seg), P(mask), P(sd), W, H, 16)
    o.vp8o_loop_filter_frame(P(ua), P(seg), P(mask), P(sd), W // 2, H // 2, 8)
    o.vp8o_loop_filter_frame(P(va), P(seg), P(mask), P(sd), W // 2, H // 2, 8)
    dy, du, dv = dev(y), dev(u), dev(v)
    eng.loop_filter_planes(dy, du, dv, dev(seg), dev(mask), dev(sd), W, H)
    assert np.array_equal(ya, host(dy)) and np.array_equal(ua, host(du)) and np.array_equal(va, host(dv))


@pytest.mark.parametrize("cluster", ["1", "2", "5"])
def test_loop_filter_other_cluster_sizes(cluster):
    """the hand-off between rows has two forms: distributed shared memory inside a cluster of rows, the global mailbox
    at cluster borders (and everywhere with VP8B200_LF_CLUSTER=1).  The default (8) is what every other test runs;
    here the same loop-filter cases run with no clusters, with pairs, and with a size that does not divide the rows
    (the variable is read once per process, hence the child process)."""
    import subprocess
    import sys
    env = dict(os.environ, VP8B200_LF_CLUSTER=cluster)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "test_loop_filter and not other_cluster", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-500:]


@pytest.mark.parametrize("w,h", [(352, 288), (208, 176), (1920, 1088)])
def test_luma_search_2step_tma_variant_identical(w, h):
    """the TMA-staged experiment kernel (vp8b200_experiment_search_2step_tma: box copies out of a replicate-padded
    plane) finds the same vectors and metrics as the production kernel, frame edges included"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import me_tma_ab
    r = me_tma_ab.ab(w, h, reps=1)
    assert r["identical"], r
    assert r["nonzero_vectors"] > 0


@pytest.mark.parametrize("w,h,q,seed", [(176, 144, (19, 24, 7, 10), 0), (352, 288, (8, 6, 4, 4), 0), (64, 48, (157, 284, 132, 284), 0),
                                        (208, 176, (4, 4, 4, 4), 0), (1920, 1088, (19, 24, 7, 10), 0), (96, 80, (7, 6, 6, 9), 5),
                                        (16, 16, (12, 12, 12, 12), 6), (32, 160, (5, 9, 9, 5), 7)])
def test_intra_frame_vs_oracle(w, h, q, seed):
    """the key-frame path (B_PRED mode decision, TM chroma, transform, quantise, reconstruct) as a
    wavefront kernel against the oracle's restatement of src/intra_part.h (itself pinned against the reference's own
    code, tests/test_intra_oracle.py): modes, coefficients, all three reconstructed planes"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_y4m
    from vp8oclenc_b200 import host as eng
    if seed == 0:
        y, u, v = (np.ascontiguousarray(p).reshape(-1) for p in gen_y4m.Clip(max(w, 128), max(h, 128)).frame(3))
        if w < 128 or h < 128:
            full = gen_y4m.Clip(max(w, 128), max(h, 128)).frame(3)
            y = np.ascontiguousarray(full[0][:h, :w]).reshape(-1)
            u = np.ascontiguousarray(full[1][:h // 2, :w // 2]).reshape(-1)
            v = np.ascontiguousarray(full[2][:h // 2, :w // 2]).reshape(-1)
    else:
        r = rng(seed)
        yy = r.integers(0, 256, size=(h, w), dtype=np.uint8)
        yy[: h // 2, : w // 2] = np.where(r.integers(0, 2, size=(h // 2, w // 2)) > 0, 255, 0)
        y = np.ascontiguousarray(yy).reshape(-1)
        u = r.integers(0, 256, size=w * h // 4, dtype=np.uint8)
        v = np.where(r.integers(0, 2, size=w * h // 4) > 0, 250, 3).astype(np.uint8)
    M = (w // 16) * (h // 16)
    qa = np.asarray(q, np.int32)
    ry, ru, rv = np.zeros(w * h, np.uint8), np.zeros(w * h // 4, np.uint8), np.zeros(w * h // 4, np.uint8)
    mb = np.zeros(M * 400, np.int16)
    modes, parts, seg = np.zeros(M * 16, np.int32), np.full(M, -1, np.int32), np.full(M, -1, np.int32)
    oracle().vp8o_intra_frame(w, h, P(y), P(u), P(v), P(ry), P(ru), P(rv), P(mb), P(modes), P(parts), P(seg), P(qa))
    d = {k: torch.from_numpy(a).cuda() for k, a in dict(y=y, u=u, v=v).items()}
    g = dict(ry=torch.zeros(w * h, dtype=torch.uint8, device="cuda"), ru=torch.zeros(w * h // 4, dtype=torch.uint8, device="cuda"),
             rv=torch.zeros(w * h // 4, dtype=torch.uint8, device="cuda"), mb=torch.zeros(M * 400, dtype=torch.int16, device="cuda"),
             modes=torch.zeros(M * 16, dtype=torch.int32, device="cuda"), parts=torch.full((M,), -1, dtype=torch.int32, device="cuda"),
             seg=torch.full((M,), -1, dtype=torch.int32, device="cuda"))
    keep = eng.intra_frame(d["y"], d["u"], d["v"], g["ry"], g["ru"], g["rv"], g["mb"], g["modes"], g["parts"], g["seg"], w, h, q)
    torch.cuda.synchronize()
    del keep
    assert np.array_equal(g["modes"].cpu().numpy(), modes), "sub-block modes"
    assert np.array_equal(g["mb"].cpu().numpy().reshape(M, 25, 16)[:, :24], mb.reshape(M, 25, 16)[:, :24]), "coefficients"
    assert np.array_equal(g["ry"].cpu().numpy(), ry), "luma reconstruction"
    assert np.array_equal(g["ru"].cpu().numpy(), ru) and np.array_equal(g["rv"].cpu().numpy(), rv), "chroma reconstruction"
    assert np.array_equal(g["parts"].cpu().numpy(), parts) and np.array_equal(g["seg"].cpu().numpy(), seg)
