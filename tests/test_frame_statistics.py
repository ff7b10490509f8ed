"""the host's per-frame O(N) reductions (get_loopfilter_strength, the chroma differences of
scene_change).  The oracle's restatement is pinned against the REAL host: what prepare_segments_data()
(src/vp8enc.cpp:129-221) uploads as segment data every frame follows from the reductor and the sharpness the host
computed for that frame, so the traced uploads must equal the table built from the oracle's numbers.  The CUDA kernel
(vp8b200_frame_statistics) is then compared with the oracle, wrap-around of the `int` accumulators at 7680x4320
included."""
import ctypes
import os
import sys

import numpy as np
import pytest

import _trace
from _libs import P, ROOT, oracle, ref

sys.path.insert(0, os.path.join(ROOT, "tools"))


def oracle_stats(y, u0, u1, v0, v1, w, h):
    out = np.zeros(4, np.int32)
    oracle().vp8o_frame_statistics(P(y) if y is not None else None, w, h, P(u0) if u0 is not None else None,
                                   P(u1) if u1 is not None else None, P(v0) if v0 is not None else None,
                                   P(v1) if v1 is not None else None, P(out))
    return out


@pytest.mark.skipif(ref() is None or not _trace.have_host(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_statistics_reproduce_the_hosts_segment_data(tmp_path):
    import gen_y4m
    from vp8oclenc_b200.hostlogic import make_segment_data
    w, h, n, qi = 352, 288, 10, 24
    d = str(tmp_path)
    y4m, trace = os.path.join(d, "clip.y4m"), os.path.join(d, "t.trace")
    gen_y4m.write_y4m(y4m, w, h, n)
    _trace.run_host(_trace.REF_DIR, d, y4m, os.path.join(d, "o.ivf"),
                    ["-qmin", qi, "-qmax", qi, "-g", 100, "-altref-range", 4, "-partitions", 2, "-threads", 2], trace=trace)
    frames, _ = _trace.split_frames(_trace.read_trace(trace))
    checked, seen = 0, set()
    for fr in frames:
        if "current_frame_Y" not in fr["w"]:
            continue  # key frame
        y = _trace.arr(fr["w"]["current_frame_Y"][0], np.uint8)
        st = oracle_stats(y, None, None, None, None, w, h)
        got = _trace.arr(fr["w"]["segments_data_gpu"][0], np.int32, (4, 11))
        # (the quantiser index of every segment is the host's rate logic; level and limits follow from the statistics)
        want = make_segment_data(tuple(int(x) for x in got[:, 0]), key=False, sharpness=int(st[1]), reductor=int(st[0]))
        assert np.array_equal(got, want), (st, got, want)
        checked += 1
        seen.add((int(st[0]), int(st[1])))
    assert checked >= n - 2
    assert all(3 <= r <= 8 and 0 <= s <= 7 for r, s in seen)


def test_scene_change_differences_match_the_planner(tmp_path):
    """the same numbers segments.plan_key_frames() derives (in numpy) from the input file"""
    import gen_y4m
    w, h = 176, 144
    clip = gen_y4m.Clip(w, h)
    (_, u0, v0), (_, u1, v1) = clip.frame(3), clip.frame(4)
    u1 = (255 - u1.astype(np.int32)).clip(0, 255).astype(np.uint8)
    st = oracle_stats(None, np.ascontiguousarray(u0), np.ascontiguousarray(u1), np.ascontiguousarray(v0), np.ascontiguousarray(v1), w, h)
    n = (w // 2) * (h // 2)
    assert st[2] == int(np.abs(u0.astype(np.int32) - u1).sum()) // n and st[2] > 7
    assert st[3] == int(np.abs(v0.astype(np.int32) - v1).sum()) // n


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,seed", [(176, 144, 1), (1920, 1088, 2), (200, 120, 3), (7680, 4320, 4)])
def test_frame_statistics_kernel_vs_oracle(w, h, seed):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vp8oclenc_b200 import host as eng
    r = np.random.default_rng(seed)
    if w == 7680:  # bright, busy content: both `int` accumulators of the reference wrap at this size
        y = r.integers(96, 256, size=w * h, dtype=np.uint8)
    else:
        y = np.clip(r.normal(120, 30, size=w * h), 0, 255).astype(np.uint8)
    planes = [r.integers(0, 256, size=(w // 2) * (h // 2), dtype=np.uint8) for _ in range(4)]
    want = oracle_stats(y, planes[0], planes[1], planes[2], planes[3], w, h)
    if w == 7680:
        assert int(y.astype(np.int64).sum()) > 2 ** 31  # the case is what it claims to be
    dy = torch.from_numpy(y).cuda()
    dp = [torch.from_numpy(p).cuda() for p in planes]
    scratch = torch.zeros(4, dtype=torch.int64, device="cuda")
    out = torch.zeros(4, dtype=torch.int32, device="cuda")
    L = eng.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    D = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    rc = L.vp8b200_frame_statistics(st, D(dy), w, h, D(dp[0]), D(dp[1]), D(dp[2]), D(dp[3]), D(scratch), D(out))
    assert rc == 0
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want), (out.cpu().numpy(), want)
    # luma only / chroma only leave the other half of the result alone
    out.fill_(-99)
    assert L.vp8b200_frame_statistics(st, D(dy), w, h, None, None, None, None, D(scratch), D(out)) == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.array_equal(got[:2], want[:2]) and (got[2:] == -99).all()
