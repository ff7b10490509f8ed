"""the intra (key-frame) macroblock path.  The oracle's restatement (oracle/vp8_oracle_intra.c) against
the REFERENCE's own code -- src/intra_part.h compiled from where it lies into oracle/_ref/libref_intra.so
(oracle/ref_intra.cpp) -- on synthetic and adversarial frames: coefficients, sub-block modes, reconstruction of all
three planes, bit for bit."""
import os
import sys

import numpy as np
import pytest

from _libs import GoldenNumpy, P, ROOT, oracle, ref_intra

sys.path.insert(0, os.path.join(ROOT, "tools"))
# VP8_GOLDEN_INTRA=record:<file> | check:<file>: golden vectors of the reference's intra path, see tests/golden/make_golden.py
GOLDEN = os.environ.get("VP8_GOLDEN_INTRA", "")
if GOLDEN:
    np = GoldenNumpy(np, GOLDEN, oracle_first=False)
CHECK_ONLY = GOLDEN.startswith("check:")
pytestmark = pytest.mark.skipif(ref_intra() is None and not CHECK_ONLY, reason="oracle/_ref/libref_intra.so not built (needs /root/reference)")


def run_both(y, u, v, w, h, quants):
    M = (w // 16) * (h // 16)
    q = np.asarray(quants, np.int32)
    res = []
    for lib, fn in ((ref_intra(), "vp8ref_intra_frame"), (oracle(), "vp8o_intra_frame")):
        ry, ru, rv = np.zeros(w * h, np.uint8), np.zeros(w * h // 4, np.uint8), np.zeros(w * h // 4, np.uint8)
        mb = np.zeros(M * 400, np.int16)
        modes, parts, seg = np.zeros(M * 16, np.int32), np.full(M, -1, np.int32), np.full(M, -1, np.int32)
        if lib is not None:  # (no reference in golden-check mode: its outputs are the stored vectors)
            getattr(lib, fn)(w, h, P(y), P(u), P(v), P(ry), P(ru), P(rv), P(mb), P(modes), P(parts), P(seg), P(q))
        res.append(dict(ry=ry, ru=ru, rv=rv, mb=mb.reshape(M, 25, 16), modes=modes, parts=parts, seg=seg))
    return res


def check(y, u, v, w, h, quants):
    a, b = run_both(y, u, v, w, h, quants)
    assert np.array_equal(a["modes"], b["modes"]), "sub-block modes"
    assert np.array_equal(a["mb"][:, :24], b["mb"][:, :24]), "coefficients"
    for k in ("ry", "ru", "rv", "parts", "seg"):
        assert np.array_equal(a[k], b[k]), k
    return b


@pytest.mark.parametrize("w,h,q", [(176, 144, (19, 24, 7, 10)), (352, 288, (8, 6, 4, 4)), (64, 48, (157, 284, 132, 284)), (208, 176, (4, 4, 4, 4))])
def test_intra_frame_clip_content(w, h, q):
    import gen_y4m
    y, u, v = (np.ascontiguousarray(p).reshape(-1) for p in gen_y4m.Clip(w, h).frame(3))
    a = check(y, u, v, w, h, q)
    assert len(set(a["modes"].tolist())) >= 6          # most of the ten modes are in play
    assert (a["parts"] == 2).all() and (a["seg"] == 0).all()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_intra_frame_adversarial_content(seed):
    """noise, hard edges and saturated blocks: 16-bit stores, the coefficient-11 rounding and the clamps all get used"""
    w, h = 96, 80
    r = np.random.default_rng(seed)
    y = r.integers(0, 256, size=(h, w), dtype=np.uint8)
    y[16:48, 16:64] = np.where(r.integers(0, 2, size=(32, 48)) > 0, 255, 0)
    y[48:, :32] = 255
    y[:16, 64:] = 0
    u = r.integers(0, 256, size=(h // 2, w // 2), dtype=np.uint8)
    v = np.where(r.integers(0, 2, size=(h // 2, w // 2)) > 0, 250, 3).astype(np.uint8)
    a = check(np.ascontiguousarray(y).reshape(-1), np.ascontiguousarray(u).reshape(-1), np.ascontiguousarray(v).reshape(-1), w, h,
              (4 + 3 * seed, 5 + seed, 6, 9))
    assert np.abs(a["mb"][:, :24].astype(np.int32)).max() > 20
