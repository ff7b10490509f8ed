"""Host logic of the shim: the three boolean-coder kernels of the reference's "CPU program"
(count_probs, num_div_denom, encode_coefficients) as re-implemented in
vp8oclenc_b200/csrc/entropy_host.cpp, against the reference's own kernels (oracle/_ref).
They decide the bytes of the coefficient partitions of every frame, so they are on the .ivf
bit-exactness path.  No GPU needed: the functions run on host threads.
"""
import ctypes
import os

import numpy as np
import pytest

from _libs import P, ROOT, ref, ref_run

SHIM = os.path.join(ROOT, "vp8oclenc_b200", "lib", "libOpenCL.so.1")
pytestmark = pytest.mark.skipif(ref() is None, reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(SHIM):
        from vp8oclenc_b200 import build
        build.build_shim()
    return ctypes.CDLL(SHIM)


def make_case(seed, mbw, mbh, density, big):
    r = np.random.default_rng(seed)
    M = mbw * mbh
    mag = r.integers(1, 2100 if big else 12, size=(M, 400))
    coef = (mag * r.choice([-1, 1], size=(M, 400)) * (r.random((M, 400)) < density)).astype(np.int16)
    parts = r.choice(np.array([0, 0, 1, 2], np.int32), size=M)
    # whole macroblocks without coefficients (skipped), some with only a Y2 block
    coef[r.random(M) < 0.2] = 0
    nz = np.zeros(M, np.int32)
    for mb in range(M):
        m = coef[mb].reshape(25, 16).astype(np.int64)
        s = np.abs(m[:16, 1:]).sum() + np.abs(m[16:24]).sum()
        s += np.abs(m[24]).sum() if parts[mb] == 0 else np.abs(m[:16, 0]).sum()
        nz[mb] = s
    return coef, parts, nz


@pytest.mark.parametrize("nparts", [1, 2, 4, 8])
@pytest.mark.parametrize("seed,density,big", [(1, 0.05, False), (2, 0.3, False), (3, 0.6, True), (4, 0.01, True)])
def test_entropy_kernels_match_reference(shim, nparts, seed, density, big):
    mbw, mbh = 11, 9
    M = mbw * mbh
    coef, parts, nz = make_case(seed * 10 + nparts, mbw, mbh, density, big)
    # the encoder uses sizeof(short)*800*mb_count/2/partitions (src/init.h:409,1190); the synthetic
    # dense cases here need more room than real frames do
    step = 2400 * M // nparts

    # reference
    probs_r = np.full(8 * 1056, 7, np.uint32)
    den_r = np.full(8 * 1056, 7, np.uint32)
    ctx_r = np.full(M * 25, 9, np.uint8)
    ref_run("count_probs", nparts, 1, [coef, nz, parts, probs_r, den_r, ctx_r, mbh, mbw, nparts, step])
    # ours
    probs_o = np.full(8 * 1056, 7, np.uint32)
    den_o = np.full(8 * 1056, 7, np.uint32)
    ctx_o = np.full(M * 25, 9, np.uint8)
    shim.vp8b200_host_count_probs(P(coef), P(nz), P(parts), P(probs_o), P(den_o), P(ctx_o), mbh, mbw, nparts)
    assert np.array_equal(probs_r, probs_o), "token statistics (numerators)"
    assert np.array_equal(den_r, den_o), "token statistics (denominators)"
    assert np.array_equal(ctx_r, ctx_o), "neighbour contexts"

    ref_run("num_div_denom", nparts, 1, [probs_r, den_r, nparts])
    shim.vp8b200_host_num_div_denom(P(probs_o), P(den_o), nparts)
    assert np.array_equal(probs_r, probs_o), "probabilities"

    out_r = np.zeros(step * nparts, np.uint8)
    out_o = np.zeros(step * nparts, np.uint8)
    sz_r = np.zeros(8, np.int32)
    sz_o = np.zeros(8, np.int32)
    ref_run("encode_coefficients", nparts, 1, [coef, nz, parts, out_r, sz_r, ctx_r, probs_r, mbh, mbw, nparts, step])
    shim.vp8b200_host_encode_coefficients(P(coef), P(nz), P(parts), P(out_o), P(sz_o), P(ctx_o), P(probs_o), mbh, mbw,
                                          nparts, step)
    assert np.array_equal(sz_r, sz_o), "partition sizes"
    assert np.array_equal(out_r, out_o), "partition bytes"
    assert sz_r[:nparts].min() >= 4


def _clz8(r):
    s = 0
    while r < 128:
        r <<= 1
        s += 1
    return s


@pytest.mark.parametrize("seed,n,mode", [(1, 0, "random"), (2, 1, "random"), (3, 7, "random"), (4, 300, "likely"),
                                         (5, 3000, "likely"), (6, 3000, "extreme"), (7, 2500, "random"),
                                         (8, 4000, "carries")])
def test_bool_coder_output_is_one_big_sum(shim, seed, n, mode):
    """The identity the GPU bool coder (entropy_kernels.cu) is built on, checked here on
    the CPU against the serial coder of entropy_host.cpp (itself pinned against the reference's kernel): the
    bytes of a partition are the big-endian digits of  sum_{bit_i = 1} split_i << (T' - T_i),  T' = 24 + 8 m,
    m + 4 bytes, m = the bytes that leave during coding.  Carry propagation is just the carries of the sum."""
    r = np.random.default_rng(seed)
    if mode == "random":
        prob = r.integers(1, 256, size=n)
        bit = r.integers(0, 2, size=n)
    elif mode == "likely":  # decisions that follow their probabilities
        prob = r.integers(1, 256, size=n)
        bit = (r.random(n) > prob / 256.0).astype(np.int64)
    elif mode == "extreme":
        prob = r.choice(np.array([1, 2, 128, 250, 254, 255]), size=n)
        bit = r.integers(0, 2, size=n)
    else:  # long runs of the improbable value: bottom creeps up to the top of the range, carries ripple
        prob = np.full(n, 255)
        bit = (r.random(n) < 0.97).astype(np.int64)
    # the stream format of the GPU tokens: bit 15 value, bits 0-10 = 1056 + fixed probability
    tokens = ((bit.astype(np.uint32) << 15) | (1056 + prob.astype(np.uint32))).astype(np.uint16)
    info = np.array([0, n, n], np.uint32)
    probs = np.zeros(1056, np.uint32)
    out = np.zeros(n + 64, np.uint8)
    size = np.zeros(8, np.int32)
    shim.vp8b200_host_encode_token_streams(P(tokens), P(info), P(probs), P(out), P(size), 1, int(out.size))
    # the sum
    R, T, L = 255, 0, []
    for p, b in zip(prob.tolist(), bit.tolist()):
        split = 1 + (((R - 1) * p) >> 8)
        if b:
            L.append((split, T))
        rr = R - split if b else split
        s = _clz8(rr)
        R, T = rr << s, T + s
    m = (T - 24) // 8 + 1 if T >= 24 else 0
    total = sum(a << (24 + 8 * m - t) for a, t in L)
    assert total < 1 << (8 * (m + 4))
    assert int(size[0]) == m + 4
    assert bytes(out[:m + 4]) == total.to_bytes(m + 4, "big")
