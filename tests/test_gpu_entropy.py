"""GPU half of the coefficient entropy coding (vp8oclenc_b200/csrc/entropy_kernels.cu) against the host
implementation (entropy_host.cpp), which tests/test_host_entropy.py pins against the reference's own
count_probs / num_div_denom / encode_coefficients kernels: statistics tables, neighbour contexts and
the bytes of every partition must be identical.
"""
import ctypes
import os

import numpy as np
import pytest

from _libs import ROOT

pytestmark = pytest.mark.gpu
SHIM = os.path.join(ROOT, "vp8oclenc_b200", "lib", "libOpenCL.so.1")


def make_case(seed, mbw, mbh, density, big):
    r = np.random.default_rng(seed)
    M = mbw * mbh
    mag = r.integers(1, 2100 if big else 12, size=(M, 400))
    coef = (mag * r.choice([-1, 1], size=(M, 400)) * (r.random((M, 400)) < density)).astype(np.int16)
    # runs of empty blocks, as real frames have them
    coef.reshape(M * 25, 16)[r.random(M * 25) < 0.5] = 0
    parts = r.choice(np.array([0, 0, 1, 2], np.int32), size=M)
    coef[r.random(M) < 0.2] = 0
    nz = np.zeros(M, np.int32)
    for mb in range(M):
        m = coef[mb].reshape(25, 16).astype(np.int64)
        s = np.abs(m[:16, 1:]).sum() + np.abs(m[16:24]).sum()
        s += np.abs(m[24]).sum() if parts[mb] == 0 else np.abs(m[:16, 0]).sum()
        nz[mb] = s
    return coef, parts, nz


@pytest.mark.parametrize("nparts", [1, 2, 4, 8])
@pytest.mark.parametrize("seed,mbw,mbh,density,big", [(1, 11, 9, 0.05, False), (2, 22, 18, 0.3, False), (3, 11, 9, 0.7, True),
                                                      (4, 45, 23, 0.02, True), (5, 120, 68, 0.03, False), (6, 9, 3, 1.0, True)])
def test_token_streams_match_host_path(nparts, seed, mbw, mbh, density, big):
    import torch
    from vp8oclenc_b200 import host as eng
    L = eng.lib()
    H = ctypes.CDLL(SHIM)
    M = mbw * mbh
    coef, parts, nz = make_case(seed * 10 + nparts, mbw, mbh, density, big)
    step = 2400 * M // nparts
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)

    # host path
    probs_h = np.full(8 * 1056, 7, np.uint32)
    den_h = np.full(8 * 1056, 7, np.uint32)
    ctx_h = np.full(M * 25, 9, np.uint8)
    H.vp8b200_host_count_probs(P(coef), P(nz), P(parts), P(probs_h), P(den_h), P(ctx_h), mbh, mbw, nparts)
    stats_h = probs_h.copy()
    H.vp8b200_host_num_div_denom(P(probs_h), P(den_h), nparts)
    out_h = np.zeros(step * nparts + 64, np.uint8)
    size_h = np.zeros(8, np.int32)
    H.vp8b200_host_encode_coefficients(P(coef), P(nz), P(parts), P(out_h), P(size_h), P(ctx_h), P(probs_h), mbh, mbw, nparts, step)

    # GPU statistics + streams, host bool coder
    dev = lambda a: torch.from_numpy(a).cuda()
    D = lambda t: ctypes.c_void_p(t.data_ptr())
    d_coef, d_nz, d_parts = dev(coef.reshape(-1)), dev(nz), dev(parts)
    d_probs = dev(np.full(8 * 1056, 7, np.uint32).view(np.int32))
    d_den = dev(np.full(8 * 1056, 7, np.uint32).view(np.int32))
    d_ctx = dev(np.full(M * 25, 9, np.uint8))
    cap = M * 400 * 20  # synthetic dense cases: up to 19 decisions per coefficient
    d_tok = torch.zeros(cap, dtype=torch.int16, device="cuda")
    d_mbt = torch.zeros(M, dtype=torch.int32, device="cuda")
    d_mbo = torch.zeros(M, dtype=torch.int32, device="cuda")
    d_info = torch.zeros(32, dtype=torch.int32, device="cuda")
    d_tail = torch.zeros(8 * 68, dtype=torch.int32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.vp8b200_entropy_tokens(st, D(d_coef), D(d_nz), D(d_parts), mbw, mbh, nparts, D(d_probs), D(d_den), D(d_ctx), D(d_tok),
                                  ctypes.c_uint32(cap), D(d_mbt), D(d_mbo), D(d_info), D(d_tail))
    assert rc == 0
    torch.cuda.synchronize()
    probs_g = d_probs.cpu().numpy().view(np.uint32).copy()
    den_g = d_den.cpu().numpy().view(np.uint32).copy()
    n = nparts * 1056
    assert np.array_equal(probs_g[:n], stats_h[:n]), "zero counts differ"
    assert np.array_equal(den_g[:n], den_h[:n]), "decision counts differ"
    assert np.array_equal(probs_g[n:], stats_h[n:]) and np.array_equal(den_g[n:], den_h[n:])  # untouched beyond P tables
    assert np.array_equal(d_ctx.cpu().numpy(), ctx_h), "contexts differ"
    info = d_info.cpu().numpy().view(np.uint32).copy()
    total = int(info[2 * nparts])
    assert total <= cap and total == int(d_mbt.sum().item())
    tokens = d_tok.cpu().numpy().view(np.uint16).copy()
    H.vp8b200_host_num_div_denom(P(probs_g), P(den_g), nparts)
    assert np.array_equal(probs_g[:1056], probs_h[:1056])
    out_g = np.zeros(step * nparts + 64, np.uint8)
    size_g = np.zeros(8, np.int32)
    H.vp8b200_host_encode_token_streams(P(tokens), P(info), P(probs_g), P(out_g), P(size_g), nparts, step)
    assert np.array_equal(size_g, size_h), (size_g, size_h)
    assert np.array_equal(out_g, out_h), "partition bytes differ"

    # the bool coder on the GPU as well (parallel formulation): same bytes, same sizes
    d_table = dev(probs_g.view(np.int32))
    d_out = torch.zeros(step * nparts + 64, dtype=torch.uint8, device="cuda")
    d_size = torch.zeros(8, dtype=torch.int32, device="cuda")
    L.vp8b200_entropy_boolcode_scratch_bytes.restype = ctypes.c_size_t
    need = L.vp8b200_entropy_boolcode_scratch_bytes(ctypes.c_uint32(total), nparts, step)
    d_scratch = torch.empty(need, dtype=torch.uint8, device="cuda")
    rc = L.vp8b200_entropy_boolcode(st, D(d_tok), D(d_info), D(d_table), D(d_out), D(d_size), nparts, step,
                                    ctypes.c_uint32(total), D(d_scratch))
    assert rc == 0
    torch.cuda.synchronize()
    size_d = d_size.cpu().numpy()
    assert np.array_equal(size_d, size_h), (size_d, size_h)
    out_d = d_out.cpu().numpy()
    for p in range(nparts):  # (bytes past a partition's size are scratch on both sides)
        assert np.array_equal(out_d[p * step:p * step + size_h[p]], out_h[p * step:p * step + size_h[p]]), "partition %d" % p


def test_token_stream_overflow_is_reported():
    """streams that do not fit the scratch are not written; the total says so (the shim then codes on the host)"""
    import torch
    from vp8oclenc_b200 import host as eng
    L = eng.lib()
    mbw, mbh, nparts = 11, 9, 4
    M = mbw * mbh
    coef, parts, nz = make_case(77, mbw, mbh, 1.0, True)
    dev = lambda a: torch.from_numpy(a).cuda()
    D = lambda t: ctypes.c_void_p(t.data_ptr())
    cap = 64
    d_tok = torch.full((cap + 1024,), 0x5a5a, dtype=torch.int16, device="cuda")
    d_info = torch.zeros(32, dtype=torch.int32, device="cuda")
    args = [dev(coef.reshape(-1)), dev(nz), dev(parts)]
    outs = [torch.zeros(8 * 1056, dtype=torch.int32, device="cuda"), torch.zeros(8 * 1056, dtype=torch.int32, device="cuda"),
            torch.zeros(M * 25, dtype=torch.uint8, device="cuda")]
    scratch = [torch.zeros(M, dtype=torch.int32, device="cuda"), torch.zeros(M, dtype=torch.int32, device="cuda")]
    tail = torch.zeros(8 * 68, dtype=torch.int32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.vp8b200_entropy_tokens(st, D(args[0]), D(args[1]), D(args[2]), mbw, mbh, nparts, D(outs[0]), D(outs[1]), D(outs[2]),
                                  D(d_tok), ctypes.c_uint32(cap), D(scratch[0]), D(scratch[1]), D(d_info), D(tail))
    assert rc == 0
    torch.cuda.synchronize()
    assert int(d_info.cpu().numpy().view(np.uint32)[2 * nparts]) > cap
    assert bool((d_tok == 0x5a5a).all())


@pytest.mark.parametrize("seed,counts,mode", [(1, [0, 5, 1000], "random"), (2, [0, 0, 0, 0], "random"),
                                              (3, [127, 128, 129, 255, 256, 257, 1, 4097], "likely"),
                                              (4, [9000, 0, 33000], "carries"), (5, [70000], "likely"),
                                              (6, [3000, 40, 700], "nomerge")])
def test_boolcode_synthetic_streams(seed, counts, mode):
    """vp8b200_entropy_boolcode on hand-made decision streams -- empty partitions, chunk-boundary lengths, long
    runs of 0xff bytes that carries ripple through, more than one round of the chain kernel, trajectories that never
    merge -- against the host bool coder"""
    import torch
    from vp8oclenc_b200 import host as eng
    L = eng.lib()
    H = ctypes.CDLL(SHIM)
    r = np.random.default_rng(seed)
    P_ = len(counts)
    toks = []
    for n in counts:
        if mode == "random":
            prob, bit = r.integers(1, 256, size=n), r.integers(0, 2, size=n)
        elif mode == "likely":
            prob = r.integers(1, 256, size=n)
            bit = (r.random(n) > prob / 256.0).astype(np.int64)
        elif mode == "nomerge":  # the range only ever steps down by one: the 128 trajectories never fall together
            prob, bit = np.full(n, 1), np.ones(n, np.int64)
        else:
            prob, bit = np.full(n, 255), (r.random(n) < 0.97).astype(np.int64)
        toks.append(((bit.astype(np.uint32) << 15) | (1056 + prob.astype(np.uint32))).astype(np.uint16))
    tokens = np.concatenate(toks + [np.zeros(8, np.uint16)])
    total = int(sum(counts))
    info = np.zeros(32, np.uint32)
    info[:P_] = np.concatenate([[0], np.cumsum(counts)[:-1]])
    info[P_:2 * P_] = counts
    info[2 * P_] = total
    step = 2 * max(counts) + 64
    probs = np.zeros(1056, np.uint32)
    pp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    out_h = np.zeros(step * P_ + 64, np.uint8)
    size_h = np.zeros(8, np.int32)
    H.vp8b200_host_encode_token_streams(pp(tokens), pp(info), pp(probs), pp(out_h), pp(size_h), P_, step)

    dev = lambda a: torch.from_numpy(a).cuda()
    D = lambda t: ctypes.c_void_p(t.data_ptr())
    d_tok, d_info, d_probs = dev(tokens.view(np.int16)), dev(info.view(np.int32)), dev(probs.view(np.int32))
    d_out = torch.zeros(step * P_ + 64, dtype=torch.uint8, device="cuda")
    d_size = torch.zeros(8, dtype=torch.int32, device="cuda")
    L.vp8b200_entropy_boolcode_scratch_bytes.restype = ctypes.c_size_t
    need = L.vp8b200_entropy_boolcode_scratch_bytes(ctypes.c_uint32(total), P_, step)
    d_scratch = torch.empty(need, dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.vp8b200_entropy_boolcode(st, D(d_tok), D(d_info), D(d_probs), D(d_out), D(d_size), P_, step,
                                      ctypes.c_uint32(total), D(d_scratch)) == 0
    torch.cuda.synchronize()
    size_d = d_size.cpu().numpy()
    assert np.array_equal(size_d[:P_], size_h[:P_]), (size_d, size_h)
    out_d = d_out.cpu().numpy()
    for p in range(P_):
        assert np.array_equal(out_d[p * step:p * step + size_h[p]], out_h[p * step:p * step + size_h[p]]), "partition %d" % p
