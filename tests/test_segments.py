"""Multi-GPU row of the scope table: the path shards only by independent
keyframe-delimited segments, gathered by host-side IVF concatenation, no collective.

CPU part (runs everywhere): the segment driver + IVF gather of vp8oclenc_b200/segments.py with the
reference-on-CPU runtime -- encoding a clip in one process with -g L must give the same bytes as
encoding its L-frame segments independently and concatenating them.  A world_size-2 gloo test
covers the rank -> segment assignment used by bench.py under torchrun.
GPU part: the same identity through the CUDA shim, segments running concurrently on one GPU.
"""
import os
import sys

import pytest

import _trace
from _libs import ROOT, ref

sys.path.insert(0, os.path.join(ROOT, "tools"))

W, H, SEG, NSEG = 176, 144, 6, 3
ARGS = ["-qmin", 28, "-qmax", 28, "-g", SEG, "-altref-range", 3, "-partitions", 2, "-threads", 4]


def _make(tmp):
    import gen_y4m
    y4m = os.path.join(tmp, "clip.y4m")
    gen_y4m.write_y4m(y4m, W, H, SEG * NSEG)
    return y4m


def _check(tmp, lib_dir, host_bin, per_device):
    from vp8oclenc_b200 import segments
    y4m = _make(tmp)
    whole = os.path.join(tmp, "whole.ivf")
    p = segments.EncoderProcess(y4m, whole, ARGS, os.path.join(tmp, "whole_run"), lib_dir, host_bin)
    stamps = p.wait()
    assert len(stamps) == SEG * NSEG
    segs = segments.split_y4m(y4m, SEG, tmp)
    assert len(segs) == NSEG
    ivfs, procs = segments.encode_segments(segs, tmp, ARGS, devices=(0,), per_device=per_device, lib_dir=lib_dir,
                                           host_bin=host_bin)
    joined = os.path.join(tmp, "joined.ivf")
    n = segments.concat_ivf(ivfs, joined)
    assert n == SEG * NSEG
    a, b = open(whole, "rb").read(), open(joined, "rb").read()
    assert a == b, "segment-parallel output differs from the serial encode"
    # every segment starts with a key frame (bit 0 of the first payload byte is 0 for key frames)
    _, frames = segments.read_ivf(joined)
    for s in range(NSEG):
        assert frames[s * SEG][1][0] & 1 == 0
        assert frames[s * SEG + 1][1][0] & 1 == 1


@pytest.mark.skipif(ref() is None or not _trace.have_host(), reason="oracle/_ref not built (needs /root/reference)")
def test_segments_equal_serial_encode_reference_runtime(tmp_path):
    _check(str(tmp_path), _trace.REF_DIR, _trace.HOST_BIN, per_device=2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    # the assignment bench.py uses: rank r encodes frames [r*n, (r+1)*n) of the clip; the only
    # communication is the barrier and the MAX over ranks of the elapsed time
    n = 7
    mine = list(range(rank * n, (rank + 1) * n))
    t = torch.tensor([float(10 + rank)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, float(t.item()), gathered))
    dist.destroy_process_group()


def test_rank_to_segment_assignment_gloo_world2():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
    for rank, tmax, gathered in res:
        assert tmax == 11.0  # max over ranks
        flat = [f for seg in gathered for f in seg]
        assert flat == list(range(14))  # disjoint, contiguous segments covering the clip


@pytest.mark.gpu
def test_segments_equal_serial_encode_cuda_shim(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vp8oclenc_b200 import segments
    if not os.path.exists(segments.HOST_BIN):
        pytest.skip("host binary not built")
    _check(str(tmp_path), segments.SHIM_DIR, segments.HOST_BIN, per_device=3)
