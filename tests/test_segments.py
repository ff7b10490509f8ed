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


CUT_FRAMES, CUT_AT, CUT_GOP = 30, 7, 10
CUT_ARGS = ["-qmin", 28, "-qmax", 28, "-g", CUT_GOP, "-altref-range", 3, "-partitions", 2, "-threads", 4]


def _scene_cut_clip(tmp):
    """chroma jumps at frame CUT_AT: the serial encoder forces a key frame there and counts its GOP from it
    (keys at 0, 7, 17, 27 -- not at 0, 10, 20)"""
    from test_gpu_e2e import _write_scene_cut_clip
    y4m = os.path.join(tmp, "cut.y4m")
    _write_scene_cut_clip(y4m, W, H, CUT_FRAMES, CUT_AT)
    return y4m


def _serial(tmp, y4m, lib_dir, host_bin, args):
    from vp8oclenc_b200 import segments
    whole = os.path.join(tmp, "whole.ivf")
    segments.EncoderProcess(y4m, whole, args, os.path.join(tmp, "whole_run"), lib_dir, host_bin).wait()
    return whole


def _check_planned(tmp, lib_dir, host_bin, per_device):
    from vp8oclenc_b200 import segments
    y4m = _scene_cut_clip(tmp)
    whole = _serial(tmp, y4m, lib_dir, host_bin, CUT_ARGS)
    keys, exact = segments.plan_key_frames(y4m, CUT_GOP)
    assert exact and keys == [0, CUT_AT, CUT_AT + CUT_GOP, CUT_AT + 2 * CUT_GOP]
    # ... which is where the serial encode has its key frames
    _, frames = segments.read_ivf(whole)
    assert [i for i, (_, p) in enumerate(frames) if p[0] & 1 == 0] == keys
    n, ivfs, exact, _ = segments.encode_clip_segment_parallel(y4m, CUT_ARGS, CUT_GOP, os.path.join(tmp, "par"), device=0,
                                                              per_device=per_device, lib_dir=lib_dir, host_bin=host_bin)
    assert n == len(keys) and sorted(ivfs) == list(range(n))
    joined = os.path.join(tmp, "joined.ivf")
    assert segments.concat_ivf([ivfs[i] for i in range(n)], joined) == CUT_FRAMES
    assert open(whole, "rb").read() == open(joined, "rb").read(), "planned cuts: concatenation differs from the serial encode"
    # cutting blindly at multiples of -g does NOT reproduce the serial stream on this clip (extra key frames)
    naive = segments.split_y4m(y4m, CUT_GOP, tmp, prefix="naive")
    nivfs, _ = segments.encode_segments(naive, os.path.join(tmp, "naive_out"), CUT_ARGS, devices=(0,), per_device=per_device,
                                        lib_dir=lib_dir, host_bin=host_bin)
    njoined = os.path.join(tmp, "naive_joined.ivf")
    assert segments.concat_ivf(nivfs, njoined) == CUT_FRAMES
    assert open(whole, "rb").read() != open(njoined, "rb").read()


@pytest.mark.skipif(ref() is None or not _trace.have_host(), reason="oracle/_ref not built (needs /root/reference)")
def test_planned_cuts_follow_forced_key_frames_reference_runtime(tmp_path):
    os.makedirs(os.path.join(str(tmp_path), "naive_out"), exist_ok=True)
    _check_planned(str(tmp_path), _trace.REF_DIR, _trace.HOST_BIN, per_device=2)


def test_assign_segments_partitions_the_clip():
    from vp8oclenc_b200 import segments
    for n in (1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            got = sorted(i for r in range(world) for i in segments.assign_segments(n, world, r))
            assert got == list(range(n))
            sizes = [len(segments.assign_segments(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, tmp, q):
    """one rank of a world-size-2 segment-parallel encode: plan, encode the own share (segments.assign_segments),
    barrier, rank 0 gathers by concatenation -- exactly what bench.py does under torchrun, on the CPU runtime"""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from vp8oclenc_b200 import segments
    y4m = os.path.join(tmp, "cut.y4m")
    n, ivfs, exact, _ = segments.encode_clip_segment_parallel(y4m, CUT_ARGS, CUT_GOP, os.path.join(tmp, "par"), rank=rank,
                                                              world=world, device=0, per_device=1, lib_dir=_trace.REF_DIR,
                                                              host_bin=_trace.HOST_BIN)
    t = torch.tensor([float(10 + rank)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the timing reduction of bench.py
    gathered = [None] * world
    dist.all_gather_object(gathered, ivfs)
    same = None
    if rank == 0:
        paths = {}
        for g in gathered:
            paths.update(g)
        joined = os.path.join(tmp, "joined.ivf")
        segments.concat_ivf([paths[i] for i in range(n)], joined)
        same = open(joined, "rb").read() == open(os.path.join(tmp, "whole.ivf"), "rb").read()
    q.put((rank, float(t.item()), sorted(ivfs), n, same))
    dist.destroy_process_group()


@pytest.mark.skipif(ref() is None or not _trace.have_host(), reason="oracle/_ref not built (needs /root/reference)")
def test_segment_parallel_encode_gloo_world2(tmp_path):
    """N > 1 path on CPU: two ranks (gloo) encode the segments segments.assign_segments() gives them with the
    reference-on-CPU runtime; rank 0's concatenation equals the serial encode byte for byte"""
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    tmp = str(tmp_path)
    y4m = _scene_cut_clip(tmp)
    _serial(tmp, y4m, _trace.REF_DIR, _trace.HOST_BIN, CUT_ARGS)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, tmp, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(60)
    assert [r[2] for r in res] == [[0, 2], [1, 3]]  # round-robin shares of the four segments
    for rank, tmax, mine, n, same in res:
        assert tmax == 11.0 and n == 4
    assert res[0][4] is True


@pytest.mark.gpu
def test_planned_cuts_follow_forced_key_frames_cuda_shim(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vp8oclenc_b200 import segments
    if not os.path.exists(segments.HOST_BIN):
        pytest.skip("host binary not built")
    os.makedirs(os.path.join(str(tmp_path), "naive_out"), exist_ok=True)
    _check_planned(str(tmp_path), segments.SHIM_DIR, segments.HOST_BIN, per_device=2)


@pytest.mark.gpu
def test_segments_equal_serial_encode_cuda_shim(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vp8oclenc_b200 import segments
    if not os.path.exists(segments.HOST_BIN):
        pytest.skip("host binary not built")
    _check(str(tmp_path), segments.SHIM_DIR, segments.HOST_BIN, per_device=3)


def test_reference_stage_split_of_the_cpu_baseline(tmp_path):
    """bench.py's per-stage split of the reference on the CPU: the per-kernel wall times the
    reference runtime writes (VP8CL_STAGES) are grouped into motion search / predict + transform / loop filter /
    entropy, search and transform averaged over the inter frames only, and what is left of a frame is the host"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    f = tmp_path / "stages.txt"
    f.write_text("total_wall 1000.0\nluma_search_1step 50 300.0\nluma_search_2step 10 500.0\ndct4x4 40 40.0\n"
                 "loop_filter_frame_luma 5 55.0\nencode_coefficients 5 22.0\ncount_probs 5 11.0\nunknown_kernel 1 9.0\n")
    stamps = [0.0, 0.25, 0.5, 0.75, 1.0]  # 1 key + 4 inter frames, 250 ms each
    st = bench.reference_stages(str(f), stamps)
    assert st["motion_search"] == pytest.approx(800.0 / 4)
    assert st["predict_transform"] == pytest.approx(40.0 / 4)
    assert st["loop_filter"] == pytest.approx(55.0 / 5)
    assert st["entropy"] == pytest.approx(33.0 / 5)
    assert st["frame"] == pytest.approx(250.0)
    assert st["host_rest"] == pytest.approx(250.0 - 200.0 - 10.0 - 11.0 - 6.6, abs=1e-3)
    assert bench.reference_stages(str(tmp_path / "missing.txt"), stamps) is None
