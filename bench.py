#!/usr/bin/env python3
"""bench.py -- encoded frames/s of vp8oclenc's inter-frame hot path on B200 (BASELINE.json).

A "step" is one pass of the hot path over one batch of synthetic input: FRAMES_PER_STEP consecutive inter frames of
each of the `segments_per_gpu` independent keyframe-delimited segments resident on the GPU.  One frame = pyramid +
hierarchical motion search over LAST / GOLDEN / ALTREF, reference selection, six-tap prediction, DCT/WHT/quantise,
dequantise/reconstruct, SSIM, filter mask and the normal loop filter (tools/gen_y4m.py makes the clip).

  value     frames/s of the CUDA engine with every input frame already resident in HBM
            (vp8b200_engine_inter_frame + vp8b200_engine_loop_filter), CUDA events, L2 flushed between timed steps
  e2e       frames/s of the UNMODIFIED reference host program running against our OpenCL shim
            (vp8oclenc_b200/lib/libOpenCL.so.1) on the same clip: Y4M in, IVF out, every host<->device copy, the
            host's intra/entropy/bitstream work and file I/O included.  The .ivf files are checked: a prefix of one
            of them against the reference encoder's own output for the same frames, instances that encode the same
            clip against each other; a mismatch voids the number.
  roofline  the dominant kernel (quarter-pel motion search) against the measured integer issue rate of the device;
            `kernels` carries one entry per stage of the frame (five full-pel search levels, predict/transform/SSIM
            against HBM, loop filter against its dependency chain), timed live inside a frame on the engine's stream
  sizes     the same measurement at 3840x2160 (BASELINE configs[2]), reduced in length
  cpu_baseline / --impl reference
            the reference itself (its own host + its own .cl kernels compiled for the CPU, oracle/_ref) on a bounded
            sample of the same clip on all host cores

  --config 4 | 5   BASELINE configs[3] / configs[4] as written: one long clip cut at its key frames into segments,
            the segments encoded in parallel over all ranks, concatenated, and compared with a serial encode.

Multi-GPU (torchrun, one rank per GPU): independent keyframe-delimited segments per GPU, no collective on the data
path ("scaling": "weak").
"""
import argparse
import ctypes
import hashlib
import json
import os
import shutil
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

ENC_ARGS = ["-qmin", "24", "-qmax", "24", "-g", "150", "-altref-range", "5", "-partitions", "8", "-threads", "12"]
ALTREF_RANGE = 5
QI = (24, 24, 24, 24)
REF_STEP_CAP = 60      # the reference arm honours --steps up to this many timed frames (about 5 frames/s at 1080p)


def padded(w, h):
    return (w + 15) // 16 * 16, (h + 15) // 16 * 16


def size_name(w, h):
    return {(1920, 1080): "1080p", (3840, 2160): "2160p", (7680, 4320): "4320p", (352, 288): "CIF"}.get((w, h), "%dx%d" % (w, h))


def make_config(w, h, segments, frames_per_step):
    ww, wh = padded(w, h)
    return {"workload": "%dx%d synthetic YUV420 (tools/gen_y4m.py), padded to %dx%d, LAST+GOLDEN+ALTREF, q=24, altref-range 5, "
                        "8 partitions, loop filter on the GPU" % (w, h, ww, wh),
            "frame_size": [w, h], "segments_per_gpu": segments, "frames_per_segment_per_step": frames_per_step,
            "step": "%d consecutive inter frames of each of the segments_per_gpu independent segments (one engine + stream "
                    "per segment)" % frames_per_step,
            "l2": "flushed between timed steps (256 MiB fill outside the timed span)"}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), polled through NVML every few
    milliseconds (the timed region of the default run is shorter than one nvidia-smi sampling period)"""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index=0):
        self.gpu_index, self.samples, self.reasons, self.stop_flag, self.thread = gpu_index, [], set(), False, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu_index])
                except (ValueError, IndexError):
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return

        def poll():
            while not self.stop_flag:
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.004)
        self.thread = threading.Thread(target=poll, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join()
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def reference_encode(tmp, w, h, frames, tag, threads=None):
    """the reference's own CPU implementation (its host + its .cl kernels compiled for the CPU, oracle/_ref, OpenMP over
    work-items) on the first `frames` frames of the clip.  -> (per-frame completion times, ivf path) or None"""
    import gen_y4m
    from vp8oclenc_b200 import segments
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    host_bin = os.path.join(ref_dir, "vp8enc")
    if not os.path.exists(host_bin) or not os.path.exists(os.path.join(ref_dir, "libOpenCL.so.1")):
        return None
    y4m = os.path.join(tmp, "ref_%s.y4m" % tag)
    gen_y4m.write_y4m(y4m, w, h, frames)
    ivf = os.path.join(tmp, "ref_%s.ivf" % tag)
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
    stage_file = os.path.join(tmp, "ref_stages_%s.txt" % tag)
    pr = segments.EncoderProcess(y4m, ivf, ENC_ARGS, os.path.join(tmp, "ref_run_%s" % tag), lib_dir=ref_dir, host_bin=host_bin,
                                 env_extra={"OMP_NUM_THREADS": str(threads or os.cpu_count() or 1), "VP8CL_STAGES": stage_file})
    stamps = pr.wait(timeout=1500)
    os.remove(y4m)
    if len(stamps) != frames:
        raise RuntimeError("reference encoder finished %d of %d frames" % (len(stamps), frames))
    reference_encode.stages = reference_stages(stage_file, stamps)
    return stamps, ivf


REF_STAGES = {  # BASELINE.md's per-stage split: the reference's kernels by name
    "motion_search": ("reset_vectors", "downsample_x2", "luma_search_1step", "luma_search_2step", "select_reference", "pack_8x8_into_16x16"),
    "predict_transform": ("prepare_predictors_and_residual", "dct4x4", "wht4x4_iwht4x4", "idct4x4", "count_SSIM_luma", "count_SSIM_chroma",
                          "gather_SSIM"),
    "loop_filter": ("prepare_filter_mask", "loop_filter_frame_luma", "loop_filter_frame_chroma"),
    "entropy": ("count_probs", "num_div_denom", "encode_coefficients"),
}


def reference_stages(stage_file, stamps):
    """ms per inter frame of the reference's stages on the CPU, from the runtime's per-kernel wall times
    (oracle/cl_host_runtime.cpp, VP8CL_STAGES).  Search and transform only run in inter frames; loop filter and entropy
    run in every frame and are averaged over all of them; host = the inter frames' wall time minus the four."""
    if not os.path.exists(stage_file) or len(stamps) < 2:
        return None
    ms = {}
    for ln in open(stage_file):
        f = ln.split()
        if len(f) == 3:
            ms[f[0]] = float(f[2])
    frames, inter = len(stamps), len(stamps) - 1
    out = {}
    for stage, names in REF_STAGES.items():
        tot = sum(ms.get(n, 0.0) for n in names)
        out[stage] = tot / (inter if stage in ("motion_search", "predict_transform") else frames)
    per_inter = 1000.0 * (stamps[-1] - stamps[0]) / inter
    out["host_rest"] = max(0.0, per_inter - sum(out.values()))  # (the key frames' denser entropy stage is in the average)
    out["frame"] = per_inter
    return {k: round(v, 3) for k, v in out.items()}


def reference_arm(args, tmp, w, h):
    warm = max(1, args.warmup)
    steps = max(1, min(args.steps, REF_STEP_CAP))
    r = reference_encode(tmp, w, h, 1 + warm + steps, "arm")
    if r is None:
        return None
    stamps, ivf = r
    dt = stamps[-1] - stamps[warm]  # frame 0 is the key frame, then `warm` untimed inter frames
    return {"value": steps / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference",
            "sample": "%d inter frames of the %s clip after 1 key + %d warm-up frames; unmodified reference host with its own "
                      ".cl kernels compiled for the CPU (oracle/_ref), OpenMP over work-items, all host cores"
                      % (steps, size_name(w, h), warm),
            "ms_per_step": 1000.0 * dt / steps, "steps": steps, "warmup": warm, "ivf": ivf,
            "stages_ms_per_frame": getattr(reference_encode, "stages", None)}


# ------------------------------------------------------------------------------------------------
def device_pipeline(w, h, S, F, K, W, rank, world, local_rank, dist, sample_clocks):
    """`value`: S engines (one stream each) with their frames resident in HBM; a step = F frames of every segment.
    -> dict(value, ms_per_step, launches, refs_per_frame, clocks, last frames of segment 0, engines' stream of segment 0)"""
    import numpy as np
    import torch
    import gen_y4m
    from vp8oclenc_b200 import host as eng
    from vp8oclenc_b200.hostlogic import HostState, make_segment_data
    ww, wh = padded(w, h)
    nframes = 1 + (W + K) * F
    clip = gen_y4m.Clip(ww, wh)
    dev_frames = []
    for sgm in range(S):
        first = (rank * S + sgm) * nframes
        dev_frames.append([[torch.from_numpy(np.ascontiguousarray(p)).cuda() for p in clip.frame(first + i)] for i in range(nframes)])
    sd = make_segment_data(QI)
    engines = [eng.Engine(ww, wh) for _ in range(S)]
    ext = [torch.cuda.ExternalStream(x.stream) for x in engines]
    master = torch.cuda.Stream()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    torch.cuda.synchronize()
    states = [HostState(10 ** 6, ALTREF_RANGE) for _ in range(S)]
    for sgm in range(S):
        states[sgm].next_frame()
        engines[sgm].set_reconstruction(*dev_frames[sgm][0])  # the key frame's reconstruction seeds LAST/GOLDEN/ALTREF

    def frame(sgm, i):
        st = states[sgm].next_frame()
        y, u, v = dev_frames[sgm][i]
        x = engines[sgm]
        x.inter_frame(y, u, v, sd, -1.0, st["prev_golden"], st["prev_altref"], st["altref_differs"])
        n1 = x.last_launch_count
        x.loop_filter(None)
        return n1 + x.last_launch_count, 1 + (not st["prev_golden"]) + (not st["prev_altref"] and st["altref_differs"])

    nxt = 1
    for _ in range(W):
        for f in range(F):
            for sgm in range(S):
                frame(sgm, nxt + f)
        nxt += F
    for x in engines:
        x.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if sample_clocks:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    done = [[torch.cuda.Event() for _ in range(S)] for _ in range(K)]
    launches = refs = 0
    for k in range(K):
        with torch.cuda.stream(master):
            flush_buf.fill_(k & 255)           # L2 flush between timed steps, outside the timed span
            ev[k][0].record(master)
        for sgm in range(S):
            ext[sgm].wait_event(ev[k][0])
        for f in range(F):
            for sgm in range(S):
                n, r = frame(sgm, nxt + f)
                launches += n
                refs += r
        nxt += F
        for sgm in range(S):
            done[k][sgm].record(ext[sgm])
            master.wait_event(done[k][sgm])
        ev[k][1].record(master)
    for x in engines:
        x.synchronize()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if sample_clocks else None
    if dist:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    return {"value": world * K * S * F / (total_ms * 1e-3), "ms_per_step": total_ms / K, "launches": launches,
            "refs_per_frame": refs / float(K * S * F), "clocks": clocks, "engines": engines, "frames": dev_frames[0],
            "flush": flush_buf, "sd": sd, "states": states}


def kernel_lines(w, h, pipe, units):
    """per-stage live timings of ONE frame that searches all three references, on one engine alone (L2 flushed before
    the frame), and the roofline each stage is held against.  Rank 0 only."""
    import torch
    from vp8oclenc_b200 import host as eng
    ww, wh = padded(w, h)
    N, M = ww * wh, (ww // 16) * (wh // 16)
    e = pipe["engines"][0]
    L = eng.lib()
    L.vp8b200_measure_int_ops_per_second.restype = ctypes.c_double
    int_peak = L.vp8b200_measure_int_ops_per_second(ctypes.c_void_p(e.stream), 5)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    ext = torch.cuda.ExternalStream(e.stream)
    frames, sd, flush = pipe["frames"], pipe["sd"], pipe["flush"]
    e.stage_timing(True)
    acc, reps = {}, 0
    # frames whose GOLDEN and ALTREF both differ from LAST: replay the last frames of the segment with those flags
    for i in range(max(1, len(frames) - 6), len(frames)):
        with torch.cuda.stream(ext):
            flush.fill_(3)
        y, u, v = frames[i]
        e.inter_frame(y, u, v, sd, -1.0, 0, 0, 1)
        e.loop_filter(None)
        for k, ms in e.stage_times().items():
            acc[k] = acc.get(k, 0.0) + ms
        reps += 1
    e.stage_timing(False)
    t = {k: v / reps for k, v in acc.items()}
    refs = 3
    out = {}

    def alu(name, ms, level_pixels, keys):
        u = units.get(name, {})
        per_px_exec = u.get("thread_instr_per_pixel_per_ref")
        per_px_min = u.get("min_ops_per_pixel_per_ref")
        ent = {"bound": "int_alu", "ms_per_launch": ms, "references_per_launch": refs, "peak": int_peak / 1e12, "unit": "Tiop/s"}
        if per_px_min:
            ent["achieved"] = per_px_min * level_pixels * refs / (ms * 1e-3) / 1e12
            ent["frac"] = ent["achieved"] * 1e12 / int_peak
            ent["min_ops_per_pixel_per_ref"] = per_px_min
        if per_px_exec:
            ent["issue_utilisation"] = per_px_exec * level_pixels * refs / (ms * 1e-3) / int_peak
            ent["executed_thread_instr_per_pixel_per_ref"] = per_px_exec
        for k in keys:
            if k in u:
                ent[k] = u[k]
        return ent

    if "search_qpel" in t:
        out["luma_search_2step"] = alu("luma_search_2step", t["search_qpel"], N, ("alu_pipe_thread_instr_per_pixel_per_ref", "dram_bytes_per_ref"))
        out["luma_search_2step"]["reference_formulation_ops_per_pixel_per_ref"] = 1381.0
    for lvl, key in ((1, "search_1x"), (2, "search_2x"), (4, "search_4x"), (8, "search_8x"), (16, "search_16x")):
        if key in t:
            ent = alu("luma_search_1step", t[key], N // (lvl * lvl), ())
            ent["level"] = "1/%d" % lvl if lvl > 1 else "1"
            out["luma_search_1step_%dx" % lvl] = ent
    if "transform" in t:
        b = 7.7 * N  # the fused predict/transform/reconstruct/SSIM path reads cur + ref and writes recon + coefficients
        out["mb_predict_transform_fused"] = {"bound": "hbm", "ms_per_launch": t["transform"], "achieved": b / (t["transform"] * 1e-3) / 1e9,
                                             "peak": hbm_peak, "unit": "GB/s", "frac": b / (t["transform"] * 1e-3) / 1e9 / hbm_peak,
                                             "algorithmic_bytes_per_pixel": 7.7, "peak_source": peak_src,
                                             "note": "issue-bound at this size (float SSIM chains, quantiser ladder): see DESIGN.md"}
    if "loop_filter" in t:
        b = 3.0 * N
        stages = ww // 16 + 2 * (wh // 16 - 1)
        out["loop_filter_planes"] = {"bound": "hbm", "ms_per_launch": t["loop_filter"], "achieved": b / (t["loop_filter"] * 1e-3) / 1e9,
                                     "peak": hbm_peak, "unit": "GB/s", "frac": b / (t["loop_filter"] * 1e-3) / 1e9 / hbm_peak,
                                     "algorithmic_bytes_per_pixel": 3.0, "peak_source": peak_src, "dependent_stages": stages,
                                     "us_per_dependent_stage": 1000.0 * t["loop_filter"] / stages,
                                     "note": "dependency-bound wavefront of mb_w+2(mb_h-1) macroblock stages, not bandwidth-bound"}
    out["frame_stage_ms"] = t
    out["key_frame"] = key_frame_line(w, h, pipe)
    return out, int_peak


def key_frame_line(w, h, pipe):
    """one key frame -- intra_transform() + loop filter -- on the engine (vp8b200_engine_key_frame, CUDA
    events on the engine's stream) next to the reference's own intra code (src/intra_part.h compiled in place,
    oracle/_ref/libref_intra.so: the cpu_baseline of this stage, one host thread as in the reference)"""
    import numpy as np
    import torch
    from vp8oclenc_b200.hostlogic import intra_quants, make_segment_data
    ww, wh = padded(w, h)
    e = pipe["engines"][0]
    ext = torch.cuda.ExternalStream(e.stream)
    y, u, v = pipe["frames"][0]
    sdk = make_segment_data(QI, key=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps, t_intra, t_lf = 5, 0.0, 0.0
    for i in range(reps + 1):
        with torch.cuda.stream(ext):
            pipe["flush"].fill_(4)
            ev[0].record(ext)
        e.key_frame(y, u, v, sdk)
        with torch.cuda.stream(ext):
            ev[1].record(ext)
        e.loop_filter(sdk)
        with torch.cuda.stream(ext):
            ev[2].record(ext)
        e.synchronize()
        if i:  # (first pass: warm-up)
            t_intra += ev[0].elapsed_time(ev[1])
            t_lf += ev[1].elapsed_time(ev[2])
    line = {"gpu_intra_ms": t_intra / reps, "gpu_loop_filter_ms": t_lf / reps, "dependent_stages": ww // 16 + 2 * (wh // 16 - 1),
            "us_per_dependent_stage": 1000.0 * t_intra / reps / (ww // 16 + 2 * (wh // 16 - 1)),
            "what": "vp8b200_engine_key_frame + vp8b200_engine_loop_filter: B_PRED mode decision over the ten sub-block modes, TM chroma, "
                    "transform, quantise, reconstruct for every macroblock (wavefront over macroblocks), then the normal loop filter"}
    so = os.path.join(ROOT, "oracle", "_ref", "libref_intra.so")
    if os.path.exists(so):
        lib = ctypes.CDLL(so)
        M = (ww // 16) * (wh // 16)
        hy, hu, hv = (np.ascontiguousarray(t.cpu().numpy()) for t in (y, u, v))
        ry, ru, rv = np.zeros(ww * wh, np.uint8), np.zeros(ww * wh // 4, np.uint8), np.zeros(ww * wh // 4, np.uint8)
        mb, modes = np.zeros(M * 400, np.int16), np.zeros(M * 16, np.int32)
        parts, seg = np.zeros(M, np.int32), np.zeros(M, np.int32)
        q = np.asarray(intra_quants(sdk), np.int32)
        P = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            lib.vp8ref_intra_frame(ww, wh, P(hy), P(hu), P(hv), P(ry), P(ru), P(rv), P(mb), P(modes), P(parts), P(seg), P(q))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        line["reference_host_intra_ms"] = 1000.0 * best
        line["reference_kind"] = "reference (src/intra_part.h compiled in place, one host thread as in the reference)"
    return line


# ------------------------------------------------------------------------------------------------
def end_to_end(args, w, h, Ke, W, rank, world, local_rank, dist, tmp, ref_ivf, distinct=None):
    """e2e: P instances of the unmodified reference host + shim per GPU, each on its own keyframe-delimited segment of
    1 + W + Ke frames; frames/s counts the frames all instances finish between "every instance is past its warm-up"
    and "the last instance is done".  The outputs are checked (see the module docstring)."""
    import torch
    import gen_y4m
    from vp8oclenc_b200 import segments
    if not os.path.exists(segments.HOST_BIN):
        return None
    n = 1 + W + Ke
    distinct = distinct or (8 if world == 1 else (4 if world == 2 else 2))  # clips per rank (they live in /dev/shm)
    clips = []
    gate_root = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()

    def run(P, tag, env_more=None):
        paths, outs = [], []
        for p in range(P):
            if p % distinct >= len(clips):
                y4m = os.path.join(tmp, "e2e_%d_%d.y4m" % (rank, p % distinct))
                # (rank 0's first clip starts at frame 0 of the sequence: the frames the reference encoder is run on)
                gen_y4m.write_y4m(y4m, w, h, n, start=(rank * distinct + p) * n)
                clips.append(y4m)
            paths.append(clips[p % distinct])
            outs.append(os.path.join(tmp, "e2e_%s_%d_%d" % (tag, rank, p)))
        # start gate (cl_shim.cu start_gate): all P x world instances come up (context creation, module loading and
        # page pinning are serialised by the driver), encode their key frame and two inter frames, then go on
        # together; frames 3..W are the common warm-up
        gate = os.path.join(gate_root, "vp8b200_gate_%s_%s" % (os.environ.get("MASTER_PORT", os.getpid()), tag))
        if rank == 0:
            shutil.rmtree(gate, ignore_errors=True)
            os.makedirs(gate)
        if dist:
            dist.barrier()
        procs = [segments.EncoderProcess(paths[p], outs[p] + ".ivf", ENC_ARGS, os.path.join(tmp, "run_%s_%d_%d" % (tag, rank, p)),
                                         device=local_rank,
                                         env_extra=dict(env_more or {}, VP8B200_STATS=outs[p] + ".stats", VP8B200_STATS_FROM=str(W),
                                                        VP8B200_START_GATE="%s:%d:3" % (gate, P * world)))
                 for p in range(P)]
        failure = None
        try:
            stamps = [pr.wait(timeout=900) for pr in procs]
            for st in stamps:
                if len(st) != n:
                    raise RuntimeError("an encoder instance finished %d of %d frames" % (len(st), n))
        except Exception as err:  # noqa: BLE001
            failure = "rank %d: %s" % (rank, err)
            for pr in procs:
                if pr.proc.poll() is None:
                    pr.proc.kill()
        # ---- the outputs: instances that encoded the same clip must agree byte for byte; the first frames of rank 0's
        # first instance must be the reference encoder's (the encoder is causal: a prefix of the clip gives a prefix
        # of the stream)
        check = {"files": P, "distinct_clips": min(P, distinct), "duplicates_identical": None, "reference_prefix_frames": 0,
                 "reference_prefix_identical": None, "entropy_host_fallbacks": None}
        if not failure:
            try:
                digests = [hashlib.md5(open(o + ".ivf", "rb").read()).hexdigest() for o in outs]
                check["duplicates_identical"] = all(digests[p] == digests[p % distinct] for p in range(P))
                check["md5_first"] = digests[0]
                if rank == 0 and ref_ivf:
                    _, got = segments.read_ivf(outs[0] + ".ivf")
                    _, want = segments.read_ivf(ref_ivf)
                    check["reference_prefix_frames"] = len(want)
                    check["reference_prefix_identical"] = len(got) >= len(want) and all(a[1] == b[1] for a, b in zip(want, got))
                stats = [json.load(open(o + ".stats")) for o in outs]
                check["entropy_host_fallbacks"] = int(sum(s.get("entropy_host_fallbacks", 0) for s in stats))
                if not check["duplicates_identical"] or check["reference_prefix_identical"] is False or check["entropy_host_fallbacks"]:
                    failure = "rank %d: output check failed: %s" % (rank, json.dumps(check))
            except Exception as err:  # noqa: BLE001
                failure = "rank %d: output check: %s" % (rank, err)
        if dist:
            ok = torch.tensor([0.0 if failure else 1.0], device="cuda", dtype=torch.float64)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 1.0 and not failure:
                failure = "an encoder instance of another rank failed"
        if failure:
            raise RuntimeError(failure)
        t0 = max(st[W] for st in stamps)
        t1 = max(st[-1] for st in stamps)
        if dist:  # one window for all ranks (perf_counter is a per-host monotonic clock; one node)
            tt = torch.tensor([t0, t1], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t0, t1 = float(tt[0].item()), float(tt[1].item())
        count = sum(1 for st in stamps for x in st if x > t0)
        # steady-state host cost per frame from the instances' own accounting (from inter frame W on)
        fr = sum(s["window_frames"] for s in stats)
        cpu = {"core_ms_per_frame": sum(s["cpu_user_ms"] + s["cpu_sys_ms"] for s in stats) / fr,
               "host_program_ms": sum(s["cpu_ms_calling_thread"] - sum(s["cpu_ms_" + k] for k in ("launch", "host_kernel", "read", "write", "map", "finish")) for s in stats) / fr,
               "shim_ms": sum(sum(s["cpu_ms_" + k] for k in ("launch", "host_kernel", "read", "write", "map", "finish")) for s in stats) / fr,
               "of_it_waiting_ms": sum(s["cpu_ms_wait"] for s in stats) / fr,
               "waits_per_frame": sum(s["waits"] for s in stats) / fr}
        if dist:
            cc = torch.tensor([float(count)], device="cuda", dtype=torch.float64)
            dist.all_reduce(cc, op=dist.ReduceOp.SUM)
            count = int(cc.item())
        s0 = stats[0]
        f0 = s0["window_frames"]
        per = {"h2d": int(s0["h2d_bytes"] / f0), "d2h": int(s0["d2h_bytes"] / f0), "launches": s0["kernel_launches"] / f0}
        if dist:
            dist.barrier()
        if rank == 0:
            shutil.rmtree(gate, ignore_errors=True)
        for o in outs:
            for suffix in (".ivf", ".stats"):
                try:
                    os.remove(o + suffix)
                except OSError:
                    pass
        return count / (t1 - t0), per, check, cpu

    # several instances per GPU share it through a private CUDA MPS daemon (one per node, started by local rank 0);
    # without MPS the contexts are time-sliced and more than ~8 instances do not pay
    daemon = segments.MpsDaemon(os.path.join(tempfile.gettempdir(), "vp8b200_mps_%s" % os.environ.get("MASTER_PORT", os.getpid())))
    if local_rank == 0 and args.e2e_procs != 1 and not args.no_mps:
        daemon.__enter__()
    if dist:
        dist.barrier()
        daemon.active = os.path.exists(os.path.join(daemon.pipe, "control"))
    cores = os.cpu_count() or 2
    big = padded(w, h)[0] * padded(w, h)[1] > 1920 * 1088
    if daemon.active:
        # one and a half instances per host core on one or two GPUs, two per core beyond (the instances sleep while they
        # wait for the GPU: the others fill the gap), at most 32 per GPU (24 above 1080p) and 128 on the node
        per_core = 3 if world <= 2 else 4
        P = args.e2e_procs or max(1, min(24 if big else 32, (per_core * cores) // (2 * world), max(4, 128 // world)))
    else:
        P = args.e2e_procs or max(1, min(8, cores // max(2, 2 * world)))
    err_text, fps1, fpsP, per, check, cpu = None, 0.0, 0.0, None, None, None
    try:
        fps1, per, check1, _ = run(1, "single")
        check = check1
        try:
            if P == 1:
                fpsP = fps1
            else:
                fpsP, _, check, cpu = run(P, "multi", daemon.env())
                check["reference_prefix_frames"] = check1["reference_prefix_frames"]
                check["reference_prefix_identical"] = check1["reference_prefix_identical"] and check["reference_prefix_identical"] is not False
        except RuntimeError as err:
            if dist or not daemon.active:
                raise
            sys.stderr.write("bench: multi-instance run under MPS failed (%s); retrying without MPS\n" % err)
            daemon.__exit__(None, None, None)
            P = max(1, min(8, cores // 2))
            fpsP, _, check, cpu = run(P, "multi_nomps", {})
    except RuntimeError as err:
        err_text = str(err)  # (run() has made sure every rank raises together)  The line is still printed, without an e2e value.
        sys.stderr.write("bench: end-to-end run failed: %s\n" % err)
    finally:
        mps_used = daemon.active
        if dist:
            dist.barrier()
        if local_rank == 0:
            daemon.__exit__(None, None, None)
    best = max(fps1, fpsP)
    if err_text or best <= 0:
        return {"value": None, "unit": "frames/s", "error": err_text or "no frames", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None}
    return {"value": best, "unit": "frames/s", "ms_per_frame": 1000.0 / best, "processes_per_gpu": P if fpsP >= fps1 else 1,
            "single_process_fps": fps1, "multi_process_fps": fpsP, "mps": bool(mps_used and P > 1), "host_cores": cores,
            "frames_per_s_per_host_core": best / cores,
            "h2d_bytes_per_step": per["h2d"], "d2h_bytes_per_step": per["d2h"], "bytes_are": "per encoded frame (one e2e step = one frame of one instance)",
            "shim_kernel_launches_per_frame": per["launches"], "timed_frames_per_instance": Ke, "output_check": check,
            "host_cpu_per_frame": cpu, "shim_mode": "VP8B200_HOST_PROFILE=reference (lazy downloads, page-locked source planes, polling waits)",
            "window": "from the moment the last instance has finished its key + warm-up frames to the moment the last instance is done; "
                      "instances wait for each other at the start of their third inter frame (start gate, inside the warm-up)",
            "what": "unmodified reference host (vp8enc.cpp + entropy_host.cpp) + libOpenCL.so.1 shim; Y4M file in, IVF file out; all "
                    "host<->device copies, host intra/entropy work and file I/O included; instances encode independent "
                    "keyframe-delimited segments (no collective)"}


# ------------------------------------------------------------------------------------------------
def segment_parallel_config(args, which, rank, world, local_rank, dist, tmp):
    """BASELINE configs[3] (--config 4: 1080p, 2400 frames, -g 150 -> 16 segments) and configs[4] (--config 5: 4320p,
    120 frames, -g 15 -> 8 segments) as written: the clip is cut at the serial encoder's key frames
    (segments.plan_key_frames), every rank encodes its share (segments.assign_segments) with P instances on its GPU,
    rank 0 concatenates and compares with a serial encode of the whole clip through the same shim."""
    import torch
    import gen_y4m
    from vp8oclenc_b200 import segments
    w, h, frames, gop = (1920, 1080, 2400, 150) if which == 4 else (7680, 4320, 120, 15)
    if args.frames:
        frames = args.frames
    enc = [a for a in ENC_ARGS]
    enc[enc.index("-g") + 1] = str(gop)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tmp
    base = os.path.join(shm, "vp8b200_cfg%d_%s" % (which, os.environ.get("MASTER_PORT", os.getpid())))
    y4m = os.path.join(base, "clip.y4m")
    if rank == 0:
        shutil.rmtree(base, ignore_errors=True)
        os.makedirs(base)
        gen_y4m.write_y4m(y4m, w, h, frames)
    if dist:
        dist.barrier()
    plan = segments.plan_key_frames(y4m, gop)
    nseg = len(plan[0])
    per_rank = (nseg + world - 1) // world
    P = args.e2e_procs or max(1, min(per_rank, 16, (3 * (os.cpu_count() or 2)) // (2 * world)))
    # (two contexts on a GPU time-slice well enough; bringing the MPS server up costs more than it saves then)
    daemon = segments.MpsDaemon(os.path.join(tempfile.gettempdir(), "vp8b200_mps_%s" % os.environ.get("MASTER_PORT", os.getpid())))
    if local_rank == 0 and not args.no_mps and P > 2:
        daemon.__enter__()
    if dist:
        dist.barrier()
        daemon.active = os.path.exists(os.path.join(daemon.pipe, "control"))
    try:
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n, ivfs, exact, procs = segments.encode_clip_segment_parallel(y4m, enc, gop, os.path.join(base, "rank%d" % rank), rank=rank,
                                                                      world=world, device=local_rank, per_device=P,
                                                                      mps_env=daemon.env(), plan=plan)
        t_mine = time.perf_counter() - t0
        steady = [(len(pr.stamps) - 2) / (pr.stamps[-1] - pr.stamps[1]) for pr in procs if len(pr.stamps) > 2]
        # where an instance's wall time goes: start-up until its key frame is coded, the inter frames, shutdown
        phases = [(pr.stamps[0] - pr.t_start, pr.stamps[-1] - pr.stamps[0], pr.t_end - pr.stamps[-1]) for pr in procs if pr.stamps]
        phase_avg = [sum(p[k] for p in phases) / max(1, len(phases)) for k in range(3)]
        if dist:
            tt = torch.tensor([t_mine], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_all = float(tt.item())
            gathered = [None] * world
            dist.all_gather_object(gathered, ivfs)
        else:
            t_all, gathered = t_mine, [ivfs]
        line = None
        if rank == 0:
            paths = {}
            for g in gathered:
                paths.update(g)
            joined = os.path.join(base, "joined.ivf")
            total = segments.concat_ivf([paths[i] for i in range(n)], joined)
            # the serial encode of the whole clip, one instance on this rank's GPU
            t1 = time.perf_counter()
            whole = os.path.join(base, "serial.ivf")
            segments.EncoderProcess(y4m, whole, enc, os.path.join(base, "serial_run"), device=local_rank).wait(timeout=3000)
            t_serial = time.perf_counter() - t1
            a, b = open(whole, "rb").read(), open(joined, "rb").read()
            line = {"metric": "encoded frames/s at %s, segment-parallel" % size_name(w, h), "value": frames / t_all, "unit": "frames/s",
                    "n_gpus": world, "higher_is_better": True, "scaling": "strong", "data": "synthetic", "dtype": "u8/int32",
                    "config": {"workload": "BASELINE configs[%d]: %dx%d, %d frames, -g %d: one clip cut at its %d key frames, segments "
                                           "encoded by independent instances of the unmodified host + shim (%d at a time per GPU), "
                                           "host-side IVF concatenation" % (which - 1, w, h, frames, gop, n, P),
                               "segments": n, "instances_per_gpu": P, "cuts_exact": bool(exact)},
                    "seconds_segment_parallel": t_all, "seconds_serial_one_instance": t_serial, "speedup_vs_serial_instance": t_serial / t_all,
                    "includes": "process start-up, CUDA context creation and the host-coded key frame of every segment",
                    "steady_frames_per_s_per_instance": sum(steady) / max(1, len(steady)),
                    "instance_seconds": {"start_to_key_frame_coded": phase_avg[0], "inter_frames": phase_avg[1], "shutdown": phase_avg[2],
                                         "note": "rank 0's instances, averaged"},
                    "mps": bool(daemon.active),
                    "frames_in_concatenation": total, "identical_to_serial_encode": a == b,
                    "md5": hashlib.md5(b).hexdigest(), "bytes": len(b)}
            print(json.dumps(line))
        if dist:
            dist.barrier()
    finally:
        if dist:
            dist.barrier()
        if local_rank == 0:
            daemon.__exit__(None, None, None)
        if rank == 0:
            shutil.rmtree(base, ignore_errors=True)
    return 0


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--segments", type=int, default=16, help="independent segments (engines, streams) per GPU in the `value` run")
    ap.add_argument("--frames-per-step", type=int, default=4, help="consecutive frames of every segment in one step")
    ap.add_argument("--size", default="1920x1080", help="frame size WxH (default: the 1080p configuration the metric is quoted on)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mps", action="store_true", help="do not start a CUDA MPS daemon for the multi-instance e2e run")
    ap.add_argument("--e2e-procs", type=int, default=0, help="encoder instances per GPU for the e2e run (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sizes", action="store_true", help="skip the 2160p block")
    ap.add_argument("--config", type=int, default=0, choices=[0, 4, 5], help="run BASELINE configs[3] / configs[4] as written instead")
    ap.add_argument("--frames", type=int, default=0, help="--config: clip length override")
    args = ap.parse_args()
    if args.impl == "b200":
        args.warmup = max(args.warmup, 3)
    w, h = (int(x) for x in args.size.lower().split("x"))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    tmp = tempfile.mkdtemp(prefix="vp8bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    S, F = max(1, args.segments), max(1, args.frames_per_step)
    ww, wh = padded(w, h)
    # resident input: S x (1 + (W + K) F) frames; keep it below ~24 GB
    while F > 1 and S * (1 + (args.warmup + args.steps) * F) * ww * wh * 1.5 > 24e9:
        F -= 1
    config = make_config(w, h, S, F)
    metric = "encoded frames/s at %s" % size_name(w, h)
    try:
        if args.impl == "reference":
            if rank != 0:
                return 0
            r = reference_arm(args, tmp, w, h)
            if r is None:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) is not present"}))
                return 0
            line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus,
                    "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic", "config": config,
                    "step_is": "one encoded inter frame (the CPU reference encodes one stream)",
                    "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "stages_ms_per_frame")},
                    "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
            return 0
        return b200_arm(args, w, h, S, F, rank, world, local_rank, tmp, config, metric)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def b200_arm(args, w, h, S, F, rank, world, local_rank, tmp, config, metric):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: vp8oclenc_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.config:
            return segment_parallel_config(args, args.config, rank, world, local_rank, dist, tmp)
        K, W = args.steps, args.warmup
        ww, wh = padded(w, h)
        N = ww * wh
        units = {}
        try:
            units = json.load(open(os.path.join(ROOT, "profiles", "r02_roofline_units.json")))
        except Exception:
            pass

        pipe = device_pipeline(w, h, S, F, K, W, rank, world, local_rank, dist, sample_clocks=(rank == 0))
        kernels = roofline = None
        if rank == 0:
            kernels, int_peak = kernel_lines(w, h, pipe, units)
            q = kernels.get("luma_search_2step")
            if q:
                roofline = dict(kernel="luma_search_2step", bound="int_alu", achieved=q.get("achieved"), peak=q["peak"], unit="Tiop/s",
                                frac=q.get("frac"), traffic=q.get("dram_bytes_per_ref"),
                                **{k: v for k, v in q.items() if k not in ("achieved", "peak", "unit", "frac", "bound", "dram_bytes_per_ref")})
                roofline["note"] = ("int-ALU roofline of the dominant kernel: achieved = minimal integer operations of the kernel's own "
                                    "de-duplicated formulation (DESIGN.md section 4, per luma pixel per reference) x pixels x references / "
                                    "live CUDA-event time of the launch inside a frame; peak = measured IMAD:add/logic 1:2 issue rate of "
                                    "this device (vp8b200_measure_int_ops_per_second); issue_utilisation = ALL executed thread-"
                                    "instructions (ncu, profiles/r02_roofline_units.json) against the same peak; traffic = ncu DRAM "
                                    "bytes per reference")
        for x in pipe["engines"]:
            x.close()
        value, ms_per_step, launches, clocks = pipe["value"], pipe["ms_per_step"], pipe["launches"], pipe["clocks"]
        refs_pf = pipe["refs_per_frame"]
        del pipe
        torch.cuda.empty_cache()

        # the reference encoder on the first frames of the clip: the CPU baseline (N=1) and the check of the e2e outputs
        cpu_baseline = ref_ivf = None
        if rank == 0 and not args.no_cpu_baseline:
            if world == 1:
                r = reference_arm(argparse.Namespace(steps=min(K, 12), warmup=2), tmp, w, h)
                if r:
                    cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "stages_ms_per_frame")}
                    ref_ivf = r["ivf"]
            else:
                r = reference_encode(tmp, w, h, 3, "check")
                ref_ivf = r[1] if r else None
        # every instance encodes 1 key + W warm-up + Ke timed frames; Ke >= 120 so that the timed window is long against
        # the scheduling noise of P processes on the host cores
        e2e = None if args.no_e2e else end_to_end(args, w, h, max(K, 120), W, rank, world, local_rank, dist, tmp, ref_ivf)

        sizes = None
        if not args.no_sizes and (w, h) == (1920, 1080):
            # BASELINE configs[2] (3840x2160), reduced in length so that the default run stays within minutes
            w2, h2 = 3840, 2160
            S2, F2, K2, W2 = 8, 2, max(3, min(K, 8)), 3
            p2 = device_pipeline(w2, h2, S2, F2, K2, W2, rank, world, local_rank, dist, sample_clocks=False)
            k2 = kernel_lines(w2, h2, p2, units)[0] if rank == 0 else None
            for x in p2["engines"]:
                x.close()
            v2, ms2 = p2["value"], p2["ms_per_step"]
            del p2
            torch.cuda.empty_cache()
            ref2 = None
            if rank == 0 and not args.no_cpu_baseline:
                r = reference_encode(tmp, w2, h2, 3, "check2160")
                ref2 = r[1] if r else None
            e2 = None if args.no_e2e else end_to_end(args, w2, h2, 60, 3, rank, world, local_rank, dist, tmp, ref2,
                                                     distinct=4 if world == 1 else 2)
            sizes = {"2160p": {"metric": "encoded frames/s at 2160p", "value": v2, "unit": "frames/s", "ms_per_step": ms2, "steps": K2,
                               "warmup": W2, "config": make_config(w2, h2, S2, F2), "me_mpix_per_s": v2 * 3840 * 2160 / 1e6,
                               "e2e": e2, "kernels": k2}}

        if rank == 0:
            line = {"metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
                    "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "u8/int32", "data": "synthetic", "config": config,
                    "frames_per_step": S * F * world, "refs_searched_per_frame": refs_pf, "me_mpix_per_s": value * N / 1e6,
                    "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "kernels": kernels,
                    "cpu_baseline": cpu_baseline, "sizes": sizes}
            print(json.dumps(line))
        return 0
    finally:
        if dist:
            dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
