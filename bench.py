#!/usr/bin/env python3
"""bench.py -- encoded frames/s of vp8oclenc's inter-frame hot path on B200 (BASELINE.json).

A "step" is one encoded inter frame: pyramid + hierarchical motion search over LAST / GOLDEN /
ALTREF, reference selection, six-tap prediction, DCT/WHT/quantise, dequantise/reconstruct,
SSIM, filter mask and the normal loop filter, on a synthetic 1080p clip (tools/gen_y4m.py).

  value     frames/s of the CUDA engine with every input frame already resident in HBM
            (vp8b200_engine_inter_frame + vp8b200_engine_loop_filter), CUDA events, L2 flushed
            between timed steps
  e2e       frames/s of the UNMODIFIED reference host program running against our OpenCL shim
            (vp8oclenc_b200/lib/libOpenCL.so.1) on the same clip: Y4M in, IVF out, every
            host<->device copy, the host's intra/entropy/bitstream work and file I/O included
  roofline  the dominant kernel (quarter-pel motion search) against the measured integer
            issue rate of the device, plus an HBM line for the loop filter
  cpu_baseline / --impl reference
            the reference itself (its own host + its own .cl kernels compiled for the CPU,
            oracle/_ref) on a bounded sample of the same clip on all host cores

Multi-GPU (torchrun, one rank per GPU): independent keyframe-delimited segments per GPU, no
collective on the data path ("scaling": "weak").
"""
import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

WIDTH, HEIGHT = 1920, 1080          # configs[1] of BASELINE.json
WRK_W, WRK_H = 1920, 1088           # padded to macroblocks as the host does (src/init.h:381-386)
ENC_ARGS = ["-qmin", "24", "-qmax", "24", "-g", "150", "-altref-range", "5", "-partitions", "8", "-threads", "12"]
ALTREF_RANGE = 5
QI = (24, 24, 24, 24)

GPU_STUB = "// vp8oclenc_b200 placeholder: kernels are built into libOpenCL.so.1; program = GPU (luma_search_1step)\n"
CPU_STUB = "// vp8oclenc_b200 placeholder: kernels are built into libOpenCL.so.1; program = CPU (encode_coefficients)\n"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), polled through
    NVML every few milliseconds (the timed region of the default run is shorter than one
    nvidia-smi sampling period)"""
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


def run_encoder_timed(host_bin, lib_dir, workdir, y4m, ivf, frames_total, env_extra=None):
    """runs the reference host program and timestamps its per-frame "-print-info" lines.
    returns (list of completion times per frame, stdout)"""
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "GPU_kernels.cl"), "w") as f:
        f.write(GPU_STUB)
    with open(os.path.join(workdir, "CPU_kernels.cl"), "w") as f:
        f.write(CPU_STUB)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = lib_dir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    env.update(env_extra or {})
    cmd = [host_bin, "-i", y4m, "-o", ivf] + ENC_ARGS + ["-print-info"]
    stdbuf = shutil.which("stdbuf")
    if stdbuf:
        cmd = [stdbuf, "-oL"] + cmd
    t0 = time.perf_counter()
    p = subprocess.Popen(cmd, cwd=workdir, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    stamps, out = [], []
    for line in p.stdout:
        out.append(line)
        if "br=" in line:  # printed once per finished frame (src/vp8enc.cpp:482-483)
            stamps.append(time.perf_counter() - t0)
    p.wait()
    if p.returncode != 777 % 256:
        sys.stderr.write("".join(out[-30:]))
        raise RuntimeError("encoder exited with %d" % p.returncode)
    if len(stamps) != frames_total:
        raise RuntimeError("expected %d frame lines, saw %d" % (frames_total, len(stamps)))
    return stamps, "".join(out)


def reference_arm(args, tmp):
    """the reference's own CPU implementation: its host + its .cl kernels compiled for the CPU
    (oracle/_ref), all host cores (OpenMP), on a bounded sample of the same workload"""
    import gen_y4m
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    host_bin = os.path.join(ref_dir, "vp8enc")
    if not os.path.exists(host_bin) or not os.path.exists(os.path.join(ref_dir, "libOpenCL.so.1")):
        return None
    warm = max(1, min(args.warmup, 2))
    steps = max(1, min(args.steps, args.ref_frames))
    n = 1 + warm + steps
    y4m = os.path.join(tmp, "ref_clip.y4m")
    gen_y4m.write_y4m(y4m, WIDTH, HEIGHT, n)
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
    stamps, _ = run_encoder_timed(host_bin, ref_dir, os.path.join(tmp, "ref_run"), y4m, os.path.join(tmp, "ref.ivf"), n,
                                  env_extra={"OMP_NUM_THREADS": str(os.cpu_count() or 1)})
    dt = stamps[-1] - stamps[warm]  # frame 0 is the key frame, then `warm` untimed inter frames
    fps = steps / dt
    return {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference",
            "sample": "%d inter frames of the 1080p clip after 1 key + %d warm-up frames; unmodified reference host with "
                      "its own .cl kernels compiled for the CPU (oracle/_ref), OpenMP over work-items" % (steps, warm),
            "ms_per_step": 1000.0 * dt / steps, "steps": steps, "warmup": warm}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-frames", type=int, default=6, help="timed frames of the CPU reference sample")
    ap.add_argument("--segments", type=int, default=16, help="independent segments (engines, streams) per GPU in the `value` run")
    ap.add_argument("--size", default="1920x1080", help="frame size WxH (default: the 1080p configuration the metric is quoted on; "
                                                         "other sizes are informational)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mps", action="store_true", help="do not start a CUDA MPS daemon for the multi-instance e2e run")
    ap.add_argument("--e2e-procs", type=int, default=0, help="encoder instances per GPU for the e2e run (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    global WIDTH, HEIGHT, WRK_W, WRK_H
    WIDTH, HEIGHT = (int(x) for x in args.size.lower().split("x"))
    WRK_W, WRK_H = (WIDTH + 15) // 16 * 16, (HEIGHT + 15) // 16 * 16
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    tmp = tempfile.mkdtemp(prefix="vp8bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    config = {"workload": "%dx%d synthetic YUV420 (tools/gen_y4m.py), padded to %dx%d, LAST+GOLDEN+ALTREF, "
                          "q=24, altref-range 5, 8 partitions, loop filter on the GPU" % (WIDTH, HEIGHT, WRK_W, WRK_H),
              "frame_size": [WIDTH, HEIGHT], "segments_per_gpu": max(1, args.segments),
              "step": "one frame of each of the segments_per_gpu independent segments (one engine + stream per segment)"}
    try:
        if args.impl == "reference":
            if rank != 0:
                return 0
            r = reference_arm(args, tmp)
            if r is None:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) is not present"}))
                return 0
            line = {"impl": "reference", "metric": "encoded frames/s at %s" % ("1080p" if HEIGHT == 1080 else "%dx%d" % (WIDTH, HEIGHT)), "value": r["value"], "unit": "frames/s",
                    "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32",
                    "data": "synthetic", "config": config,
                    "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
            return 0
        return b200_arm(args, rank, world, local_rank, tmp, config)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def b200_arm(args, rank, world, local_rank, tmp, config):
    import numpy as np
    import torch
    import gen_y4m
    from vp8oclenc_b200 import host as eng
    from vp8oclenc_b200.hostlogic import HostState, make_segment_data

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: vp8oclenc_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, args.warmup
    M = (WRK_W // 16) * (WRK_H // 16)
    N = WRK_W * WRK_H

    # ---- synthetic frames: every rank encodes S keyframe-delimited segments of its own ----------
    # The path shards by independent segments.  A step is one frame of EACH of the S
    # segments resident on this GPU, each segment on its own engine and stream: the dependency-bound
    # loop filter of one segment overlaps the motion search of the others.
    S = max(1, args.segments)
    nframes = 1 + W + K
    clip = gen_y4m.Clip(WRK_W, WRK_H)
    dev_frames = []
    for sgm in range(S):
        first = (rank * S + sgm) * nframes
        host_frames = [clip.frame(first + i) for i in range(nframes)]
        dev_frames.append([[torch.from_numpy(np.ascontiguousarray(p)).cuda() for p in f] for f in host_frames])
    sd = make_segment_data(QI)

    engines = [eng.Engine(WRK_W, WRK_H) for _ in range(S)]
    e = engines[0]
    stream_ptr = e.stream
    ext_streams = [torch.cuda.ExternalStream(x.stream) for x in engines]
    ext_stream = ext_streams[0]
    master = torch.cuda.Stream()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    torch.cuda.synchronize()
    states = [HostState(10 ** 6, ALTREF_RANGE) for _ in range(S)]
    for sgm in range(S):
        states[sgm].next_frame()
        engines[sgm].set_reconstruction(*dev_frames[sgm][0])  # the key frame's reconstruction seeds LAST/GOLDEN/ALTREF

    def step(sgm, i):
        st = states[sgm].next_frame()
        y, u, v = dev_frames[sgm][i]
        x = engines[sgm]
        x.inter_frame(y, u, v, sd, -1.0, st["prev_golden"], st["prev_altref"], st["altref_differs"])
        n1 = x.last_launch_count
        x.loop_filter(None)
        return n1 + x.last_launch_count, 1 + (not st["prev_golden"]) + (not st["prev_altref"] and st["altref_differs"])

    for i in range(1, 1 + W):
        for sgm in range(S):
            step(sgm, i)
    for x in engines:
        x.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    done = [[torch.cuda.Event() for _ in range(S)] for _ in range(K)]
    launches, refs_searched = 0, 0
    for k in range(K):
        with torch.cuda.stream(master):
            flush_buf.fill_(k & 255)           # L2 flush between timed steps, outside the timed span
            ev[k][0].record(master)
        for sgm in range(S):
            ext_streams[sgm].wait_event(ev[k][0])
            n, r = step(sgm, 1 + W + k)
            done[k][sgm].record(ext_streams[sgm])
            launches += n
            refs_searched += r
        for sgm in range(S):
            master.wait_event(done[k][sgm])
        ev[k][1].record(master)
    for x in engines:
        x.synchronize()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * K * S / (total_ms * 1e-3)
    refs_searched /= S
    dev_frames = dev_frames[0]

    # ---- roofline of the dominant kernel, timed live on the stream it is launched on -------------
    roofline = roofline_hbm = None
    if rank == 0:
        L = eng.lib()
        L.vp8b200_measure_int_ops_per_second.restype = ctypes.c_double
        int_peak = L.vp8b200_measure_int_ops_per_second(ctypes.c_void_p(stream_ptr), 5)
        cur_y = dev_frames[-1][0]
        ref_y = dev_frames[-2][0]
        nb = N // 64
        net = torch.zeros((nb, 2), dtype=torch.int16, device="cuda")
        outn = torch.zeros((nb, 2), dtype=torch.int16, device="cuda")
        met = torch.zeros(nb, dtype=torch.int32, device="cuda")
        reps = 10
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        with torch.cuda.stream(ext_stream):
            for _ in range(3):
                eng.luma_search_2step(cur_y, ref_y, net, outn, met, WRK_W, WRK_H)
            for a, b in evs:
                flush_buf.fill_(1)
                a.record(ext_stream)
                eng.luma_search_2step(cur_y, ref_y, net, outn, met, WRK_W, WRK_H)
                b.record(ext_stream)
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / reps
        # Per-unit figure: thread-instructions this kernel executes per luma pixel per reference, from the ncu
        # capture committed under profiles/ (tools/ncu_summary.py units).  The reference's scalar formulation is
        # 1381 int-ops per pixel; the kernel needs fewer (dp4a six-tap, packed lanes, shared loads),
        # so counting 1381 against the issue peak would overstate the fraction.
        units = {}
        try:
            units = json.load(open(os.path.join(ROOT, "profiles", "r01_roofline_units.json")))
        except Exception:
            pass
        per_px = float(units.get("thread_instr_per_pixel_per_ref", 1381.0))
        ops = per_px * N
        roofline = {"kernel": "luma_search_2step", "bound": "int_alu", "achieved": ops / (ms * 1e-3) / 1e12,
                    "peak": int_peak / 1e12, "unit": "Tiop/s", "frac": (ops / (ms * 1e-3)) / int_peak,
                    "traffic": units.get("dram_bytes_per_ref"), "ms_per_launch": ms,
                    "ops_per_pixel": per_px, "reference_ops_per_pixel": 1381.0,
                    "reference_formulation_tiops": 1381.0 * N / (ms * 1e-3) / 1e12,
                    "note": "int-ALU issue roofline: achieved = executed thread-instructions per pixel (ncu, "
                            "profiles/r01_roofline_units.json) x pixels / live CUDA-event time of one launch for one "
                            "reference; peak = measured IMAD:add/logic 1:2 issue rate of this device "
                            "(vp8b200_measure_int_ops_per_second); traffic = ncu DRAM bytes per reference-launch"}
        # loop filter: HBM line (3N bytes read + written once each -> 3N algorithmic bytes per SURVEY 8d)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        seg = torch.zeros(M, dtype=torch.int32, device="cuda")
        mask = torch.full((M,), -1, dtype=torch.int32, device="cuda")
        sd_dev = torch.from_numpy(sd).cuda()
        planes = [t.clone() for t in dev_frames[-1]]
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        with torch.cuda.stream(ext_stream):
            for _ in range(2):
                eng.loop_filter_planes(planes[0], planes[1], planes[2], seg, mask, sd_dev, WRK_W, WRK_H)
            for a, b in evs:
                flush_buf.fill_(2)
                a.record(ext_stream)
                eng.loop_filter_planes(planes[0], planes[1], planes[2], seg, mask, sd_dev, WRK_W, WRK_H)
                b.record(ext_stream)
        torch.cuda.synchronize()
        ms_lf = sum(a.elapsed_time(b) for a, b in evs) / reps
        gbs = 3.0 * N / (ms_lf * 1e-3) / 1e9
        roofline_hbm = {"kernel": "loop_filter_planes", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": gbs / hbm_peak, "traffic": None, "ms_per_launch": ms_lf,
                        "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                        "note": "dependency-bound wavefront of mb_w+2(mb_h-1)=%d stages, not bandwidth-bound" % (WRK_W // 16 + 2 * (WRK_H // 16 - 1))}
    for x in engines:
        x.close()

    # ---- end to end: the unmodified reference host against our OpenCL shim -------------------------
    # Every rank runs `procs` encoder instances on its GPU, each on its own keyframe-delimited
    # segment of 1 + W + K frames (vp8oclenc_b200/segments.py); frames/s counts the frames all
    # instances finish between "every instance is past its warm-up" and "the last instance is done".
    e2e = None
    if not args.no_e2e:
        from vp8oclenc_b200 import segments
        if os.path.exists(segments.HOST_BIN):
            # every instance encodes 1 key + W warm-up + Ke timed frames; Ke >= 120 so that the timed window is long
            # against the scheduling noise of P processes on the host cores
            Ke = max(K, 120)
            n = 1 + W + Ke
            distinct = 8 if world == 1 else (4 if world == 2 else 2)  # clips per rank (they live in /dev/shm)
            clips = []
            gate_root = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()

            def run(P, tag, env_more=None):
                # a few distinct segments per rank, written once; further instances re-encode one of them into their
                # own output
                paths, outs = [], []
                for p in range(P):
                    if p % distinct >= len(clips):
                        y4m = os.path.join(tmp, "e2e_%d_%d.y4m" % (rank, p % distinct))
                        gen_y4m.write_y4m(y4m, WIDTH, HEIGHT, n, start=(rank * distinct + p) * n)
                        clips.append(y4m)
                    paths.append(clips[p % distinct])
                    outs.append(os.path.join(tmp, "e2e_%s_%d_%d" % (tag, rank, p)))
                # start gate (cl_shim.cu start_gate): all P x world instances come up (15-30 s for 32 of them: context
                # creation, module loading and page pinning are serialised by the driver), encode their key frame
                # and two inter frames, then go on together; frames 3..W are the common warm-up
                gate = os.path.join(gate_root, "vp8b200_gate_%s_%s" % (os.environ.get("MASTER_PORT", os.getpid()), tag))
                if rank == 0:
                    shutil.rmtree(gate, ignore_errors=True)
                    os.makedirs(gate)
                if dist:
                    dist.barrier()
                procs = [segments.EncoderProcess(paths[p], outs[p] + ".ivf", ENC_ARGS,
                                                 os.path.join(tmp, "run_%s_%d_%d" % (tag, rank, p)), device=local_rank,
                                                 env_extra=dict(env_more or {}, VP8B200_STATS=outs[p] + ".stats",
                                                                VP8B200_START_GATE="%s:%d:3" % (gate, P * world)))
                         for p in range(P)]
                # a failure on one rank must not leave the others waiting in a collective: agree on it first
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
                if dist:
                    cc = torch.tensor([float(count)], device="cuda", dtype=torch.float64)
                    dist.all_reduce(cc, op=dist.ReduceOp.SUM)
                    count = int(cc.item())
                h2d = d2h = launches_ps = None
                try:
                    s0 = json.load(open(outs[0] + ".stats"))
                    h2d, d2h, launches_ps = int(s0["h2d_bytes"] / n), int(s0["d2h_bytes"] / n), s0["kernel_launches"] / n
                except Exception:
                    pass
                if dist:
                    dist.barrier()
                if rank == 0:
                    shutil.rmtree(gate, ignore_errors=True)
                for pth in set(o + ".ivf" for o in outs):
                    try:
                        os.remove(pth)
                    except OSError:
                        pass
                return count / (t1 - t0), count, h2d, d2h, launches_ps

            # several instances per GPU share it through a private CUDA MPS daemon (one per node, started by
            # local rank 0); without MPS the contexts are time-sliced and more than ~8 instances do not pay
            import tempfile as _tf
            daemon = segments.MpsDaemon(os.path.join(_tf.gettempdir(), "vp8b200_mps_%s" % os.environ.get("MASTER_PORT", os.getpid())))
            if local_rank == 0 and args.e2e_procs != 1 and not args.no_mps:
                daemon.__enter__()
            if dist:
                dist.barrier()
                daemon.active = os.path.exists(os.path.join(daemon.pipe, "control"))
            cores = os.cpu_count() or 2
            if daemon.active:
                # two instances per host core, at most 32 per GPU and 128 on the node (the driver brings contexts up
                # one after the other, about half a second each)
                P = args.e2e_procs or max(1, min(32, (2 * cores) // world, max(4, 128 // world)))
            else:
                P = args.e2e_procs or max(1, min(8, cores // max(2, 2 * world)))
            e2e_error = None
            fps1 = fpsP = 0.0
            h2d = d2h = lps = None
            try:
                fps1, _, h2d, d2h, lps = run(1, "single")
                try:
                    fpsP, cnt, _, _, _ = (fps1, Ke, 0, 0, 0) if P == 1 else run(P, "multi", dict(daemon.env(), VP8B200_SYNC="yield"))
                except RuntimeError as err:
                    if dist or not daemon.active:
                        raise
                    # the instances could not run under the MPS daemon on this box: time-sliced contexts instead
                    sys.stderr.write("bench: multi-instance run under MPS failed (%s); retrying without MPS\n" % err)
                    daemon.__exit__(None, None, None)
                    P = max(1, min(8, cores // 2))
                    fpsP, cnt, _, _, _ = run(P, "multi_nomps", {})
            except RuntimeError as err:
                # (run() has made sure every rank raises together)  The line is still printed, without an e2e value.
                e2e_error = str(err)
                sys.stderr.write("bench: end-to-end run failed: %s\n" % err)
            finally:
                mps_used = daemon.active
                if dist:
                    dist.barrier()
                if local_rank == 0:
                    daemon.__exit__(None, None, None)
            best = max(fps1, fpsP)
            if e2e_error or best <= 0:
                e2e = {"value": None, "unit": "frames/s", "error": e2e_error or "no frames", "h2d_bytes_per_step": None,
                       "d2h_bytes_per_step": None}
            else:
                e2e = {"value": best, "unit": "frames/s", "ms_per_step": 1000.0 / best,
                       "processes_per_gpu": P if fpsP >= fps1 else 1, "single_process_fps": fps1,
                       "multi_process_fps": fpsP, "mps": bool(mps_used and P > 1), "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "shim_kernel_launches_per_step": lps, "timed_frames_per_instance": Ke,
                       "window": "from the moment the last instance has finished its key + warm-up frames to the moment the "
                                 "last instance is done; instances wait for each other at the start of their third inter "
                                 "frame (start gate, inside the warm-up)",
                       "what": "unmodified reference host (vp8enc.cpp + entropy_host.cpp) + libOpenCL.so.1 shim; Y4M file in, "
                               "IVF file out; all host<->device copies, host intra/entropy work and file I/O included; "
                               "instances encode independent keyframe-delimited segments (no collective)"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = reference_arm(args, tmp)
        if r:
            cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "encoded frames/s at %s" % ("1080p" if HEIGHT == 1080 else "%dx%d" % (WIDTH, HEIGHT)), "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
                "config": dict(config, l2="flushed between timed steps (256 MiB fill outside the timed span)",
                               refs_searched_per_frame=refs_searched / K,
                               me_mpix_per_s=value * N / 1e6),
                "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline,
                "roofline_hbm": roofline_hbm, "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
