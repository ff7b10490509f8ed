#!/usr/bin/env python3
"""Where do the HOST CORES go in the end-to-end path?  Runs P instances of the unmodified reference host + shim
per experiment (private MPS daemon for P > 1, start gate as in bench.py) and prints, per encoded frame:
frames/s, core-ms of the instances (rusage user / sys), the part of it inside the shim's entry points by kind, the
part the waiting thread burned in waits, faults / mprotects / waits per frame, busy cores of the whole machine
(/proc/stat) and the CPU time of the MPS server.  GPU only.

    python tools/e2e_cpu.py [--size WxH] [--frames N] [--devices 0,1,..] [--json out.json] SPEC [SPEC ...]
    SPEC = P[:KEY=VALUE[,KEY=VALUE...]]      e.g.  16:VP8B200_SYNC=block   32:VP8B200_SYNC=yield,VP8B200_ELIDE=track
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_y4m  # noqa: E402
from vp8oclenc_b200 import segments  # noqa: E402

ENC_ARGS = ["-qmin", "24", "-qmax", "24", "-g", "150", "-altref-range", "5", "-partitions", "8", "-threads", "12"]
WARM = 5


def proc_stat():
    with open("/proc/stat") as f:
        v = [int(x) for x in f.readline().split()[1:]]
    idle = v[3] + v[4]
    return sum(v), idle


def mps_server_cpu():
    """user+sys seconds of every nvidia-cuda-mps-server process (read-only look at /proc)"""
    tot, hz = 0.0, os.sysconf("SC_CLK_TCK")
    for pid in os.listdir("/proc"):
        if not pid.isdigit():
            continue
        try:
            with open("/proc/%s/comm" % pid) as f:
                if not f.read().startswith("nvidia-cuda-mps"):
                    continue
            with open("/proc/%s/stat" % pid) as f:
                parts = f.read().rsplit(")", 1)[1].split()
            tot += (int(parts[11]) + int(parts[12])) / hz
        except OSError:
            pass
    return tot


def run(spec, clips, tmp, n, devices, tag):
    P = int(spec.split(":")[0])
    env = dict(kv.split("=", 1) for kv in spec.split(":", 1)[1].split(",")) if ":" in spec else {}
    total = P * len(devices)
    gate = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tmp, "gate_%d_%s" % (os.getpid(), tag))
    shutil.rmtree(gate, ignore_errors=True)
    os.makedirs(gate)
    daemon = segments.MpsDaemon(os.path.join(tmp, "mps_%s" % tag))
    if total > 1 and not env.pop("NOMPS", None):
        daemon.__enter__()
    try:
        t_start = time.perf_counter()
        procs = []
        for d in devices:
            for p in range(P):
                i = len(procs)
                procs.append(segments.EncoderProcess(
                    clips[i % len(clips)], os.path.join(tmp, "o_%s_%d.ivf" % (tag, i)), ENC_ARGS, os.path.join(tmp, "run_%s_%d" % (tag, i)),
                    device=d, env_extra=dict(daemon.env(), **env, VP8B200_STATS=os.path.join(tmp, "st_%s_%d.json" % (tag, i)),
                                             VP8B200_STATS_FROM=str(WARM),
                                             VP8B200_START_GATE="%s:%d:3" % (gate, total))))
        # CPU accounting window: from "everybody is through the gate" (approximately: first stamp index WARM) to the end
        while min(len(pr.stamps) for pr in procs) <= WARM and all(pr.proc.poll() is None for pr in procs):
            time.sleep(0.01)
        s0, m0, w0 = proc_stat(), mps_server_cpu(), time.perf_counter()
        while any(pr.proc.poll() is None for pr in procs):
            time.sleep(0.01)
            if min(len(pr.stamps) for pr in procs) >= n - 2:
                break
        s1, m1, w1 = proc_stat(), mps_server_cpu(), time.perf_counter()
        stamps = [pr.wait(timeout=600) for pr in procs]
    finally:
        daemon.__exit__(None, None, None)
        shutil.rmtree(gate, ignore_errors=True)
    t0 = max(st[WARM] for st in stamps)
    t1 = max(st[-1] for st in stamps)
    count = sum(1 for st in stamps for x in st if x > t0)
    fps = count / (t1 - t0)
    st = [json.load(open(os.path.join(tmp, "st_%s_%d.json" % (tag, i)))) for i in range(total)]
    frames = sum(x["window_frames"] for x in st)  # steady state: from inter frame WARM on (no start-up, no key frame)
    per = lambda k: sum(x.get(k, 0.0) for x in st) / frames  # noqa: E731
    hz = os.sysconf("SC_CLK_TCK")
    busy = ((s1[0] - s0[0]) - (s1[1] - s0[1])) / hz / max(1e-9, w1 - w0)
    res = {
        "spec": spec, "instances": total, "fps": fps, "startup_s": t0 - t_start,
        "core_ms_per_frame_machine": 1000.0 * busy / fps if fps else None, "busy_cores": busy, "cores": os.cpu_count(),
        "cpu_user_ms": per("cpu_user_ms"), "cpu_sys_ms": per("cpu_sys_ms"),
        "calling_thread_cpu_ms": per("cpu_ms_calling_thread"),
        "shim_cpu_ms": {k[7:]: per(k) for k in st[0] if k.startswith("cpu_ms_") and k not in ("cpu_ms_calling_thread", "cpu_ms_wait")},
        "shim_wall_ms": {k[3:]: per(k) for k in st[0] if k.startswith("ms_") and k not in ("ms_total", "ms_wait")},
        "wait_wall_ms": per("ms_wait"), "wait_cpu_ms": per("cpu_ms_wait"), "waits": per("waits"), "faults": per("faults"),
        "waits_skipped": per("waits_skipped"), "vol_ctx_switches": per("vol_ctx_switches"), "invol_ctx_switches": per("invol_ctx_switches"),
        "mprotects": per("mprotects"), "launches": per("kernel_launches"), "host_kernels": per("host_kernels"),
        "h2d_mb": per("h2d_bytes") / 1e6, "d2h_mb": per("d2h_bytes") / 1e6, "elided_mb": per("elided_bytes") / 1e6,
        "wall_ms_per_frame_instance": per("ms_total"), "mps_server_cpu_cores": (m1 - m0) / max(1e-9, w1 - w0),
    }
    shim_cpu = sum(res["shim_cpu_ms"].values())
    res["host_program_cpu_ms"] = res["calling_thread_cpu_ms"] - shim_cpu
    res["other_threads_cpu_ms"] = res["cpu_user_ms"] + res["cpu_sys_ms"] - res["calling_thread_cpu_ms"]
    return res


def show(r):
    print("== %s  (%d instances)  %.1f frames/s   busy cores %.1f of %d -> %.2f core-ms/frame (machine)   start-up %.1f s" %
          (r["spec"], r["instances"], r["fps"], r["busy_cores"], r["cores"], r["core_ms_per_frame_machine"], r["startup_s"]))
    print("   per frame: instance CPU user %.2f + sys %.2f ms | calling thread %.2f ms = host program %.2f + shim %.2f "
          "(of it waits %.2f) | other threads %.2f" %
          (r["cpu_user_ms"], r["cpu_sys_ms"], r["calling_thread_cpu_ms"], r["host_program_cpu_ms"],
           sum(r["shim_cpu_ms"].values()), r["wait_cpu_ms"], r["other_threads_cpu_ms"]))
    print("   shim CPU ms by kind: " + "  ".join("%s %.3f" % kv for kv in r["shim_cpu_ms"].items()))
    print("   shim wall ms by kind: " + "  ".join("%s %.3f" % kv for kv in r["shim_wall_ms"].items()) +
          "   wall per frame per instance %.2f" % r["wall_ms_per_frame_instance"])
    print("   waits skipped %.1f  context switches voluntary %.1f involuntary %.1f" % (r["waits_skipped"], r["vol_ctx_switches"], r["invol_ctx_switches"]))
    print("   waits %.1f (wall %.2f ms, cpu %.2f ms)  faults %.1f  mprotects %.1f  launches %.1f  h2d %.2f MB  d2h %.2f MB  "
          "d2d-elided %.1f MB  mps server %.2f cores" %
          (r["waits"], r["wait_wall_ms"], r["wait_cpu_ms"], r["faults"], r["mprotects"], r["launches"], r["h2d_mb"], r["d2h_mb"],
           r["elided_mb"], r["mps_server_cpu_cores"]), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--frames", type=int, default=66)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--clips", type=int, default=4)
    ap.add_argument("--json", default=None)
    ap.add_argument("specs", nargs="+")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    devices = [int(x) for x in a.devices.split(",")]
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    results = []
    with tempfile.TemporaryDirectory(dir=base) as tmp:
        clips = []
        for c in range(a.clips):
            y4m = os.path.join(tmp, "clip%d.y4m" % c)
            gen_y4m.write_y4m(y4m, w, h, a.frames, start=c * a.frames)
            clips.append(y4m)
        for i, spec in enumerate(a.specs):
            try:
                r = run(spec, clips, tmp, a.frames, devices, "x%d" % i)
                results.append(r)
                show(r)
            except Exception as e:  # noqa: BLE001
                print("== %s FAILED: %s" % (spec, e), flush=True)
            for f in os.listdir(tmp):
                if f.endswith(".ivf"):
                    os.remove(os.path.join(tmp, f))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(results, f, indent=1)
