"""Segment-parallel encoding: the only way the path shards.

Frames depend on their references, so one stream cannot be split inside a group of pictures.
It can be cut at multiples of the GOP size: the serial encoder places a forced key frame exactly
there (frames_until_key reloads to -g at every key, src/intra_part.h:1091) and every piece of
host state is reset or re-sent at a key frame.  Each segment is encoded by its own instance of
the UNMODIFIED reference host against our OpenCL shim, pinned to one GPU with VP8B200_DEVICE;
several instances share a GPU to hide the host's serial work.  There is no collective: the
"gather" is host-side concatenation of the IVF files --

  * keep the first 32-byte IVF header, drop the others,
  * rewrite each 12-byte frame header's timestamp to the global frame index (src/encIO.h:48),
  * set the header's frame count to total + 1, as the reference writes it (src/encIO.h:122-131).
"""
import os
import struct
import subprocess
import sys
import threading
import time

PKG = os.path.dirname(os.path.abspath(__file__))
SHIM_DIR = os.path.join(PKG, "lib")
HOST_BIN = os.environ.get("VP8B200_HOST_BIN") or os.path.join(PKG, "bin", "vp8enc")

GPU_STUB = "// vp8oclenc_b200 placeholder: kernels are built into libOpenCL.so.1; program = GPU (luma_search_1step)\n"
CPU_STUB = "// vp8oclenc_b200 placeholder: kernels are built into libOpenCL.so.1; program = CPU (encode_coefficients)\n"


def read_ivf(path):
    """-> (header bytes, [(timestamp, payload), ...])"""
    with open(path, "rb") as f:
        data = f.read()
    assert data[:4] == b"DKIF", "not an IVF file: %s" % path
    hdr_len = struct.unpack_from("<H", data, 6)[0]
    frames, pos = [], hdr_len
    while pos + 12 <= len(data):
        size, ts = struct.unpack_from("<IQ", data, pos)
        frames.append((ts, data[pos + 12:pos + 12 + size]))
        pos += 12 + size
    return data[:hdr_len], frames


def concat_ivf(paths, out_path, timestep=1):
    """host-side gather of independently encoded segments"""
    header, out = None, []
    for p in paths:
        h, frames = read_ivf(p)
        if header is None:
            header = bytearray(h)
        out.extend(payload for _, payload in frames)
    struct.pack_into("<I", header, 24, len(out) + 1)  # the reference's frame count is frames + 1 (Q14)
    with open(out_path, "wb") as f:
        f.write(bytes(header))
        for i, payload in enumerate(out):
            f.write(struct.pack("<IQ", len(payload), i * timestep))
            f.write(payload)
    return len(out)


def split_y4m(path, seg_frames, out_dir, prefix="seg"):
    """cuts a Y4M file into files of seg_frames frames; returns their paths"""
    with open(path, "rb") as f:
        header = f.readline()
        toks = header.split()
        w = int([t for t in toks if t[:1] == b"W"][0][1:])
        h = int([t for t in toks if t[:1] == b"H"][0][1:])
        fsz = 6 + w * h * 3 // 2
        paths, idx = [], 0
        while True:
            blob = f.read(fsz * seg_frames)
            if not blob:
                break
            p = os.path.join(out_dir, "%s%03d.y4m" % (prefix, idx))
            with open(p, "wb") as g:
                g.write(header)
                g.write(blob)
            paths.append(p)
            idx += 1
    return paths


class MpsDaemon:
    """CUDA Multi-Process Service for the encoder instances that share a GPU.

    Without it the contexts of the instances are time-sliced: kernels and copies of different
    instances never overlap, and on a B200 eight instances already saturate at about 320 frames/s
    (1080p).  Under MPS they share the SMs and the copy engines (690 frames/s with 16 instances on
    16 host cores).  The daemon is private to this object (its own pipe directory), so CUDA
    processes that are not handed `env()` do not see it.  If the control binary is missing or the
    daemon does not come up, `active` stays False and the instances run time-sliced as before."""

    def __init__(self, base_dir):
        self.pipe = os.path.join(base_dir, "mps_pipe")
        self.log = os.path.join(base_dir, "mps_log")
        self.active = False

    def env(self):
        return {"CUDA_MPS_PIPE_DIRECTORY": self.pipe, "CUDA_MPS_LOG_DIRECTORY": self.log} if self.active else {}

    def _control(self, text=None, args=()):
        env = dict(os.environ, CUDA_MPS_PIPE_DIRECTORY=self.pipe, CUDA_MPS_LOG_DIRECTORY=self.log)
        return subprocess.run(["nvidia-cuda-mps-control", *args], input=text, env=env, text=True, capture_output=True,
                              timeout=30)

    def __enter__(self):
        import shutil
        if os.environ.get("VP8B200_NO_MPS") or not shutil.which("nvidia-cuda-mps-control"):
            return self
        try:
            os.makedirs(self.pipe, exist_ok=True)
            os.makedirs(self.log, exist_ok=True)
            if self._control(args=("-d",)).returncode == 0:
                for _ in range(50):  # the control pipe appears once the daemon listens
                    if os.path.exists(os.path.join(self.pipe, "control")):
                        self.active = True
                        break
                    time.sleep(0.05)
        except Exception:
            self.active = False
        return self

    def __exit__(self, *exc):
        if self.active:
            try:
                self._control("quit\n")
            except Exception:
                pass
            self.active = False
        return False


class EncoderProcess:
    """one instance of the reference host program; records the completion time of every frame"""

    def __init__(self, y4m, ivf, args, workdir, lib_dir=SHIM_DIR, host_bin=HOST_BIN, device=None, env_extra=None,
                 print_info=True):
        os.makedirs(workdir, exist_ok=True)
        for name, txt in (("GPU_kernels.cl", GPU_STUB), ("CPU_kernels.cl", CPU_STUB)):
            with open(os.path.join(workdir, name), "w") as f:
                f.write(txt)
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = lib_dir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
        if device is not None:
            env["VP8B200_DEVICE"] = str(device)
        env.update(env_extra or {})
        cmd = [host_bin, "-i", y4m, "-o", ivf] + [str(a) for a in args] + (["-print-info"] if print_info else [])
        import shutil
        stdbuf = shutil.which("stdbuf")
        if stdbuf:
            cmd = [stdbuf, "-oL"] + cmd
        self.stamps, self.output = [], []
        self.t_start = time.perf_counter()
        self.proc = subprocess.Popen(cmd, cwd=workdir, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.output.append(line)
            if "br=" in line:  # printed once per finished frame (src/vp8enc.cpp:482-483)
                self.stamps.append(time.perf_counter())

    def wait(self, timeout=None):
        try:
            self.proc.wait(timeout)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            self.proc.wait()
            self.thread.join()
            raise RuntimeError("encoder did not finish within %.0f s" % timeout)
        self.thread.join()
        self.t_end = time.perf_counter()
        if self.proc.returncode != 777 % 256:  # main() returns 777
            sys.stderr.write("".join(self.output[-30:]))
            raise RuntimeError("encoder exited with %d" % self.proc.returncode)
        return self.stamps


def encode_segments(seg_paths, out_dir, args, devices=(0,), per_device=1, lib_dir=SHIM_DIR, host_bin=HOST_BIN, mps=True):
    """encodes every segment, at most len(devices)*per_device at a time, round-robin over the
    devices; returns (list of ivf paths, list of EncoderProcess).  With several instances per
    device they run under a private MPS daemon when one can be started (see MpsDaemon)."""
    if mps and per_device > 1:
        with MpsDaemon(out_dir) as daemon:
            return _encode_segments(seg_paths, out_dir, args, devices, per_device, lib_dir, host_bin, daemon.env())
    return _encode_segments(seg_paths, out_dir, args, devices, per_device, lib_dir, host_bin, {})


def _encode_segments(seg_paths, out_dir, args, devices, per_device, lib_dir, host_bin, env_extra):
    slots = [d for d in devices for _ in range(per_device)]
    ivfs = [os.path.join(out_dir, "seg%03d.ivf" % i) for i in range(len(seg_paths))]
    procs, running = [None] * len(seg_paths), {}
    nxt = 0
    free = list(range(len(slots)))
    while nxt < len(seg_paths) or running:
        while free and nxt < len(seg_paths):
            s = free.pop(0)
            procs[nxt] = EncoderProcess(seg_paths[nxt], ivfs[nxt], args, os.path.join(out_dir, "run%03d" % nxt), lib_dir,
                                        host_bin, device=slots[s], env_extra=env_extra)
            running[nxt] = s
            nxt += 1
        done = [i for i in running if procs[i].proc.poll() is not None]
        for i in done:
            procs[i].wait()
            free.append(running.pop(i))
        if not done:
            time.sleep(0.002)
    return ivfs, procs
