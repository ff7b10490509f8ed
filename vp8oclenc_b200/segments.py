"""Segment-parallel encoding: the only way the path shards.

Frames depend on their references, so one stream cannot be split inside a group of pictures.
It can be cut where the serial encoder places its key frames: every piece of host state is reset
or re-sent at a key frame.  Where those are follows from the input alone as long as only the GOP
counter and the colour-difference scene detector force them: frames_until_key reloads to -g at
EVERY key frame, forced ones included (src/intra_part.h:1091), so after a scene cut at frame 70
with -g 150 the next key is frame 220, not 150.  plan_key_frames() replays that logic on the
chroma planes of the input (src/vp8enc.cpp:265-311, 364-412) and split_y4m_at() cuts there; the
concatenation then equals the serial encode byte for byte.  Key frames the serial encoder forces
AFTER an inter pass (bad SSIM / too many intra-replaced macroblocks, src/vp8enc.cpp:443-453)
depend on the encode itself and cannot be planned: with such content the concatenation is a valid
stream with the same frames but other key positions than the serial encode.  Plain cuts at
multiples of -g (split_y4m) are exact only when no key frame is forced at all.
Each segment is encoded by its own instance of
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


def _y4m_geometry(header):
    toks = header.split()
    w = int([t for t in toks if t[:1] == b"W"][0][1:])
    h = int([t for t in toks if t[:1] == b"H"][0][1:])
    return w, h


def plan_key_frames(path, gop):
    """-> (sorted key frame indices of the serial encode of `path` with -g `gop`, exact)

    Replays main()'s key-frame logic on the input: the GOP counter (src/vp8enc.cpp:364-368, reloaded by
    intra_transform(), src/intra_part.h:1091-1093) and scene_change() (src/vp8enc.cpp:265-311: mean absolute
    difference of the padded chroma planes against the previous input frame, with its hold-over for detections
    less than four frames after a key).  `exact` is False when a cut would not reproduce the detector's state
    (a key frame placed while a hold-over is pending); such a key is not reported as a cut."""
    import numpy as np
    with open(path, "rb") as f:
        w, h = _y4m_geometry(f.readline())
        cw, ch = w // 2, h // 2
        wrk_cw, wrk_ch = ((w + 15) // 16 * 16) // 2, ((h + 15) // 16 * 16) // 2
        fsz = w * h + 2 * cw * ch

        def padded(plane, right_fill):
            out = np.zeros((wrk_ch, wrk_cw), np.int32)
            out[:ch, :cw] = plane
            if wrk_cw > cw:  # copy_with_padding() (src/encIO.h:140-200) extends U to the right; V's padding is never written
                out[:ch, cw:] = plane[:, -1:] if right_fill else 0
            out[ch:, :] = out[ch - 1, :]
            return out

        keys, exact = [], True
        until_key, last_key_detect, holdover = 1, 0, 0
        last_u = last_v = None
        n = 0
        while True:
            tag = f.read(6)
            if len(tag) < 6:
                break
            blob = f.read(fsz)
            if len(blob) < fsz:
                break
            u = padded(np.frombuffer(blob, np.uint8, cw * ch, w * h).reshape(ch, cw), True)
            v = padded(np.frombuffer(blob, np.uint8, cw * ch, w * h + cw * ch).reshape(ch, cw), False)
            until_key -= 1
            key = until_key < 1
            if not key:
                size = wrk_cw * wrk_ch
                ud = int(np.abs(last_u - u).sum()) // size
                vd = int(np.abs(last_v - v).sum()) // size
                detect = ud > 7 or vd > 7 or ud + vd > 10
                since = n - last_key_detect
                if detect and since < 4:
                    last_key_detect, holdover = n, 1
                elif detect:
                    key = True
                elif holdover and since >= 4:
                    holdover, key = 0, True
            if key:
                if holdover and n > 0:
                    exact = False  # a fresh instance would start without the pending hold-over: do not cut here
                else:
                    keys.append(n)
                until_key, last_key_detect = gop, n
            last_u, last_v = u, v
            n += 1
    return keys, exact


def split_y4m_at(path, cuts, out_dir, prefix="seg"):
    """cuts a Y4M file at the given frame indices (cuts[0] must be 0); returns the paths of the pieces"""
    assert cuts and cuts[0] == 0
    with open(path, "rb") as f:
        header = f.readline()
        w, h = _y4m_geometry(header)
        fsz = 6 + w * h * 3 // 2
        paths = []
        for i, start in enumerate(cuts):
            count = (cuts[i + 1] - start) if i + 1 < len(cuts) else None
            blob = f.read(fsz * count) if count is not None else f.read()
            if not blob:
                break
            p = os.path.join(out_dir, "%s%03d.y4m" % (prefix, i))
            with open(p, "wb") as g:
                g.write(header)
                g.write(blob)
            paths.append(p)
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
        # the binary started here IS the reference host (build_host.py): the shim may use the modes that rely on its
        # habits (csrc/cl_shim.cu, VP8B200_HOST_PROFILE); env_extra can still override every single one of them
        env.setdefault("VP8B200_HOST_PROFILE", "reference")
        if device is not None:
            # one visible GPU per instance: a CUDA process that sees all eight of a node spends seconds initialising the
            # seven it will never use.  `device` counts within what this process may see.
            vis = [v for v in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if v != ""]
            env["CUDA_VISIBLE_DEVICES"] = vis[device] if device < len(vis) else str(device)
            env["VP8B200_DEVICE"] = "0"
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


def assign_segments(n_segments, world, rank):
    """which segments rank `rank` of `world` encodes: round-robin, so that every rank gets early and late
    (possibly shorter) segments alike.  The ranks share nothing but this rule and the input."""
    return list(range(rank, n_segments, world))


def encode_clip_segment_parallel(y4m, args, gop, work_dir, rank=0, world=1, device=0, per_device=1, lib_dir=SHIM_DIR,
                                 host_bin=HOST_BIN, mps_env=None, plan=None):
    """One rank's share of a segment-parallel encode of `y4m` (config 4/5 of BASELINE.json): plans the cuts
    (plan_key_frames, the same on every rank), writes and encodes the segments assign_segments() gives this rank,
    `per_device` at a time on `device`.  Returns (n_segments, {segment index: ivf path}, exact, [EncoderProcess]).
    The gather is concat_ivf() over all ranks' paths in segment order on one rank (files on a shared directory;
    the only communication the caller needs is a barrier)."""
    keys, exact = plan if plan is not None else plan_key_frames(y4m, gop)
    mine = assign_segments(len(keys), world, rank)
    os.makedirs(work_dir, exist_ok=True)
    with open(y4m, "rb") as f:
        header = f.readline()
        w, h = _y4m_geometry(header)
        fsz = 6 + w * h * 3 // 2
        base = f.tell()
    total_frames = (os.path.getsize(y4m) - base) // fsz
    seg_paths = {i: os.path.join(work_dir, "seg%04d.y4m" % i) for i in mine}

    def cut(i):  # (in-kernel copy where the file system offers it; the pieces are written concurrently)
        start = base + fsz * keys[i]
        count = fsz * ((keys[i + 1] if i + 1 < len(keys) else total_frames) - keys[i])
        with open(y4m, "rb") as src, open(seg_paths[i], "wb") as dst:
            dst.write(header)
            dst.flush()
            done = 0
            try:
                while done < count:
                    n = os.copy_file_range(src.fileno(), dst.fileno(), min(count - done, 1 << 30), start + done, len(header) + done)
                    if n <= 0:
                        break
                    done += n
            except (OSError, AttributeError):
                pass
            if done < count:
                src.seek(start + done)
                dst.seek(len(header) + done)
                while done < count:
                    blob = src.read(min(count - done, 64 << 20))
                    if not blob:
                        break
                    dst.write(blob)
                    done += len(blob)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(16, len(mine)))) as pool:
        list(pool.map(cut, mine))
    order = sorted(seg_paths)
    ivfs, procs = _encode_segments([seg_paths[i] for i in order], work_dir, args, (device,), per_device, lib_dir, host_bin,
                                   mps_env or {}, names=["seg%04d" % i for i in order])
    return len(keys), dict(zip(order, ivfs)), exact, procs


def encode_segments(seg_paths, out_dir, args, devices=(0,), per_device=1, lib_dir=SHIM_DIR, host_bin=HOST_BIN, mps=True):
    """encodes every segment, at most len(devices)*per_device at a time, round-robin over the
    devices; returns (list of ivf paths, list of EncoderProcess).  With several instances per
    device they run under a private MPS daemon when one can be started (see MpsDaemon)."""
    if mps and per_device > 1:
        with MpsDaemon(out_dir) as daemon:
            return _encode_segments(seg_paths, out_dir, args, devices, per_device, lib_dir, host_bin, daemon.env())
    return _encode_segments(seg_paths, out_dir, args, devices, per_device, lib_dir, host_bin, {})


def _encode_segments(seg_paths, out_dir, args, devices, per_device, lib_dir, host_bin, env_extra, names=None):
    slots = [d for d in devices for _ in range(per_device)]
    names = names or ["seg%03d" % i for i in range(len(seg_paths))]
    ivfs = [os.path.join(out_dir, n + ".ivf") for n in names]
    procs, running = [None] * len(seg_paths), {}
    nxt = 0
    free = list(range(len(slots)))
    while nxt < len(seg_paths) or running:
        while free and nxt < len(seg_paths):
            s = free.pop(0)
            procs[nxt] = EncoderProcess(seg_paths[nxt], ivfs[nxt], args, os.path.join(out_dir, "run_" + names[nxt]), lib_dir,
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
