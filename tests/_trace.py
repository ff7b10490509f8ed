"""Running the reference host binary and reading the OpenCL call traces it leaves.

oracle/_ref/vp8enc is the UNMODIFIED reference host (src/vp8enc.cpp + src/entropy_host.cpp);
which libOpenCL.so.1 it loads decides where the kernels run:
  oracle/_ref            -> the reference's own kernels on the CPU (the pin)
  vp8oclenc_b200/lib     -> the CUDA engine (the product)
Both runtimes write the same trace format when VP8CL_TRACE=<file> is set:
  {u32 kind, u32 mem index, u64 offset, u64 size, payload}
  kind 0 create buffer, 1 create image (offset=w, size=h), 2 WriteBuffer, 3 ReadBuffer,
  4 WriteImage, 5 UnmapMemObject (whole buffer as handed back by the host)
Memory objects are numbered in creation order, which init_all() fixes (src/init.h:442-593).
"""
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
HOST_BIN = os.path.join(REF_DIR, "vp8enc")

# creation order of the cl_mem objects in init_all() for GOP>1 without -loop-filter-on-gpu
MEM_NAMES = [
    "predictors_Y", "predictors_U", "predictors_V", "residual_Y", "residual_U", "residual_V",
    "current_frame_Y", "current_Y_by2", "current_Y_by4", "current_Y_by8", "current_Y_by16",
    "current_frame_U", "current_frame_V",
    "last_Y_by2", "last_Y_by4", "last_Y_by8", "last_Y_by16",
    "golden_Y_by2", "golden_Y_by4", "golden_Y_by8", "golden_Y_by16",
    "altref_Y_by2", "altref_Y_by4", "altref_Y_by8", "altref_Y_by16",
    "reconstructed_frame_Y", "reconstructed_frame_U", "reconstructed_frame_V",
    "golden_frame_Y", "altref_frame_Y",
    "last_vnet1", "golden_vnet1", "altref_vnet1", "last_vnet2", "golden_vnet2", "altref_vnet2",
    "metrics1", "metrics2", "metrics3",
    "macroblock_coeffs_gpu", "macroblock_non_zero_coeffs_gpu", "macroblock_parts_gpu",
    "macroblock_reference_frame_gpu", "macroblock_segment_id_gpu", "macroblock_SSIM_gpu",
    "macroblock_vectors_gpu",
    "mb_mask", "cpu_frame_Y", "cpu_frame_U", "cpu_frame_V", "segments_data_gpu", "segments_data_cpu",
    "last_frame_Y_image", "last_frame_U_image", "last_frame_V_image",
    "golden_frame_Y_image", "golden_frame_U_image", "golden_frame_V_image",
    "altref_frame_Y_image", "altref_frame_U_image", "altref_frame_V_image",
    "partitions", "partitions_sizes", "third_context", "coeff_probs", "coeff_probs_denom",
    "macroblock_coeffs_cpu", "macroblock_non_zero_coeffs_cpu", "macroblock_parts_cpu",
    "macroblock_segment_id_cpu",
]
IDX = {n: i for i, n in enumerate(MEM_NAMES)}

GPU_STUB = "// vp8oclenc_b200 placeholder: kernels are built into libOpenCL.so.1; program = GPU (luma_search_1step)\n"
CPU_STUB = "// vp8oclenc_b200 placeholder: kernels are built into libOpenCL.so.1; program = CPU (encode_coefficients)\n"


def have_host():
    return os.path.exists(HOST_BIN)


def run_host(lib_dir, workdir, y4m, ivf, args, trace=None, env_extra=None, timeout=3600):
    """runs the reference host in workdir against the libOpenCL.so.1 found in lib_dir"""
    os.makedirs(workdir, exist_ok=True)
    # the host insists on two program files of >= 32 bytes in the CWD (src/init.h:159-171)
    with open(os.path.join(workdir, "GPU_kernels.cl"), "w") as f:
        f.write(GPU_STUB)
    with open(os.path.join(workdir, "CPU_kernels.cl"), "w") as f:
        f.write(CPU_STUB)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = lib_dir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    if trace:
        env["VP8CL_TRACE"] = trace
    env.update(env_extra or {})
    cmd = [HOST_BIN, "-i", y4m, "-o", ivf] + [str(a) for a in args]
    p = subprocess.run(cmd, cwd=workdir, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    # main() returns 777 (src/vp8enc.cpp:498) -> exit status 9
    if p.returncode != 777 % 256:
        sys.stderr.write(p.stdout.decode(errors="replace")[-4000:])
        raise RuntimeError("reference host exited with %d" % p.returncode)
    return p.stdout.decode(errors="replace")


def read_trace(path):
    """-> list of (kind, name, offset, size, payload-bytes-or-None)"""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        kind, idx, off, size = struct.unpack_from("<IIQQ", data, pos)
        pos += 24
        payload = None
        if kind >= 2:
            payload = data[pos:pos + size]
            pos += size
        out.append((kind, MEM_NAMES[idx] if idx < len(MEM_NAMES) else "mem%d" % idx, off, size, payload))
    return out


def split_frames(events):
    """groups the events of one encode into frames.  A frame's device work ends with the host's
    upload of MB_segment_id to the CPU device (src/vp8enc.cpp:469); everything up to the next
    such upload belongs to the next frame."""
    frames = []
    cur = {"w": {}, "r": {}, "img": {}, "unmap": {}, "order": []}
    for kind, name, off, size, payload in events:
        if kind < 2:
            continue
        key = {2: "w", 3: "r", 4: "img", 5: "unmap"}[kind]
        if kind == 5 and name == "macroblock_coeffs_cpu" and cur["order"]:
            pass
        cur[key].setdefault(name, []).append(payload)
        cur["order"].append((key, name))
        if kind == 2 and name == "macroblock_segment_id_cpu":
            frames.append(cur)
            cur = {"w": {}, "r": {}, "img": {}, "unmap": {}, "order": []}
    frames_tail = cur
    return frames, frames_tail


def arr(payload, dtype, shape=None):
    a = np.frombuffer(payload, dtype=dtype).copy()
    return a.reshape(shape) if shape is not None else a


class HostState:
    """the reference-frame bookkeeping of main() (src/vp8enc.cpp:340-374)"""

    def __init__(self, gop, altref_range):
        self.gop, self.altref_range = gop, altref_range
        self.until_key, self.until_altref = 1, 2
        self.n = 0
        self.golden_no = self.altref_no = -1
        self.cur_key = self.cur_golden = self.cur_altref = 0

    def next_frame(self, forced_key=False):
        self.prev_key, self.prev_golden, self.prev_altref = self.cur_key, self.cur_golden, self.cur_altref
        self.until_key -= 1
        self.until_altref -= 1
        self.cur_key = int(self.until_key < 1)
        self.cur_golden = self.cur_key
        self.cur_altref = int(self.until_altref < 1 or self.cur_key)
        if self.until_altref < 1 or self.cur_key:
            self.until_altref = self.altref_range
        if self.cur_golden:
            self.golden_no = self.n
        if self.cur_altref:
            self.altref_no = self.n
        if self.cur_key:
            self.until_key = self.gop  # intra_transform(), src/intra_part.h:1091
        st = dict(n=self.n, key=self.cur_key, prev_golden=self.prev_golden, prev_altref=self.prev_altref,
                  altref_differs=int(self.altref_no != self.golden_no))
        self.n += 1
        return st
