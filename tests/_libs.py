"""Loaders for the test-side libraries.

  oracle()  -> ctypes handle of oracle/liboracle.so (our plain-C restatement; built on demand)
  ref()     -> ctypes handle of oracle/_ref/libOpenCL.so.1 (the REFERENCE's own kernels compiled
               for the CPU, see oracle/Makefile) or None when it has not been built
  ref_run() -> run one reference __kernel over an NDRange on numpy buffers

Nothing here is imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_oracle = None
_ref = None
_ref_tried = False

c_i16p = ctypes.POINTER(ctypes.c_int16)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f32p = ctypes.POINTER(ctypes.c_float)


class SegmentData(ctypes.Structure):
    """segment_data, src/vp8enc.h:80-92"""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "y_ac_i", "y_dc_idelta", "y2_dc_idelta", "y2_ac_idelta", "uv_dc_idelta", "uv_ac_idelta",
        "loop_filter_level", "mbedge_limit", "sub_bedge_limit", "interior_limit", "hev_threshold")]


def make_segment_data(qi=(24, 24, 24, 24), key=False, lf_level=None, sharpness=0):
    """numpy int32[4][11] filled the way prepare_segments_data() does (src/vp8enc.cpp:129-221)
    for an inter frame, with an explicit loop-filter level instead of the image-derived one."""
    dc = [4, 5, 6, 7, 8, 9, 10, 10, 11, 12, 13, 14, 15, 16, 17, 17, 18, 19, 20, 20, 21, 21, 22, 22, 23, 23, 24, 25,
          25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 46, 47, 48,
          49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 66, 67, 68, 69, 70, 71, 72, 73, 74,
          75, 76, 76, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 91, 93, 95, 96, 98, 100, 101, 102, 104,
          106, 108, 110, 112, 114, 116, 118, 122, 124, 126, 128, 130, 132, 134, 136, 138, 140, 143, 145, 148, 151,
          154, 157]
    sd = np.zeros((4, 11), np.int32)
    for s in range(4):
        sd[s, 0] = qi[s]
    sd[0, 1] = 15
    sd[0, 4] = 0 if key else -15
    sd[0, 5] = 0 if key else -15
    for s in range(4):
        lvl = lf_level[s] if lf_level is not None else min(63, dc[min(127, qi[s] + 15)] // 4)
        il = lvl
        if sharpness:
            il >>= 2 if sharpness > 4 else 1
            il = min(il, 9 - sharpness)
        il = il or 1
        sd[s, 6] = lvl
        sd[s, 7] = (lvl + 2) * 2 + il
        sd[s, 8] = lvl * 2 + il
        sd[s, 9] = il
        sd[s, 10] = 0 if key else (3 if lvl >= 40 else 2 if lvl >= 20 else 1 if lvl >= 15 else 0)
    return sd


def oracle():
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "vp8_oracle.c")
        srcs = [src, os.path.join(ORACLE_DIR, "vp8_oracle_intra.c"), os.path.join(ORACLE_DIR, "vp8_oracle.h")]
        if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(x) for x in srcs):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
        _oracle = ctypes.CDLL(so)
        _oracle.vp8o_ctx_create.restype = ctypes.c_void_p
    return _oracle


def ref_intra():
    """ctypes handle of oracle/_ref/libref_intra.so (the reference's own intra path, oracle/ref_intra.cpp) or None"""
    if os.environ.get("VP8_NO_REF"):
        return None
    so = os.path.join(ORACLE_DIR, "_ref", "libref_intra.so")
    if not os.path.exists(so) and os.path.isdir("/root/reference/src"):
        subprocess.call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(so) if os.path.exists(so) else None


def ref():
    global _ref, _ref_tried
    if os.environ.get("VP8_NO_REF"):  # tests/test_oracle_vs_golden.py: behave as on a machine without the reference
        return None
    if not _ref_tried:
        _ref_tried = True
        so = os.path.join(ORACLE_DIR, "_ref", "libOpenCL.so.1")
        if not os.path.exists(so) and os.path.isdir("/root/reference/src"):
            subprocess.call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)
        if os.path.exists(so):
            _ref = ctypes.CDLL(so)
            _ref.vp8ref_run_kernel.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p]
    return _ref


class RefImage(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("width", ctypes.c_int), ("height", ctypes.c_int)]


class Image:
    """marks a numpy uint8 [h, w] array that must be passed as an image2d_t"""

    def __init__(self, arr):
        assert arr.dtype == np.uint8 and arr.ndim == 2 and arr.flags["C_CONTIGUOUS"]
        self.arr = arr


def P(a, byte_offset=0):
    """void* to a numpy buffer (+ offset)"""
    return ctypes.c_void_p(a.ctypes.data + byte_offset)


def ref_run(name, global_size, local_size, args):
    """args: numpy arrays (pointer params), Image (image params), ctypes scalars, or raw c_void_p."""
    lib = ref()
    keep = []
    slots = []
    for a in args:
        if isinstance(a, np.ndarray):
            v = ctypes.c_void_p(a.ctypes.data)
        elif isinstance(a, Image):
            img = RefImage(a.arr.ctypes.data, a.arr.shape[1], a.arr.shape[0])
            keep.append(img)
            v = ctypes.c_void_p(ctypes.addressof(img))
        elif isinstance(a, int):
            v = ctypes.c_int32(a)
        elif isinstance(a, float):
            v = ctypes.c_float(a)
        else:
            v = a
        keep.append(v)
        slots.append(ctypes.cast(ctypes.pointer(v), ctypes.c_void_p))
    argv = (ctypes.c_void_p * len(slots))(*slots)
    rc = lib.vp8ref_run_kernel(name.encode(), int(global_size), int(local_size), argv)
    assert rc == 0, "reference kernel %s not found" % name


class GoldenNumpy:
    """numpy stand-in for tests/test_oracle_vs_ref.py (VP8_GOLDEN=record:<file> | check:<file>).  Those tests compare
    `np.array_equal(oracle_output, reference_output)`.  In record mode (reference available) the comparison is the real
    one and the sha256 of every REFERENCE output is stored per test; in check mode (no reference needed) the ORACLE
    output's sha256 has to equal the stored one.  The committed table is tests/golden/reference_kernels.json; the
    script that makes it is tests/golden/make_golden.py."""

    def __init__(self, real, spec, oracle_first=True):
        import atexit
        import json
        self._np = real
        self._oracle_first = oracle_first  # which argument of array_equal is the oracle's output
        self._mode, self._path = spec.split(":", 1)
        self._table = {}
        self._count = {}
        if self._mode == "check":
            self._table = json.load(open(self._path))
        else:
            atexit.register(lambda: json.dump(self._table, open(self._path, "w"), indent=0, sort_keys=True))

    def __getattr__(self, name):
        return getattr(self._np, name)

    def array_equal(self, a, b):
        import hashlib
        key = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0].split("::", 1)[-1]
        i = self._count.get(key, 0)
        self._count[key] = i + 1
        if not self._oracle_first:
            a, b = b, a
        if self._mode == "record":
            self._table.setdefault(key, []).append(hashlib.sha256(self._np.ascontiguousarray(b).tobytes()).hexdigest())
            return self._np.array_equal(a, b)
        want = self._table.get(key, [])
        assert i < len(want), "no golden vector %d for %s" % (i, key)
        return hashlib.sha256(self._np.ascontiguousarray(a).tobytes()).hexdigest() == want[i]
