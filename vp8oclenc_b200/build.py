"""Builds the native parts of vp8oclenc_b200 in-tree (no JIT cache: the .so files travel with the
repo snapshot to the GPU box).

  lib/libvp8b200.so    CUDA engine, kernel-level C ABI of include/vp8b200.h      (nvcc, sm_100a)
  lib/libOpenCL.so.1   the drop-in OpenCL shim of include/CL/cl.h over the engine (nvcc, sm_100a)

nvcc cross-compiles without a GPU.  `python -m vp8oclenc_b200.build` rebuilds what is stale.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
INC = os.path.join(ROOT, "include")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", HOST_CXX, "-I", INC, "-I", CSRC]

ENGINE_SRCS = ["me_kernels.cu", "transform_kernels.cu", "loopfilter_kernels.cu", "capi_misc.cu", "engine.cu", "mb_fused_kernel.cu",
               "entropy_kernels.cu", "host_reductions.cu", "intra_kernels.cu"]
# the float SSIM must not be contracted into FMAs the reference source does not have
EXTRA = {"transform_kernels.cu": ["-fmad=false"], "mb_fused_kernel.cu": ["-fmad=false"]}
SHIM_SRCS = ["cl_shim.cu", "entropy_host.cpp"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))] + \
           [os.path.join(INC, "vp8b200.h"), os.path.join(INC, "CL", "cl.h")]


def build_engine(verbose=False, ptxas_info=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(os.path.join(PKG, "_obj"), exist_ok=True)
    objs = []
    hdrs = _headers()
    for s in ENGINE_SRCS:
        src = os.path.join(CSRC, s)
        obj = os.path.join(PKG, "_obj", s + ".o")
        if _stale(obj, [src] + hdrs) or ptxas_info:
            cmd = [NVCC] + ARCH + COMMON + EXTRA.get(s, []) + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
            _run(cmd, verbose)
        objs.append(obj)
    so = os.path.join(LIB, "libvp8b200.so")
    if _stale(so, objs):
        _run([NVCC] + ARCH + ["-shared", "-ccbin", HOST_CXX, "-o", so] + objs + ["-lcudart"], verbose)
    return so


def build_shim(verbose=False):
    engine = build_engine(verbose)
    srcs = [os.path.join(CSRC, s) for s in SHIM_SRCS]
    if not all(os.path.exists(s) for s in srcs):
        return None
    so = os.path.join(LIB, "libOpenCL.so.1")
    if _stale(so, srcs + _headers() + [engine]):
        cmd = [NVCC] + ARCH + COMMON + ["-shared", "-Xlinker", "-soname,libOpenCL.so.1", "-o", so] + srcs + \
              ["-L", LIB, "-lvp8b200", "-Xlinker", "-rpath,$ORIGIN", "-lcudart", "-lpthread"]
        _run(cmd, verbose)
        link = os.path.join(LIB, "libOpenCL.so")
        if os.path.lexists(link):
            os.remove(link)
        os.symlink("libOpenCL.so.1", link)
    return so


def build_all(verbose=False):
    e = build_engine(verbose)
    s = build_shim(verbose)
    return e, s


if __name__ == "__main__":
    v = "-q" not in sys.argv
    if "--ptxas" in sys.argv:
        build_engine(True, True)
    print(build_all(v))
