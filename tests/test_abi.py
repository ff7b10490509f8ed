"""The drop-in boundary without a GPU: both C-ABI libraries load and export exactly what include/*.h declares
(no compute calls), and the shim refuses to become a platform when there is no CUDA device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from _libs import ROOT

LIB = os.path.join(ROOT, "vp8oclenc_b200", "lib")
ENGINE = os.path.join(LIB, "libvp8b200.so")
SHIM = os.path.join(LIB, "libOpenCL.so.1")

pytestmark = pytest.mark.skipif(not (os.path.exists(ENGINE) and os.path.exists(SHIM)),
                                reason="native libraries not built (python -m vp8oclenc_b200.build)")


def declared(header, pattern):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(pattern, text)))


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_engine_exports_every_declared_entry_point():
    names = declared("vp8b200.h", r"\b(vp8b200_\w+)\s*\(")
    assert len(names) >= 30
    have = exported(ENGINE)
    missing = [n for n in names if n not in have]
    assert not missing, missing
    lib = ctypes.CDLL(ENGINE)
    for n in names:
        assert getattr(lib, n) is not None
    lib.vp8b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.vp8b200_version()


def test_shim_exports_exactly_the_opencl_calls_of_the_header():
    names = declared(os.path.join("CL", "cl.h"), r"\b(cl[A-Z]\w+)\s*\(")
    have = {s for s in exported(SHIM) if re.match(r"cl[A-Z]", s)}
    assert sorted(have) == names, (sorted(have - set(names)), sorted(set(names) - have))
    assert len(names) == 28  # the calls the reference host makes


def test_shim_has_no_cpu_fallback():
    """without a CUDA device clGetPlatformIDs reports no platform (and says why) instead of running anything"""
    code = ("import ctypes,sys; L=ctypes.CDLL(%r); n=ctypes.c_uint(7); rc=L.clGetPlatformIDs(0,None,ctypes.byref(n)); "
            "print(rc, n.value)" % SHIM)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run(["python", "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0, p.stderr
    rc, n = (int(x) for x in p.stdout.split())
    assert rc == -1 and n == 0  # CL_DEVICE_NOT_FOUND
    assert "no CPU fallback" in p.stderr
