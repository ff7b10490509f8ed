import os, sys, ctypes
import numpy as np, torch
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_y4m
from vp8oclenc_b200 import host as eng
w, h = 1920, 64
y, u, v = (torch.from_numpy(np.ascontiguousarray(p).reshape(-1)).cuda() for p in gen_y4m.Clip(w, 128).frame(1))
y, u, v = y[: w * h].contiguous(), u[: w * h // 4].contiguous(), v[: w * h // 4].contiguous()
M = (w // 16) * (h // 16)
out = [torch.zeros(n, dtype=t, device="cuda") for n, t in ((w * h, torch.uint8), (w * h // 4, torch.uint8), (w * h // 4, torch.uint8), (M * 400, torch.int16), (M * 16, torch.int32), (M, torch.int32), (M, torch.int32))]
for _ in range(3):
    keep = eng.intra_frame(y, u, v, *out, w, h, (19, 24, 7, 10))
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["VP8B200_ENGINE_LIB"])
t = np.zeros((4, 128, 16), dtype=np.uint64)
print(lib.vp8b200_intra_debug_timeline(t.ctypes.data_as(ctypes.c_void_p)))
t = t.astype(np.int64); t0 = t[0, 0, 0]
for r in range(4):
    for c in (5, 6, 60):
        d = np.diff(t[r, c, :15]) / 1000.0
        print(r, c, " ".join("%5.2f" % x for x in d))
ph = np.zeros((10, 8), dtype=np.int64)
print(lib.vp8b200_intra_debug_phases(ph.ctypes.data_as(ctypes.c_void_p)))
print("cycles inside the steps of macroblock (0,5): edges | predict+fdct+weight | min | winner: quantise, reconstruct, store")
for t in range(10):
    print(t, " ".join("%6d" % x for x in np.diff(ph[t, :5])), "  -> next step %6d" % ((ph[t + 1, 0] - ph[t, 4]) if t < 9 else 0))
