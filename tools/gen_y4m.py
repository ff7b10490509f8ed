#!/usr/bin/env python3
"""Deterministic synthetic YUV420 clips for parity tests and benchmarks.

Integer-only numpy arithmetic with fixed PCG64 seeds, so every machine produces the same
bytes.  Content: a smooth-ish noise texture plus a diagonal gradient that pans by
(+3,+1) px/frame with wrap-around (known global motion), six textured 64x64 patches that
move on their own (known local motion, two of them in half-pel steps so the quarter-pel
search stage has work), and low-amplitude chroma so the reference's chroma-difference scene
detector (src/vp8enc.cpp:285) stays quiet.

The Y4M header carries exactly the tokens the reference parser needs
(src/init.h:1628-1729, src/encIO.h:243-248): "YUV4MPEG2 W<w> H<h> F30:1 Ip A1:1 C420" and a
literal "FRAME\n" before each frame.
"""
import argparse
import hashlib
import sys

import numpy as np


def _upsample4(small, h, w):
    """integer bilinear x4 upsample of a uint16 image to (h, w) with wrap."""
    sh, sw = small.shape
    ys = np.arange(h)
    xs = np.arange(w)
    y0 = (ys // 4) % sh
    y1 = (y0 + 1) % sh
    fy = (ys % 4).astype(np.int32)
    x0 = (xs // 4) % sw
    x1 = (x0 + 1) % sw
    fx = (xs % 4).astype(np.int32)
    s = small.astype(np.int32)
    top = s[y0][:, x0] * (4 - fx)[None, :] + s[y0][:, x1] * fx[None, :]
    bot = s[y1][:, x0] * (4 - fx)[None, :] + s[y1][:, x1] * fx[None, :]
    return (top * (4 - fy)[:, None] + bot * fy[:, None] + 8) // 16


def _texture(seed, h, w, lo, hi, grad):
    rng = np.random.Generator(np.random.PCG64(seed))
    small = rng.integers(0, 256, size=((h + 3) // 4 + 1, (w + 3) // 4 + 1), dtype=np.uint16)
    img = _upsample4(small, h, w)
    fine = rng.integers(0, 32, size=(h, w), dtype=np.int32)  # a little per-pixel detail
    yy, xx = np.mgrid[0:h, 0:w]
    img = (img * 3 + fine * 2) // 4 + ((xx + yy) * grad // (h + w)) - grad // 2
    return np.clip(img, lo, hi).astype(np.uint8)


class Clip:
    """frame(i) -> (Y, U, V) uint8 arrays."""

    PATCH = 64
    # (dx2, dy2) per frame in HALF pixels: two patches move in half-pel steps
    MOTION2 = [(10, 4), (-10, -4), (4, -8), (-4, 8), (14, 0), (5, 3)]

    def __init__(self, w, h):
        self.w, self.h = w, h
        self.base_y = _texture(0x5EED0001, h, w, 16, 235, 96)
        self.base_u = _texture(0x5EED0003, h // 2, w // 2, 0, 255, 0).astype(np.int32)
        self.base_v = _texture(0x5EED0004, h // 2, w // 2, 0, 255, 0).astype(np.int32)
        self.base_u = (128 + (self.base_u - 128) * 12 // 128).astype(np.uint8)
        self.base_v = (128 + (self.base_v - 128) * 12 // 128).astype(np.uint8)
        # patches are rendered at 2x and box-averaged, so half-pel positions are exact
        p2 = _texture(0x5EED0002, 2 * self.PATCH * 6, 2 * self.PATCH + 2, 16, 235, 64)
        self.patch2 = [p2[i * 2 * self.PATCH:(i + 1) * 2 * self.PATCH] for i in range(6)]
        rng = np.random.Generator(np.random.PCG64(0x5EED0005))
        self.start = [(int(rng.integers(0, max(1, w - self.PATCH))), int(rng.integers(0, max(1, h - self.PATCH))))
                      for _ in range(6)]

    def _patch(self, k, half_x, half_y):
        """64x64 patch k sampled at a half-pel phase (half_x, half_y in {0,1})."""
        p = self.patch2[k].astype(np.int32)
        p = p[half_y:half_y + 2 * self.PATCH - 1 + 1, half_x:half_x + 2 * self.PATCH]
        p = p[:2 * self.PATCH - (half_y), :]
        hh = (p.shape[0] // 2) * 2
        ww = (p.shape[1] // 2) * 2
        p = p[:hh, :ww]
        q = (p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + 2) // 4
        out = np.zeros((self.PATCH, self.PATCH), np.int32)
        out[:q.shape[0], :q.shape[1]] = q
        if q.shape[0] < self.PATCH:
            out[q.shape[0]:, :] = out[q.shape[0] - 1:q.shape[0], :]
        if q.shape[1] < self.PATCH:
            out[:, q.shape[1]:] = out[:, q.shape[1] - 1:q.shape[1]]
        return out.astype(np.uint8)

    def frame(self, i):
        w, h = self.w, self.h
        y = np.roll(self.base_y, (1 * i, 3 * i), axis=(0, 1)).copy()
        u = np.roll(self.base_u, ((1 * i) // 2, (3 * i) // 2), axis=(0, 1)).copy()
        v = np.roll(self.base_v, ((1 * i) // 2, (3 * i) // 2), axis=(0, 1)).copy()
        for k, (dx2, dy2) in enumerate(self.MOTION2):
            if w < 2 * self.PATCH or h < 2 * self.PATCH:
                break
            px2 = self.start[k][0] * 2 + dx2 * i
            py2 = self.start[k][1] * 2 + dy2 * i
            px, hx = (px2 // 2) % (w - self.PATCH), px2 % 2
            py, hy = (py2 // 2) % (h - self.PATCH), py2 % 2
            y[py:py + self.PATCH, px:px + self.PATCH] = self._patch(k, hx, hy)
            cu = 128 + ((k * 37) % 24) - 12
            u[py // 2:py // 2 + self.PATCH // 2, px // 2:px // 2 + self.PATCH // 2] = cu
            v[py // 2:py // 2 + self.PATCH // 2, px // 2:px // 2 + self.PATCH // 2] = 256 - cu
        return y, u, v


def write_y4m(path, w, h, frames, start=0):
    """writes frames [start, start+frames) of the clip; returns the md5 of the file."""
    clip = Clip(w, h)
    md5 = hashlib.md5()
    with open(path, "wb") as f:
        hdr = ("YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420\n" % (w, h)).encode()
        f.write(hdr)
        md5.update(hdr)
        for i in range(start, start + frames):
            y, u, v = clip.frame(i)
            blob = b"FRAME\n" + y.tobytes() + u.tobytes() + v.tobytes()
            f.write(blob)
            md5.update(blob)
    return md5.hexdigest()


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("-W", "--width", type=int, default=352)
    ap.add_argument("-H", "--height", type=int, default=288)
    ap.add_argument("-n", "--frames", type=int, default=60)
    ap.add_argument("--start", type=int, default=0, help="index of the first frame (for segments)")
    a = ap.parse_args(argv)
    print(write_y4m(a.out, a.width, a.height, a.frames, a.start))


if __name__ == "__main__":
    sys.exit(main())
