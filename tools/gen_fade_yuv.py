#!/usr/bin/env python
"""Synthetic 8-bit 4:2:0 YUV with a brightness fade (so that the encoder's explicit weighted prediction picks non-trivial
weights and offsets).  Used for tests/golden/s_wpp_240p and s_wpb_240p (see tools/make_corpus.sh).
usage: gen_fade_yuv.py out.yuv [--width 416 --height 240 --frames 9 --seed 23]"""
import argparse
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("out")
ap.add_argument("--width", type=int, default=416)
ap.add_argument("--height", type=int, default=240)
ap.add_argument("--frames", type=int, default=9)
ap.add_argument("--seed", type=int, default=23)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
W, H = a.width, a.height
yy, xx = np.mgrid[0:H, 0:W]
base = 110 + 60 * np.sin(xx / 23.0) * np.cos(yy / 17.0) + 25 * np.sin((xx + yy) / 7.0)
tex = rng.normal(0, 6, (H, W))
cb = 128 + 30 * np.sin(xx[::2, ::2] / 31.0)
cr = 128 + 30 * np.cos(yy[::2, ::2] / 29.0)
with open(a.out, "wb") as f:
    for i in range(a.frames):
        g = 0.35 + 0.65 * i / max(1, a.frames - 1)            # fade in
        sh = 2 * i                                             # slow pan
        y = np.clip(np.roll(base + tex, sh, axis=1) * g + 12 * (1 - g), 0, 255).astype(np.uint8)
        u = np.clip(128 + (np.roll(cb, sh // 2, axis=1) - 128) * g, 0, 255).astype(np.uint8)
        v = np.clip(128 + (np.roll(cr, sh // 2, axis=1) - 128) * g, 0, 255).astype(np.uint8)
        f.write(y.tobytes()); f.write(u.tobytes()); f.write(v.tobytes())
