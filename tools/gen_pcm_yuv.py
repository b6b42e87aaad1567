#!/usr/bin/env python
"""Synthetic 8-bit 4:2:0 YUV whose left half is uniform noise (costlier to code than to send raw: the encoder picks I_PCM)
and whose right half is smooth.  Used for tests/golden/s_pcm_240p (see tools/make_corpus.sh).
usage: gen_pcm_yuv.py out.yuv [--width 416 --height 240 --frames 5 --seed 21]"""
import argparse
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("out")
ap.add_argument("--width", type=int, default=416)
ap.add_argument("--height", type=int, default=240)
ap.add_argument("--frames", type=int, default=5)
ap.add_argument("--seed", type=int, default=21)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
W, H = a.width, a.height
with open(a.out, "wb") as f:
    for i in range(a.frames):
        y = rng.integers(0, 256, (H, W), dtype=np.uint8)
        y[:, W // 2:] = (np.linspace(40, 200, W // 2)[None, :] + 10 * np.sin(np.arange(H)[:, None] / 9.0 + i)).astype(np.uint8)
        u = rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8)
        v = rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8)
        u[:, W // 4:] = 128
        v[:, W // 4:] = 120
        f.write(y.tobytes()); f.write(u.tobytes()); f.write(v.tobytes())
