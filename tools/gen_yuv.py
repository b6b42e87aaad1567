#!/usr/bin/env python
"""Seeded synthetic planar YUV generator (SURVEY.md §8d recipe).

Content = 4 octaves of value noise (cells 64/16/4/1 px, weights .45/.30/.15/.10) scaled to the
legal video range, panned (+3,+2) px/frame (--pan: other, also fractional, velocities), plus fresh N(0,3) noise per frame and six inverted
160x160 squares moving at distinct velocities.  That forces intra blocks, bi-prediction, all TU
sizes and non-zero motion vectors; smooth gradients compress to nothing and exercise nothing.

Output: planar Y,Cb,Cr per frame; 8-bit -> 1 byte/sample, >8-bit -> little-endian uint16
(TVideoIOYuv.cpp:109-150 file format).
"""
import argparse
import numpy as np


def _octave(rng, h, w, cell):
    gh, gw = h // cell + 3, w // cell + 3
    g = rng.random((gh, gw), dtype=np.float32)
    if cell == 1:
        return g[:h, :w]
    ys = np.arange(h, dtype=np.float32) / cell
    xs = np.arange(w, dtype=np.float32) / cell
    y0 = ys.astype(np.int32)
    x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def make_field(rng, h, w):
    f = np.zeros((h, w), np.float32)
    for cell, wt in ((64, .45), (16, .30), (4, .15), (1, .10)):
        f += wt * _octave(rng, h, w, cell)
    return f


def generate(path, width, height, frames, bitdepth, seed, chroma="420", pan=(3.0, 2.0)):
    rng = np.random.default_rng(seed)
    # canvas big enough for the pan
    ch, cw = height + int(np.ceil(pan[1] * frames)) + 8, width + int(np.ceil(pan[0] * frames)) + 8
    luma = make_field(rng, ch, cw)
    cb = make_field(rng, ch, cw)
    cr = make_field(rng, ch, cw)
    scale = 1 << (bitdepth - 8)
    lo, hi = 16 * scale, 235 * scale
    maxv = (1 << bitdepth) - 1
    sq = [(rng.integers(0, max(1, width - 160)), rng.integers(0, max(1, height - 160)),
           int(rng.integers(-9, 10)), int(rng.integers(-7, 8))) for _ in range(6)]
    sx = 1 if chroma == "444" else 2
    sy = 1 if chroma == "444" else (1 if chroma == "422" else 2)
    dt = np.uint8 if bitdepth <= 8 else np.dtype("<u2")
    with open(path, "wb") as fh:
        for t in range(frames):
            oy, ox = int(pan[1] * t), int(pan[0] * t)
            fy, fx = np.float32(pan[1] * t - oy), np.float32(pan[0] * t - ox)
            planes = []
            for k, fld in enumerate((luma, cb, cr)):
                p = fld[oy:oy + height, ox:ox + width]
                if fx or fy:   # sub-sample pan: bilinear resample of the canvas (forces fractional motion vectors)
                    q = fld[oy:oy + height + 1, ox:ox + width + 1]
                    p = (q[:-1, :-1] * (1 - fx) + q[:-1, 1:] * fx) * (1 - fy) + (q[1:, :-1] * (1 - fx) + q[1:, 1:] * fx) * fy
                p = p * (hi - lo) + lo
                p = p + rng.normal(0.0, 3.0 * scale / 4 if k else 3.0 * scale / 2, p.shape).astype(np.float32)
                for (qx, qy, vx, vy) in sq:
                    x = int((qx + vx * t) % max(1, width - 160))
                    y = int((qy + vy * t) % max(1, height - 160))
                    p[y:y + 160, x:x + 160] = (hi + lo) - p[y:y + 160, x:x + 160]
                p = np.clip(np.rint(p), 0, maxv)
                if k:
                    p = p[::sy, ::sx]
                planes.append(p.astype(dt))
            for p in planes[:1] if chroma == "400" else planes:      # 4:0:0: the luma plane only
                fh.write(p.tobytes())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--width", type=int, required=True)
    ap.add_argument("--height", type=int, required=True)
    ap.add_argument("--frames", type=int, required=True)
    ap.add_argument("--bitdepth", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--chroma", default="420")
    ap.add_argument("--pan", default="3,2", help="pan in samples per frame (x,y); fractional values resample the canvas")
    a = ap.parse_args()
    generate(a.out, a.width, a.height, a.frames, a.bitdepth, a.seed, a.chroma, tuple(float(v) for v in a.pan.split(",")))
