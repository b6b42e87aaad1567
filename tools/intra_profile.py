#!/usr/bin/env python
"""Cycle accounting of the intra wavefront's chain warps (needs libhmrecon.so built with EXTRA=-DINTRA_PROFILE).
usage: intra_profile.py dump.hmr[.gz] --frames 0 [--reps 2]     prints, per TU class, count and mean cycles of prep / token wait / turn."""
import sys, os, argparse, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from libhm_b200 import records, engine

ap = argparse.ArgumentParser()
ap.add_argument("dump")
ap.add_argument("--frames", default="0")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
sel = [int(x) for x in a.frames.split(",")]
frames = records.read_dump(a.dump)[: max(sel) + 1]
eng = engine.Engine(0)
lib = engine.load()
handles = [eng.upload(f) for f in frames]
for h in handles:
    eng.run_resident(h)
eng.sync()
out = (C.c_ulonglong * 64)()
rc = lib.hmr_debug_intra_profile(out, 1)
if rc != 0:
    raise SystemExit(f"hmr_debug_intra_profile rc={rc}: rebuild with `make -f libhm_b200/csrc/Makefile EXTRA=-DINTRA_PROFILE`")
eng.enable_timing(True)
eng.stage_times()
for r in range(a.reps):
    for i in sel:
        eng.run_resident(handles[i])
eng.sync()
t, nf, nl = eng.stage_times()
print({k: round(v / max(nf, 1) * 1000, 1) for k, v in t.items()}, "us/frame over", nf, "frames")
lib.hmr_debug_intra_profile(out, 0)
v = list(out)
for lg in range(2, 6):
    b = 8 * (lg - 2)
    own, helper = v[b], v[b + 6]
    if not own:
        continue
    n = 1 << lg
    print(f"{n:2d}x{n:<2d} warp-turns {own:8d}: prep {v[b + 1] / own:7.1f}  token(nominal) {v[b + 2] / own:6.1f}  line/DC+token {v[b + 3] / own:7.1f}  turn {v[b + 5] / own:7.1f} cycles")
ctus = max(v[33], 1)
print(f"per chain warp and CTU: wait for the staged CTU {v[32] / ctus:.0f} cycles; {ctus} warp-CTUs; chain warps busy+waiting in row jobs: {v[34] / 4:.0f} warp-cycles per warp-set")
rows = (C.c_ulonglong * (3 * 128 * 4))()
if lib.hmr_debug_intra_rows(rows) == 0:
    r = list(rows)
    t0 = min(x for x in r[0::4] if x)
    print("wavefront of the last launch (us since the first job started): comp row  job-start  first-CTU-staged  last-TU-done")
    for comp in range(3):
        for row in range(0, 128):
            b = (comp * 128 + row) * 4
            if r[b] and (row < 6 or row % 8 == 0 or r[(comp * 128 + row + 1) * 4] == 0):
                print(f"  {comp} {row:3d}  {(r[b] - t0) / 1e3:9.1f} {(r[b + 1] - t0) / 1e3:9.1f} {(r[b + 2] - t0) / 1e3:9.1f}")
