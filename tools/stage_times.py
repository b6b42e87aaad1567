#!/usr/bin/env python
"""Replay a record dump on the GPU with per-stage CUDA-event timing (development / profiling aid).
usage: stage_times.py dump.hmr[.gz] [--reps N] [--per-frame]"""
import sys, os, time, argparse
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from libhm_b200 import records, engine

ap = argparse.ArgumentParser()
ap.add_argument("dump")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--per-frame", action="store_true")
ap.add_argument("--check", action="store_true", help="compare final pictures with the golden MD5s")
a = ap.parse_args()
frames = records.read_dump(a.dump)
eng = engine.Engine(0)
handles = [eng.upload(f) for f in frames]
for f, h in zip(frames, handles):      # warm-up pass (allocations, first-touch)
    eng.run_resident(h)
    if a.check:
        got = eng.read_picture(int(f.h["out_slot"]))
        ok = (records.picture_md5(got, [f.bit_depth(c) for c in range(3)]) == f.gold[2]).all()
        if not ok:
            print("MISMATCH at POC", int(f.h["poc"]))
eng.sync()
eng.enable_timing(True)
if a.per_frame:
    for i, (f, h) in enumerate(zip(frames, handles)):
        eng.run_resident(h)
        t, nf, nl = eng.stage_times()
        st = "IPB"[2 - int(f.h["slice_type"])] if int(f.h["slice_type"]) <= 2 else "?"
        print(f"frame {i:3d} poc {int(f.h['poc']):3d} {st} tu {int(f.h['n_tu']):6d} coef {int(f.h['n_coef']):8d} intra {int(f.h['n_intra']):6d} pu {int(f.h['n_pu']):6d} tiles {int(f.h['n_mc_tiles']):6d} | "
              + " ".join(f"{k}={v*1000:7.1f}us" for k, v in t.items() if k != "h2d") + f" | launches {nl}")
eng.stage_times()
t0 = time.perf_counter()
for r in range(a.reps):
    for h in handles:
        eng.run_resident(h)
eng.sync()
wall = time.perf_counter() - t0
t, nf, nl = eng.stage_times()
tot = sum(v for k, v in t.items() if k != "h2d")
print(f"{a.dump}: {len(frames)} frames x {a.reps} reps, wall {wall*1000:.1f} ms -> {nf/wall:.1f} fps; sum of kernel time {tot:.1f} ms ({tot/nf*1000:.1f} us/frame), {nl/nf:.1f} launches/frame")
for k, v in t.items():
    print(f"   {k:10s} {v/nf*1000:9.1f} us/frame  {100*v/max(tot,1e-9):5.1f}%")
