#!/usr/bin/env python
"""Replay selected pictures of a record dump on the GPU (profiling aid: keeps an ncu capture short).
usage: replay_frames.py dump.hmr[.gz] --frames 0,1 --reps 3   (all earlier pictures are run once first so the DPB is valid)"""
import sys, os, argparse
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from libhm_b200 import records, engine

ap = argparse.ArgumentParser()
ap.add_argument("dump")
ap.add_argument("--frames", default="0")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
sel = [int(x) for x in a.frames.split(",")]
frames = records.read_dump(a.dump)[: max(sel) + 1]
eng = engine.Engine(0)
handles = [eng.upload(f) for f in frames]
for h in handles:
    eng.run_resident(h)
eng.sync()
eng.enable_timing(True)
eng.stage_times()
for r in range(a.reps):
    for i in sel:
        eng.run_resident(handles[i])
eng.sync()
t, nf, nl = eng.stage_times()
print({k: round(v / max(nf, 1) * 1000, 1) for k, v in t.items()}, "us/frame over", nf, "frames")
