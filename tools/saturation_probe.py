#!/usr/bin/env python
"""Diagnostic: resident multi-stream throughput with only some stages enabled (hmr_set_stage_mask), to see which stage
bounds `value` at saturation.  Pictures are NOT correct with stages missing — timing only.
usage: python tools/saturation_probe.py bench_data/c3_ra10_2160p.hmr.gz [--streams 8] [--steps 5]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libhm_b200 import engine, records

ap = argparse.ArgumentParser()
ap.add_argument("dump")
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
frames = records.read_dump(a.dump)
engines = [engine.Engine(0) for _ in range(a.streams)]
handles = [[e.upload(f) for f in frames] for e in engines]
names = {63: "all", 59: "all but intra", 62: "all but mc", 61: "all but resid", 39: "all but deblock", 4: "intra only", 1: "mc only", 2: "resid only", 24: "deblock only", 32: "sao only"}
for mask, name in names.items():
    for e in engines:
        e.set_stage_mask(mask)
    for _ in range(2):
        for e, hs in zip(engines, handles):
            e.run_resident_list(hs)
    for e in engines:
        e.sync()
    engines[0].timer_begin()
    for _ in range(a.steps):
        for e, hs in zip(engines, handles):
            e.run_resident_list(hs)
    for e in engines[1:]:
        engines[0].timer_join(e)
    ms = engines[0].timer_end()
    n = a.steps * a.streams * len(frames)
    print(f"{name:18s} mask {mask:2d}: {ms / n * 1000:7.1f} us per picture  ({n / ms * 1000:8.0f} pictures/s)")
