#!/usr/bin/env python
"""Several engines (CUDA streams) reconstruct the same sequence concurrently on one GPU, records resident in HBM; afterwards the last
picture of every engine — it depends on every picture before it — must carry HM's MD5.  Catches races that only show when kernels of
different bitstreams share the SMs.   usage: multistream_check.py dump.hmr.gz [--streams 8] [--steps 4]"""
import sys, os, argparse
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from libhm_b200 import records, engine

ap = argparse.ArgumentParser()
ap.add_argument("dump")
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
frames = records.read_dump(a.dump)
engs = [engine.Engine(0) for _ in range(a.streams)]
handles = [[e.upload(f) for f in frames] for e in engs]
bad = 0
for step in range(a.steps):
    for e, hs in zip(engs, handles):
        e.run_resident_list(hs)
    for e in engs:
        e.sync()
    fr = frames[-1]
    for i, e in enumerate(engs):
        got = records.picture_md5(e.read_picture(int(fr.h["out_slot"])), [fr.bit_depth(c) for c in range(3)])
        if not (got == fr.gold[2]).all():
            bad += 1
            print(f"step {step} engine {i}: last picture differs from HM's MD5")
print(f"{a.streams} streams x {a.steps} steps x {len(frames)} pictures: {'OK' if not bad else str(bad) + ' MISMATCHES'}")
sys.exit(1 if bad else 0)
