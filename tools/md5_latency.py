#!/usr/bin/env python
"""Latency / throughput of the device MD5 service on the 2160p workload (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from libhm_b200 import records, engine
frames = records.read_dump("bench_data/c3_ra10_2160p.hmr.gz")[:2]
eng = engine.Engine(0)
for f in frames: eng.submit(f)
eng.sync()
slot = int(frames[1].h["out_slot"])
for n in (1, 1, 4, 7):
    t0 = time.perf_counter()
    jobs = [eng.md5_submit(slot) for _ in range(n)]
    t1 = time.perf_counter()
    for j in jobs: d = eng.md5_result(j)
    t2 = time.perf_counter()
    print(f"{n} job(s): submit {1e3*(t1-t0):.2f} ms, all digests after {1e3*(t2-t0):.1f} ms; ok={(d == frames[1].gold[2]).all()}", flush=True)
