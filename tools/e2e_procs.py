#!/usr/bin/env python
"""e2e exploration: P decoder processes x T threads (P*T = cores) through the libHMDec_* drop-in, common start time."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MT = os.path.join(ROOT, "frontend", "_build", "hmdec_mt")


def run(bitstream, procs, threads, repeat, extra=(), lead=25.0):
    start = time.time() + lead
    ps = []
    for p in range(procs):
        cmd = [MT, "-b", bitstream, "--threads", str(threads), "--repeat", str(repeat), "--pin", str(p * threads), "--start-at", f"{start:.3f}", *extra]
        ps.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
    res = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in ps]
    pics = sum(r["pictures"] for r in res)
    late = max(r["t_start"] for r in res) - start
    wall = max(r["t_end"] for r in res) - min(r["t_start"] for r in res)
    return {"procs": procs, "threads": threads, "pictures": pics, "wall": round(wall, 3), "fps": round(pics / wall, 2), "failures": sum(r["failures"] for r in res), "late_start_s": round(late, 3)}


if __name__ == "__main__":
    bs = sys.argv[1]
    cores = os.cpu_count()
    for procs in (cores, cores // 2, cores // 4, 2, 1):
        print(json.dumps(run(bs, procs, cores // procs, 2)), flush=True)
    print(json.dumps(dict(run(bs, cores // 4, 4, 2, ("--no-hash",)), hash=False)), flush=True)
