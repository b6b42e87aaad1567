#!/usr/bin/env python
"""bench.py — decoded frames/s of the HEVC reconstruction hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...   # the unmodified HM TAppDecoder on the host cores

Workload (config.workload): BASELINE.json configs[2] — 3840x2160 Main10 random-access stream (33 pictures, deblocking
and SAO on) produced offline by the reference's own TAppEncoder from seeded synthetic YUV (tools/make_corpus.sh).

One "step" = every stream of this rank reconstructs the whole 33-picture sequence once.
  value : frames/s with the per-picture records already RESIDENT in HBM (hmr_upload_frame / hmr_run_resident), timed
          with CUDA events on the engines' streams; several independent streams per GPU (config.streams_per_gpu) keep the
          GPU busy — a single bitstream does not shard.  N GPUs = N ranks, each with its own streams (weak).
  e2e   : the same metric through the reference-facing API: libHMDec_push_nal_unit / libHMDec_get_picture /
          libHMDEC_get_image_plane on the Annex-B BYTES of the stream (host CABAC parse -> pinned H2D of the records ->
          kernels -> D2H of every output plane), decoder threads on all host cores, SEI MD5 check on as in the reference.
  roofline / kernels : per-kernel CUDA-event durations of a single-stream pass, against algorithmic bytes.
  cpu_baseline : oracle/_ref/TAppDecoderStatic (the reference itself), one process per host core, same stream.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "c3_ra10_2160p"
DATA = os.path.join(ROOT, "bench_data")
TAPPDEC = os.path.join(ROOT, "oracle", "_ref", "TAppDecoderStatic")
CLI = os.path.join(ROOT, "frontend", "_build", "hmdec_mt")


def _paths(name):
    return os.path.join(DATA, name + ".hmr.gz"), os.path.join(DATA, name + ".bin")


def _rank_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = sorted(int(float(r[0])) for r in rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(rows[0][1])) if rows else None,
                "samples": len(rows), "reasons": sorted(reasons)}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel (profiles/*_full_raw.csv,
    produced by tools/profile_run.sh); None when no capture is shipped.  Returns (bytes, file name)."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{kernel}_kernel_full_raw.csv")))
    if not files:
        return None, None
    rows = list(csv.reader(open(files[-1])))
    if len(rows) < 3:
        return None, None
    hdr, units, vals = rows[0], rows[1], rows[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if name not in hdr:
            return None, None
        i = hdr.index(name)
        total += float(vals[i].replace(",", "")) * scale.get(units[i], 1.0)
    return total, os.path.basename(files[-1])


def algorithmic_bytes(fr):
    """Algorithmic bytes one launch of each kernel moves for this picture (definitions: DESIGN.md §d / SURVEY.md §8d)."""
    import numpy as np
    h = fr.h
    fmt = int(h["chroma_format"])
    cf = {1: 0.5, 2: 1.0, 3: 2.0}[fmt]                      # chroma samples per luma sample (both planes)
    S_b = 2.0 * int(h["width"]) * int(h["height"]) * (1 + cf)
    w4, h4 = (int(h["width"]) + 3) // 4, (int(h["height"]) + 3) // 4
    w8, h8 = (int(h["width"]) + 7) // 8, (int(h["height"]) + 7) // 8
    pu, tu, it = fr.pu, fr.tu, fr.intra
    out = {}
    if len(pu):
        area = pu["w"].astype(np.int64) * pu["h"].astype(np.int64)
        nl = (pu["lists"] & 1) + ((pu["lists"] >> 1) & 1)
        out["mc"] = 2.0 * (1 + cf) * float((area * nl).sum() + area.sum()) + 16.0 * len(pu)
    else:
        out["mc"] = 0.0
    if len(tu):
        area = (1 << (2 * tu["log2_size"].astype(np.int64)))
        coded = (tu["flags"] & 1) != 0
        inter = (tu["flags"] & 2) == 0
        out["resid"] = 2.0 * float(area[coded].sum()) + 20.0 * len(tu) + 2.0 * float(area[~inter].sum()) + 4.0 * float(area[inter].sum())
    else:
        out["resid"] = 0.0
    if len(it):
        n = (1 << it["log2_size"].astype(np.int64))
        has_res = it["resid_off"] != 0xFFFFFFFF
        out["intra"] = 2.0 * float((n * n).sum()) + 2.0 * float((n * n)[has_res].sum()) + 2.0 * float((4 * n + 1).sum()) + 16.0 * len(it)
    else:
        out["intra"] = 0.0
    # deblocking: what a pass must READ — the 4 lines x 8 samples around every edge segment with BS > 0 on the 8x8 grid, the
    # chroma lines of BS-2 segments on the chroma grid, and the BS map itself (writes are data dependent and not counted)
    out["deblock_v"] = out["deblock_h"] = 0.0
    if (int(h["flags"]) & 2) and fr.bs is not None and len(fr.bs):
        bs = np.asarray(fr.bs).reshape(h4, w4)
        sx, sy = (1, 1) if fmt == 1 else ((1, 0) if fmt == 2 else (0, 0))
        for name, seg, grid_sel, nlines in (("deblock_v", bs[:, ::2] & 3, np.arange(0, w4, 2) % (2 << sx) == 0, 4 >> sy),
                                            ("deblock_h", (bs[::2, :] >> 2) & 3, np.arange(0, h4, 2) % (2 << sy) == 0, 4 >> sx)):
            strong = seg == 2
            n_chroma = int(strong[:, grid_sel].sum()) if name == "deblock_v" else int(strong[grid_sel, :].sum())
            out[name] = 64.0 * int((seg != 0).sum()) + 2 * nlines * 8.0 * n_chroma + float(w4 * h4)
    out["sao"] = (2.0 * S_b + 36.0 * int(h["n_ctu"])) if (int(h["flags"]) & 4) else 0.0     # pictures without SAO launch nothing
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_many(cmds, env=None):
    """Run the commands concurrently (one process each); returns wall seconds and the list of return codes."""
    t0 = time.perf_counter()
    procs = [subprocess.Popen(c, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env) for c in cmds]
    rcs = [p.wait() for p in procs]
    return time.perf_counter() - t0, rcs


def count_frames(bitstream_records):
    from libhm_b200 import records
    return len(records.read_dump(bitstream_records))


def reference_pass(bitstream, nproc, passes=1):
    """One process of the unmodified TAppDecoder per host core, each decoding the whole stream `passes` times."""
    if not os.path.exists(TAPPDEC):
        return None
    wall = 0.0
    for _ in range(passes):
        cmds = [["taskset", "-c", str(i), TAPPDEC, "-b", bitstream, "-d", "0"] for i in range(nproc)]
        w, rcs = run_many(cmds)
        if any(rcs):
            raise RuntimeError(f"TAppDecoderStatic failed: {rcs}")
        wall += w
    return wall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=8, help="independent bitstreams reconstructed concurrently per GPU")
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything else any library writes to fd 1 (NCCL's version banner, for one) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    rank, local_rank, world = _rank_env()
    rec_path, bin_path = _paths(a.workload)
    ncores = os.cpu_count() or 1
    cfg = {"workload": f"{a.workload}: 3840x2160 Main10 random-access (encoder_randomaccess_main10.cfg, QP32, deblock+SAO on), 33 pictures, synthetic YUV seed 3"
           if a.workload == WORKLOAD else a.workload}

    # ---------------------------------------------------------------- reference arm: HM's own CPU decoder
    if a.impl == "reference":
        if rank != 0:
            return 0
        from libhm_b200 import records
        nframes = len(records.read_dump(rec_path))
        for _ in range(min(a.warmup, 1)):
            reference_pass(bin_path, ncores)
        wall = reference_pass(bin_path, ncores, passes=a.steps)
        fps = nframes * ncores * a.steps / wall
        line = {"impl": "reference", "metric": "decoded frames/s", "value": round(fps, 3), "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": min(a.warmup, 1), "ms_per_step": round(1000 * wall / a.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int16", "data": "synthetic", "config": dict(cfg, note="unmodified TAppDecoderStatic -d 0 (SEI MD5 check on), one process per host core, all cores; 1 warm-up pass at most"),
                "cpu_baseline": {"value": round(fps, 3), "unit": "frames/s", "cores": ncores, "kind": "reference", "sample": f"{a.steps} pass(es) of the {nframes}-picture stream per core"},
                "e2e": {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ---------------------------------------------------------------- our arm
    import numpy as np
    import torch
    import torch.distributed as dist
    from libhm_b200 import engine, records, sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the reconstruction engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    frames = records.read_dump(rec_path)
    F = len(frames)
    S = a.streams
    engines = [engine.Engine(local_rank) for _ in range(S)]
    handles = [[e.upload(f) for f in frames] for e in engines]
    rec_bytes = sum(f.nbytes() for f in frames)
    plane_bytes = sum(2 * w * h for (w, h) in (frames[0].comp_size(c) for c in range(3)))

    def step():
        for e, hs in zip(engines, handles):
            e.run_resident_list(hs)

    def barrier():
        for e in engines:
            e.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    # parity guard inside the bench: the last picture of every stream must carry HM's MD5
    for e in engines:
        got = e.read_picture(int(frames[-1].h["out_slot"]))
        assert (records.picture_md5(got, [frames[-1].bit_depth(c) for c in range(3)]) == frames[-1].gold[2]).all(), "bench: GPU picture != HM golden MD5"
    for e in engines:
        e.stage_times()
    sampler = ClockSampler(local_rank)
    barrier()
    engines[0].timer_begin()
    for _ in range(a.steps):
        step()
    for e in engines[1:]:
        engines[0].timer_join(e)
    ms = engines[0].timer_end()
    barrier()
    clocks = sampler.stop()
    launches = sum(e.stage_times()[2] for e in engines)
    frames_total, ms = sharding.reduce_measurement(S * F * a.steps, ms, device="cuda")     # frames summed, time = max over ranks
    value = sharding.frames_per_second(frames_total, ms)

    # ---- per-kernel durations, single stream, CUDA events around every launch (live, same process)
    kern = {}
    roof = None
    if rank == 0:
        e0 = engines[0]
        e0.enable_timing(True)
        e0.stage_times()
        reps = max(3, min(a.steps, 10))
        for _ in range(reps):
            e0.run_resident_list(handles[0])
        e0.sync()
        t, nf, nl = e0.stage_times()
        e0.enable_timing(False)
        alg = {}
        for f in frames:
            for k, v in algorithmic_bytes(f).items():
                alg[k] = alg.get(k, 0.0) + v
        peak, peak_src = _peaks()
        tot = sum(v for k, v in t.items() if k != "h2d")
        for k in ("mc", "resid", "intra", "deblock_v", "deblock_h", "sao"):
            dur_ms = t[k] / reps                     # per pass over the F pictures
            gbs = alg[k] / (dur_ms * 1e-3) / 1e9 if dur_ms > 0 else 0.0
            kern[k] = {"us_per_picture": round(1000 * dur_ms / F, 2), "share": round(t[k] / tot, 4), "alg_MB_per_picture": round(alg[k] / F / 1e6, 3),
                       "achieved_GBs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        top = max(kern, key=lambda k: kern[k]["share"])
        traffic, traffic_src = ncu_traffic({"deblock_v": "deblock", "deblock_h": "deblock"}.get(top, top))
        roof = {"kernel": top, "bound": "hbm", "achieved": kern[top]["achieved_GBs"], "peak": peak, "unit": "GB/s", "frac": kern[top]["frac"],
                "traffic": traffic, "traffic_source": (traffic_src + ": ncu --set full, the launch of the I picture (the heaviest launch of the pass; `achieved` averages all pictures)") if traffic_src else None,
                "peak_source": peak_src, "note": "algorithmic bytes / CUDA-event duration, single-stream pass; all kernels in `kernels`; intra is bounded by its dependency chain, not by HBM (DESIGN.md K3)"}
    for e, hs in zip(engines, handles):
        for h in hs:
            e.free_resident(h)
        e.close()

    # ---- end to end through libHMDec_* on the bitstream bytes: one decoder process per host core of this rank
    e2e = None
    my_cores = sharding.host_cores_of_rank(ncores, world, local_rank)
    cores_rank = len(my_cores)
    if not a.no_e2e and os.path.exists(CLI) and os.path.exists(bin_path):
        # decoder front ends: ONE process per rank (one CUDA context, process-wide buffer pools and MD5 service), 1.5 decoder
        # threads per host core of this rank: the surplus threads fill the ~0.13 s a finishing decoder waits for its last
        # MD5 chains
        thr = cores_rank + cores_rank // 2
        nproc = 1
        passes = 3
        env = dict(os.environ, HMDEC_B200_DEVICE=str(local_rank), HMDEC_B200_QUIET="1")
        t0 = torch.tensor([time.time() + 20.0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.broadcast(t0, 0)
        start = float(t0.item())
        first_core = my_cores[0]
        cmd = [CLI, "-b", bin_path, "--threads", str(thr), "--repeat", str(passes), "--start-at", f"{start:.3f}"]
        if world > 1:
            cmd = ["taskset", "-c", f"{first_core}-{first_core + cores_rank - 1}"] + cmd
        ps = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for p in range(nproc)]
        outs = [p.communicate()[0] for p in ps]
        try:
            res = [json.loads(o.strip().splitlines()[-1]) for o in outs]
            good = all(r["failures"] == 0 for r in res) and sum(r["pictures"] for r in res) == nproc * thr * passes * F and all(p.returncode == 0 for p in ps)
            t_first, t_last = min(r["t_start"] for r in res), max(r["t_end"] for r in res)
        except Exception:
            good, t_first, t_last = False, 0.0, 0.0
        tt = torch.tensor([-t_first if good else 0.0, t_last if good else 1e30], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)          # latest end, earliest start (negated)
        wall = float(tt[1].item()) + float(tt[0].item())
        if wall < 1e20:
            e2e = {"value": round(world * nproc * thr * passes * F / wall, 3), "unit": "frames/s",
                   "h2d_bytes_per_step": int(rec_bytes), "d2h_bytes_per_step": int(plane_bytes * F),
                   "host_cores": cores_rank, "decoder_threads": thr,
                   "note": f"libHMDec_* drop-in on Annex-B bytes: 1 process x {thr} decoder threads per GPU on {cores_rank} host cores; host CABAC parse (HM) + pinned H2D of records + kernels + DMA of every output picture into the planes libHMDEC_get_image_plane returns (caller touches all 3 planes of every picture) + SEI MD5 of every picture verified (device-side chains); a new decoder per pass, {passes} passes of the {F}-picture stream per thread after one warm-up pass, common start, wall clock to the last finisher"}
        else:
            e2e = {"value": None, "unit": "frames/s", "error": "hmdec_mt failed"}

    # ---- CPU baseline: the reference decoder itself on this box's cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and os.path.exists(TAPPDEC) and os.path.exists(bin_path):
        wall = reference_pass(bin_path, ncores)
        cpu = {"value": round(F * ncores / wall, 3), "unit": "frames/s", "cores": ncores, "kind": "reference",
               "sample": f"one pass of the {F}-picture stream per core, TAppDecoderStatic -d 0 (SEI MD5 check on), one process per core"}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {"metric": "decoded frames/s", "value": round(value, 3), "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": dict(cfg, streams_per_gpu=S, pictures_per_stream=F, l2="working set (DPB + work planes of all streams) > 126 MB L2; no explicit flush",
                           parallelism=f"{world} GPU x {S} independent streams, no collective"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "kernels": kern, "cpu_baseline": cpu}
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
