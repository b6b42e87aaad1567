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
DUMP_CLI = os.path.join(ROOT, "frontend", "_build", "hmdec_cli")


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
    cf = {0: 0.0, 1: 0.5, 2: 1.0, 3: 2.0}[fmt]                      # chroma samples per luma sample (both planes)
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


def count_pictures(bitstream):
    """Coded pictures of an Annex-B stream: VCL NAL units (type < 32) whose first_slice_segment_in_pic_flag is set."""
    data = open(bitstream, "rb").read()
    n, pos = 0, data.find(b"\x00\x00\x01")
    while pos >= 0 and pos + 5 < len(data):
        if ((data[pos + 3] >> 1) & 0x3f) < 32 and (data[pos + 5] & 0x80):
            n += 1
        pos = data.find(b"\x00\x00\x01", pos + 3)
    return n


_FRAMES = {}


def load_frames(name):
    """Per-picture records of a workload: the committed dump (bench_data/<name>.hmr.gz, with HM's golden MD5s), or — streams shipped
    as bitstream only — a records-only dump made on the spot by the drop-in's own host parser (hmdec_cli --dump: no GPU, nothing reconstructed)."""
    if name in _FRAMES:
        return _FRAMES[name]
    from libhm_b200 import records
    rec, bitstream = _paths(name)
    if os.path.exists(rec):
        fr = records.read_dump(rec)
    else:
        with tempfile.TemporaryDirectory() as td:
            out = os.path.join(td, name + ".hmr")
            subprocess.run([DUMP_CLI, "-b", bitstream, "--dump", out, "--no-hash"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=dict(os.environ, HMDEC_B200_QUIET="1"))
            fr = records.read_dump(out)
        for f in fr:
            f.gold = None
    _FRAMES[name] = fr
    return fr


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
    ap.add_argument("--no-extra", action="store_true", help="skip the further workloads (configs[4] eight LD-B streams, fractional-pan, QP27)")
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
        nframes = count_pictures(bin_path)
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
    my_cores = sharding.host_cores_of_rank(ncores, world, local_rank)
    cores_rank = len(my_cores)

    def barrier(engines):
        for e in engines:
            e.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def check_last_picture(e, frames, name):
        """Parity guard inside the bench: the last picture must carry the MD5 HM computed for it (GOLD section of the dump, or the
        line the unmodified TAppDecoder printed: bench_data/<name>.md5, decoding order)."""
        fr = frames[-1]
        got = records.picture_md5(e.read_picture(int(fr.h["out_slot"])), [fr.bit_depth(c) for c in range(3)])
        if fr.gold is not None:
            want = fr.gold[2]
        else:
            last = [l for l in open(os.path.join(DATA, name + ".md5")) if "MD5:" in l][-1]
            want = np.frombuffer(bytes.fromhex("".join(last.split("[MD5:")[1].split(",(")[0].split(","))), np.uint8).reshape(3, 16)
        assert (got == want).all(), f"bench: GPU picture != HM's MD5 ({name})"

    def resident_value(stream_names, steps, warmup, submit=False, sample_clocks=False):
        """frames/s of this rank's engines, one engine (CUDA stream, DPB) per entry of stream_names, every engine reconstructing its
        whole sequence once per step.  submit=False: records resident in HBM (hmr_run_resident).  submit=True: every picture's
        records go from host memory through hmr_submit_frame — validation, packing into the page-locked ring and the H2D copy are
        inside the timed region — fed by one host thread per engine.  Returns (frames per step, ms of the timed steps, launches, engines' frame lists)."""
        from concurrent.futures import ThreadPoolExecutor
        framesets = {n: load_frames(n) for n in set(stream_names)}
        engines = [engine.Engine(local_rank) for _ in stream_names]
        for e in engines:
            e.lib.hmr_set_validation(e.h, 1)                 # the bench's records come from this repo's own emitter
        if submit:
            descs = [[f.desc() for f in framesets[n]] for n in stream_names]
            pool = ThreadPoolExecutor(len(engines))

            def feed(i):
                e = engines[i]
                for d in descs[i]:
                    e.submit_desc(d)

            def step():
                list(pool.map(feed, range(len(engines))))
        else:
            handles = [[e.upload(f) for f in framesets[n]] for e, n in zip(engines, stream_names)]

            def step():
                for e, hs in zip(engines, handles):
                    e.run_resident_list(hs)
        for _ in range(max(warmup, 3)):
            step()
        barrier(engines)
        for e, n in zip(engines, stream_names):
            e.sizes = [framesets[n][0].comp_size(c) for c in range(3)]
            check_last_picture(e, framesets[n], n)
            e.stage_times()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier(engines)
        t_wall = time.time()
        engines[0].timer_begin()
        for _ in range(steps):
            step()
        for e in engines[1:]:
            engines[0].timer_join(e)
        ms = engines[0].timer_end()
        barrier(engines)
        launches = sum(e.stage_times()[2] for e in engines)
        if sampler:
            # nvidia-smi reports every 200 ms: keep the same load running (untimed) until the sampler has seen it for >= 1.2 s
            while time.time() - t_wall < 1.2:
                step()
                for e in engines:
                    e.sync()
            clock_box.append(sampler.stop())
        per_step = sum(len(framesets[n]) for n in stream_names)
        if submit:
            pool.shutdown()
        return per_step, ms, launches, engines, (None if submit else handles), framesets

    def close_engines(engines, handles):
        for i, e in enumerate(engines):
            if handles:
                for h in handles[i]:
                    e.free_resident(h)
            e.close()

    def e2e_run(bins, thr, passes, extra_args=()):
        """hmdec_mt (frontend/): ONE process per rank (one CUDA context, process-wide buffer pools and MD5 service), `thr` decoder threads,
        thread t on stream t % len(bins), through the libHMDec_* entry points on the Annex-B bytes.  Returns (fps over all ranks, result of this rank)."""
        env = dict(os.environ, HMDEC_B200_DEVICE=str(local_rank), HMDEC_B200_QUIET="1")
        t0 = torch.tensor([time.time() + 12.0 + 0.4 * thr], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.broadcast(t0, 0)
        cmd = [CLI] + [x for b in bins for x in ("-b", b)] + ["--threads", str(thr), "--repeat", str(passes), "--start-at", f"{float(t0.item()):.3f}"] + list(extra_args)
        if world > 1:
            cmd = ["taskset", "-c", f"{my_cores[0]}-{my_cores[0] + cores_rank - 1}"] + cmd
        res, good = None, False
        for attempt in range(2 if world == 1 else 1):            # one more try on a single rank (no common start time to renegotiate)
            if not bins or good:
                break
            if attempt:
                cmd[cmd.index("--start-at") + 1] = f"{time.time() + 12.0 + 0.4 * thr:.3f}"
            pr = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
            try:
                res = json.loads(pr.stdout.strip().splitlines()[-1])
                good = res["failures"] == 0 and pr.returncode == 0
            except Exception:
                res, good = None, False
            if not good:
                sys.stderr.write(f"bench: hmdec_mt failed (rc {pr.returncode}, attempt {attempt + 1}): {' '.join(cmd)}\n{pr.stdout[-500:]}\n{pr.stderr[-1500:]}\n")
        tt = torch.tensor([-res["t_start"] if good else 0.0, res["t_end"] if good else 1e30, float(res["pictures"]) if good else 0.0], device="cuda", dtype=torch.float64)
        if world > 1:
            both = tt[:2].clone()
            dist.all_reduce(both, op=dist.ReduceOp.MAX)          # latest end, earliest start (negated)
            dist.all_reduce(tt[2:], op=dist.ReduceOp.SUM)
            tt[:2] = both
        wall = float(tt[1].item()) + float(tt[0].item())
        return (float(tt[2].item()) / wall if wall < 1e20 and wall > 0 else None), res

    def reference_fps(bins):
        """The unmodified TAppDecoder, one process per host core, core i on stream i % len(bins); frames/s over all cores."""
        if not os.path.exists(TAPPDEC):
            return None
        from libhm_b200 import records as _r
        cmds = [["taskset", "-c", str(i), TAPPDEC, "-b", bins[i % len(bins)], "-d", "0"] for i in range(ncores)]
        wall, rcs = run_many(cmds)
        if any(rcs):
            return None
        pics = sum(count_pictures(bins[i % len(bins)]) for i in range(ncores))
        return {"value": round(pics / wall, 3), "unit": "frames/s", "cores": ncores, "kind": "reference", "sample": "one pass per core, TAppDecoderStatic -d 0 (SEI MD5 check on)"}

    # 2 decoder threads per host core: a decoder that has pushed its last NAL waits ~0.14 s for the MD5 chains of its last pictures (a serial
    # chain per plane, on the device); the surplus threads parse meanwhile (tools/gpu_e2e_waits.sh: 16 / 24 / 32 threads on 16 cores give
    # 569 / 742 / 808 frames/s, 820 without the hash check)
    thr = 2 * cores_rank
    have_cli = os.path.exists(CLI)

    # ---- headline: `value` = 8 copies of the stream per GPU, records resident in HBM
    S = a.streams
    clock_box = []
    per_step, ms, launches, engines, handles, framesets = resident_value([a.workload] * S, a.steps, a.warmup, sample_clocks=True)
    clocks = clock_box[0]
    frames = framesets[a.workload]
    F = len(frames)
    rec_bytes = sum(f.nbytes() for f in frames)
    plane_bytes = sum(2 * w * h for (w, h) in (frames[0].comp_size(c) for c in range(3)))
    frames_total, ms_all = sharding.reduce_measurement(per_step * a.steps, ms, device="cuda")     # frames summed, time = max over ranks
    value = sharding.frames_per_second(frames_total, ms_all)

    # ---- per-kernel durations, single stream, CUDA events around every launch (live, same process)
    kern = {}
    roof = None
    plane_sum_per_pass = None
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
                "peak_source": peak_src, "single_stream_us_per_picture": round(1000 * tot / reps / F, 1),
                "note": "algorithmic bytes / CUDA-event duration, single-stream pass; all kernels in `kernels`; intra is bounded by its dependency chain, not by HBM (DESIGN.md K3)"}
        # what a caller that reads EVERY sample must find: sum over the pictures of one pass of (sample & 0xfff), from the engine's own planes
        tot_sum = 0
        for f, h in zip(frames, handles[0]):
            e0.run_resident(h)
            tot_sum += sum(int((p.astype(np.int64) & 0xfff).sum()) for p in e0.read_picture(int(f.h["out_slot"])))
        plane_sum_per_pass = tot_sum
    close_engines(engines, handles)

    # ---- value_submit: the same job with every picture's records coming from host memory through hmr_submit_frame (H2D in the timed region)
    sub_steps = max(2, a.steps // 2)
    ps2, ms2, _, eng2, _, _ = resident_value([a.workload] * S, sub_steps, 3, submit=True)
    ft2, ms2 = sharding.reduce_measurement(ps2 * sub_steps, ms2, device="cuda")
    value_submit = {"value": round(sharding.frames_per_second(ft2, ms2), 3), "unit": "frames/s", "h2d_bytes_per_step": int(rec_bytes * S), "steps": sub_steps,
                    "note": f"records of every picture from host memory through hmr_submit_frame: validation + packing into the page-locked ring + cudaMemcpyAsync inside the timed region; {S} engines per GPU, one feeding host thread each; no parse, no D2H"}
    close_engines(eng2, None)

    # ---- end to end through libHMDec_* on the bitstream bytes
    e2e = e2e_single = e2e_read = None
    if not a.no_e2e and have_cli and os.path.exists(bin_path):
        passes = 3
        fps, res = e2e_run([bin_path], thr, passes)
        if fps:
            e2e = {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": int(rec_bytes), "d2h_bytes_per_step": int(plane_bytes * F),
                   "host_cores": cores_rank, "decoder_threads": thr,
                   "note": f"libHMDec_* drop-in on Annex-B bytes: 1 process x {thr} decoder threads per GPU on {cores_rank} host cores; host CABAC parse (HM) + pinned H2D of records + kernels + DMA of every output picture into the planes libHMDEC_get_image_plane returns + SEI MD5 of every picture verified on the GPU; {passes} passes of {F} pictures per thread after 1 warm-up pass; bytes are per stream pass"}
        else:
            e2e = {"value": None, "unit": "frames/s", "error": "hmdec_mt failed"}
        if rank == 0 and world == 1:
            # what ONE caller gets (a YUView-style player: one decoder, one thread), and the full rate with the caller reading every sample
            fps1, _ = e2e_run([bin_path], 1, 2)
            e2e_single = {"value": round(fps1, 3) if fps1 else None, "unit": "frames/s", "decoder_threads": 1, "note": "one decoder thread, one bitstream: parse-bound (HM's CABAC on one core)"}
            fpsr, resr = e2e_run([bin_path], thr, 1, ["--sum-planes"])
            ok = bool(resr) and plane_sum_per_pass is not None and int(resr["plane_sum"]) == plane_sum_per_pass * thr * 1
            e2e_read = {"value": round(fpsr, 3) if fpsr else None, "unit": "frames/s", "decoder_threads": thr, "plane_sum_matches_engine": ok,
                        "note": "as e2e, but the caller reads EVERY sample of every returned plane (64-bit sum); the sum equals the one computed from the engine's own planes"}

    # ---- CPU baseline: the reference decoder itself on this box's cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and os.path.exists(TAPPDEC) and os.path.exists(bin_path):
        wall = reference_pass(bin_path, ncores)
        cpu = {"value": round(F * ncores / wall, 3), "unit": "frames/s", "cores": ncores, "kind": "reference",
               "sample": f"one pass of the {F}-picture stream per core, TAppDecoderStatic -d 0 (SEI MD5 check on), one process per core"}

    # ---- further workloads (same method; their own e2e and reference numbers)
    extra = {}
    if not a.no_extra:
        # BASELINE.json configs[4]: EIGHT DISTINCT 2160p Main10 low-delay-B streams, 8 / 4 / 2 / 1 per GPU on 1 / 2 / 4 / 8 GPUs (a fixed job: strong scaling)
        c5 = [f"c5_ld10_2160p_s{k}" for k in range(50, 58)]
        if all(os.path.exists(os.path.join(DATA, n + ".bin")) for n in c5):
            mine = [c5[i] for i in sharding.assign_streams(len(c5), world, rank)] if world <= len(c5) else []
            ent = {"config": {"workload": "c5_ld10_2160p_s50..57: BASELINE.json configs[4], eight distinct 3840x2160 Main10 low-delay-B streams (encoder_lowdelay_main10.cfg), 17 pictures each, decoded concurrently",
                              "streams_total": len(c5), "streams_per_gpu": len(mine), "scaling": "strong"}}
            if mine:
                ps, msx, _, engx, hx, _ = resident_value(mine, a.steps, 3)
                close_engines(engx, hx)
            else:
                ps, msx = 0, 0.0
            ft, msx = sharding.reduce_measurement(ps * a.steps, msx, device="cuda")
            ent["value"] = round(sharding.frames_per_second(ft, msx), 3)
            if have_cli and not a.no_e2e:
                fps, _ = e2e_run([os.path.join(DATA, n + ".bin") for n in mine], thr, 4)
                ent["e2e"] = {"value": round(fps, 3) if fps else None, "unit": "frames/s", "decoder_threads": thr, "host_cores": cores_rank}
            if rank == 0 and world == 1 and not a.no_cpu_baseline:
                ent["reference"] = reference_fps([os.path.join(DATA, n + ".bin") for n in c5])
            extra["c5_ld10_2160p"] = ent
        if world == 1:
            for name, what in (("f_ra10_2160p", "the headline configuration on a source that pans by 2.75 / 1.25 samples per picture: 70 % of the predicted luma area uses fractional motion vectors (the headline source pans by whole samples)"),
                               ("q27_ra10_2160p", "the headline configuration at QP 27 (SURVEY.md §8d's second operating point): 4.6 MB for 33 pictures, 0.13 bit/pixel, 3.8x the bits of the QP32 stream")):
                if not os.path.exists(os.path.join(DATA, name + ".bin")):
                    continue
                ent = {"config": {"workload": f"{name}: {what}", "streams_per_gpu": S}}
                ps, msx, _, engx, hx, fsx = resident_value([name] * S, max(2, a.steps // 2), 3)
                close_engines(engx, hx)
                ent["value"] = round(sharding.frames_per_second(ps * max(2, a.steps // 2), msx), 3)
                ent["record_MB_per_picture"] = round(sum(f.nbytes() for f in fsx[name]) / len(fsx[name]) / 1e6, 3)
                if have_cli and not a.no_e2e:
                    fps, _ = e2e_run([os.path.join(DATA, name + ".bin")], thr, 2)
                    ent["e2e"] = {"value": round(fps, 3) if fps else None, "unit": "frames/s", "decoder_threads": thr, "host_cores": cores_rank}
                if not a.no_cpu_baseline:
                    ent["reference"] = reference_fps([os.path.join(DATA, name + ".bin")])
                extra[name] = ent

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {"metric": "decoded frames/s", "value": round(value, 3), "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": round(ms_all / a.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": dict(cfg, streams_per_gpu=S, pictures_per_stream=F, l2="working set (DPB + work planes of all streams) > 126 MB L2; no explicit flush",
                           parallelism=f"{world} GPU x {S} independent streams, no collective"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "kernels": kern, "cpu_baseline": cpu,
            "value_submit": value_submit, "e2e_single_stream": e2e_single, "e2e_planes_read": e2e_read, "extra": extra}
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
