"""GPU parity (run on the B200 box with -m gpu): the CUDA engine, called through its C ABI (libhmrecon.so), against
(1) the golden MD5s recorded from HM itself after CU reconstruction, deblocking and SAO, and (2) the CPU oracle,
sample by sample, with the first differing sample reported.  Integer work: the bar is bit-exact."""
import os
import numpy as np
import pytest
from conftest import GOLDEN, GPU_STREAMS as STREAMS, ROOT, hm_digests
from libhm_b200 import records

pytestmark = pytest.mark.gpu
OUT = os.path.join(ROOT, "gpurun_out")


@pytest.fixture(scope="module")
def eng_mod():
    from libhm_b200 import engine
    return engine


def _report(name, fr, stage, got, oracle_planes):
    os.makedirs(OUT, exist_ok=True)
    lines = []
    for c in range(3):
        bad = np.argwhere(got[c] != oracle_planes[c])
        if len(bad):
            y, x = bad[0]
            lines.append(f"{name} poc {int(fr.h['poc'])} stage {stage} comp {c}: {len(bad)} samples differ, first at x={x} y={y} "
                         f"oracle={oracle_planes[c][y, x]} gpu={got[c][y, x]}; bbox x[{bad[:,1].min()},{bad[:,1].max()}] y[{bad[:,0].min()},{bad[:,0].max()}]")
    with open(os.path.join(OUT, "parity_failures.log"), "a") as fh:
        fh.write("\n".join(lines) + "\n")
    return "; ".join(lines)


@pytest.mark.parametrize("name", STREAMS)
def test_engine_matches_hm_and_oracle(name, eng_mod):
    from oracle import oracle
    frames = records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz"))
    eng = eng_mod.Engine(0)
    dec = oracle.Decoder()
    pre = eng_mod.STAGE_MC | eng_mod.STAGE_RESID | eng_mod.STAGE_INTRA
    try:
        for fr in frames:
            bds = [fr.bit_depth(c) for c in range(3)]
            slot = int(fr.h["out_slot"])
            for stage, mask in ((0, pre), (1, pre | eng_mod.STAGE_DBV | eng_mod.STAGE_DBH), (2, eng_mod.STAGE_ALL)):
                eng.set_stage_mask(mask)
                eng.submit(fr)
                got = eng.read_picture(slot) if stage == 2 else eng.read_work_picture()
                if not (records.picture_md5(got, bds) == fr.gold[stage]).all():
                    ref = dec.frame(fr, mask)
                    ref = ref.planes if stage == 2 else dec.work.planes
                    pytest.fail(_report(name, fr, stage, got, ref))
            dec.frame(fr)     # keep the oracle's DPB in step for failure reports
    finally:
        eng.close()


@pytest.mark.parametrize("name", ["s_crc_240p", "s_cksum_240p", "s_ra8_240p", "s_ra10_240p"])
def test_picture_hashes_match_hm(name, eng_mod):
    """checksum / CRC kernels (SEI hash methods 3 and 2, TComPicYuvMD5.cpp:87-175) against the digests the unmodified TAppDecoder
    printed and verified against the SEI (s_crc_240p: method 2, s_cksum_240p: method 3, every picture), and — on the two MD5 streams,
    for both methods — against the oracle functions that tests/test_oracle_golden.py pins to those same HM digests."""
    import ctypes as C
    from oracle import oracle
    frames = records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz"))
    hm = {poc: (kind, [int(h, 16) for h in hx]) for poc, kind, hx in hm_digests(open(os.path.join(GOLDEN, name + ".md5")).read()) if kind != "MD5"}
    eng = eng_mod.Engine(0)
    try:
        for fr in frames:
            eng.submit(fr)
            slot = int(fr.h["out_slot"])
            poc = int(fr.h["poc"])
            if poc in hm:
                kind, want = hm[poc]
                assert eng.picture_hash(slot, 2 if kind == "CRC" else 3) == want, (name, poc, kind)
            planes = eng.read_picture(slot)
            for kind, fn in ((3, oracle.lib().orc_checksum_plane), (2, oracle.lib().orc_crc_plane)):
                exp = [int(fn(p.ctypes.data_as(C.c_void_p), C.c_int(p.shape[1]), C.c_int(p.shape[0]), C.c_int(p.shape[1]), C.c_int(fr.bit_depth(c))))
                       for c, p in enumerate(planes)]
                assert eng.picture_hash(slot, kind) == exp, (name, poc, kind)
        assert name not in ("s_crc_240p", "s_cksum_240p") or len(hm) == len(frames)
    finally:
        eng.close()


def test_resident_replay_is_deterministic(eng_mod):
    """Records resident in HBM (hmr_upload_frame / hmr_run_resident, the bench path) give the same pictures as submit."""
    frames = records.read_dump(os.path.join(GOLDEN, "s_ra8_240p.hmr.gz"))
    eng = eng_mod.Engine(0)
    try:
        handles = [eng.upload(fr) for fr in frames]
        for rep in range(2):
            for fr, h in zip(frames, handles):
                eng.run_resident(h)
                if rep == 1:
                    got = eng.read_picture(int(fr.h["out_slot"]))
                    assert (records.picture_md5(got, [fr.bit_depth(c) for c in range(3)]) == fr.gold[2]).all()
        for h in handles:
            eng.free_resident(h)
    finally:
        eng.close()


@pytest.mark.parametrize("name", ["s_ra8_odd", "s_ra10_240p", "s_rext444_240p"])
def test_device_md5_matches_hm_golden(name, eng_mod):
    """Asynchronous device MD5 (hmr_md5_submit/result: one warp per plane chain, private copy of the picture) against the MD5
    HM itself computed for the final picture (8-bit = 1 byte/sample, 10/12-bit = 2 bytes LE; odd sizes exercise the tail)."""
    frames = records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz"))
    eng = eng_mod.Engine(0)
    try:
        jobs = []
        for fr in frames:
            eng.submit(fr)
            jobs.append((eng.md5_submit(int(fr.h["out_slot"])), fr))
            if len(jobs) == 6:                      # several digests in flight while later pictures overwrite the DPB slots
                for job, f in jobs:
                    assert (eng.md5_result(job) == f.gold[2]).all(), (name, int(f.h["poc"]))
                jobs = []
        for job, f in jobs:
            assert (eng.md5_result(job) == f.gold[2]).all(), (name, int(f.h["poc"]))
    finally:
        eng.close()


@pytest.mark.parametrize("name", ["s_ra8_240p", "s_intra10_240p_q22", "s_rext444_240p"])
def test_concurrent_streams_on_one_gpu_stay_bit_exact(name, eng_mod):
    """Eight engines (CUDA streams) reconstruct the same sequence CONCURRENTLY, records resident in HBM, several passes: kernels of
    different bitstreams then share the SMs, which changes every relative timing inside the intra wavefront's CTAs (that is how a
    hand-over race was found that a single bitstream never showed).  Every picture of every engine must still carry HM's MD5."""
    frames = records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz"))
    engs = [eng_mod.Engine(0) for _ in range(8)]
    try:
        handles = [[e.upload(f) for f in frames] for e in engs]
        for rep in range(6):
            for e, hs in zip(engs, handles):
                e.run_resident_list(hs)
        for e in engs:
            e.sync()
        # picture by picture, all engines in step, so that every picture can be checked before its slot is reused
        for i, fr in enumerate(frames):
            for e, hs in zip(engs, handles):
                e.run_resident(hs[i])
            for k, e in enumerate(engs):
                got = e.read_picture(int(fr.h["out_slot"]))
                assert (records.picture_md5(got, [fr.bit_depth(c) for c in range(3)]) == fr.gold[2]).all(), (name, k, int(fr.h["poc"]))
    finally:
        for e in engs:
            e.close()


def test_concurrent_full_size_streams_stay_bit_exact(eng_mod):
    """The same at BASELINE.json's headline size (2160p Main10 random access, 8 streams x 4 passes): the last picture of every engine
    depends on every picture before it."""
    path = os.path.join(ROOT, "bench_data", "c3_ra10_2160p.hmr.gz")
    if not os.path.exists(path):
        pytest.skip("bench stream not present")
    frames = records.read_dump(path)
    engs = [eng_mod.Engine(0) for _ in range(8)]
    try:
        handles = [[e.upload(f) for f in frames] for e in engs]
        fr = frames[-1]
        for rep in range(4):
            for e, hs in zip(engs, handles):
                e.run_resident_list(hs)
            for k, e in enumerate(engs):
                got = e.read_picture(int(fr.h["out_slot"]))
                assert (records.picture_md5(got, [fr.bit_depth(c) for c in range(3)]) == fr.gold[2]).all(), (rep, k)
    finally:
        for e in engs:
            e.close()
