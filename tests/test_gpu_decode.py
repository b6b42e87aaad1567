"""End-to-end on the B200 box (-m gpu): the libHMDec_* drop-in (frontend/_build/libHMDecoder_b200.so = HM's parser +
record emitter + CUDA engine) decodes the golden bitstreams; every picture's SEI MD5 check must pass in-process and the
printed MD5s must equal the ones the unmodified TAppDecoder printed for the same stream (tests/golden/*.md5)."""
import os
import re
import subprocess
import pytest
from conftest import GOLDEN, GPU_STREAMS as STREAMS, ROOT, hm_digests

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "frontend", "_build", "hmdec_cli")
CLI_VERIFY = os.path.join(ROOT, "frontend", "_build", "hmdec_cli_verify")   # links libHMDecoder_b200_verify.so (HM's CPU reconstruction next to the engine)


def _md5s(text):
    """(poc, kind, digests) of every picture whose hash check passed: MD5, CRC and checksum SEI methods alike."""
    return hm_digests(text)


@pytest.mark.parametrize("name", STREAMS)
def test_dropin_decoder_matches_tappdecoder(name):
    if not os.path.exists(CLI):
        pytest.skip("frontend/_build/hmdec_cli not built (needs the reference sources at build time)")
    r = subprocess.run([CLI, "-b", os.path.join(GOLDEN, name + ".bin"), "--touch-planes"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "***ERROR***" not in r.stdout
    got = _md5s(r.stdout)
    ref = _md5s(open(os.path.join(GOLDEN, name + ".md5")).read())
    assert got == ref and len(ref) > 0


def test_dropin_decoder_selfcheck_against_hm_cpu_recon():
    """The verify build (libHMDecoder_b200_verify.so; the product library holds no CPU reconstruction) with HMDEC_B200_VERIFY=1
    also runs HM's CPU reconstruction and compares every fetched plane byte for byte."""
    if not os.path.exists(CLI_VERIFY):
        pytest.skip("frontend not built")
    env = dict(os.environ, HMDEC_B200_VERIFY="1")
    r = subprocess.run([CLI_VERIFY, "-b", os.path.join(GOLDEN, "s_ra10_240p.bin"), "--touch-planes"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("(OK)") == 17


BENCH = os.path.join(ROOT, "bench_data")
BENCH_STREAMS = ["c3_ra10_2160p", "c4_rext444_1080p", "c5_ld10_2160p_s50", "m_ra10_1080p", "f_ra10_1080p"]


@pytest.mark.parametrize("name", BENCH_STREAMS)
def test_full_size_streams_pass_the_sei_md5_check(name):
    """BASELINE.json's full sizes (2160p Main10 RA / LD, 1080p 4:4:4 12-bit): every picture the drop-in decodes must pass
    the encoder-embedded SEI MD5 (the same check the unmodified TAppDecoder performs), and, where the reference's
    printed MD5 list is shipped, equal it."""
    path = os.path.join(BENCH, name + ".bin")
    if not os.path.exists(CLI) or not os.path.exists(path):
        pytest.skip("frontend or bench stream not present")
    r = subprocess.run([CLI, "-b", path, "--touch-planes"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "***ERROR***" not in r.stdout and "(unk)" not in r.stdout
    got = _md5s(r.stdout)
    assert len(got) >= 4
    ref_file = os.path.join(BENCH, name + ".md5")
    if os.path.exists(ref_file):
        assert got == _md5s(open(ref_file).read())


CORPUS = os.path.join(ROOT, "corpus")


def test_config1_1080p_main8_random_access_64_pictures():
    """BASELINE.json configs[1]: 1920x1080 Main 8-bit random access, 64 pictures (corpus/c2_ra8_1080p)."""
    path = os.path.join(CORPUS, "c2_ra8_1080p.bin")
    if not os.path.exists(CLI) or not os.path.exists(path):
        pytest.skip("frontend or corpus stream not present")
    r = subprocess.run([CLI, "-b", path, "--touch-planes"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "***ERROR***" not in r.stdout and "(unk)" not in r.stdout
    got = _md5s(r.stdout)
    assert len(got) == 64 and got == _md5s(open(os.path.join(CORPUS, "c2_ra8_1080p.md5")).read())


def test_config4_eight_low_delay_streams_concurrently():
    """BASELINE.json configs[4]: 8 independent 2160p Main10 low-delay-B streams decoded CONCURRENTLY on one GPU (8 processes
    here; the threads-in-one-process form is what bench.py's e2e runs).  Every picture of every stream must carry the MD5
    the unmodified TAppDecoder printed (corpus/c5_ld10_2160p_s5?.md5)."""
    names = ["c5_ld10_2160p_s%d" % k for k in range(50, 58)]
    if not os.path.exists(CLI) or not all(os.path.exists(os.path.join(CORPUS, n + ".bin")) for n in names):
        pytest.skip("frontend or corpus streams not present")
    procs = [subprocess.Popen([CLI, "-b", os.path.join(CORPUS, n + ".bin"), "--touch-planes"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for n in names]
    for n, pr in zip(names, procs):
        out, err = pr.communicate(timeout=600)
        assert pr.returncode == 0, (n, err[-2000:])
        assert "***ERROR***" not in out and "(unk)" not in out, n
        got = _md5s(out)
        assert len(got) == 17 and got == _md5s(open(os.path.join(CORPUS, n + ".md5")).read()), n


def test_packed_output_equals_tappdecoder_o(tmp_path):
    """Output wire format: `hmdec_cli -o --packed [-d N]` = conformance-window crop + bit-depth conversion +
    8/16-bit packing on the GPU (hmr_read_packed) must give the very file `TAppDecoder -o [-d N]` writes.  The stream is
    202x134 10-bit (coded 208x136, non-zero conformance window); tests/golden/s_crop10.yuvmd5 holds the reference MD5s."""
    import hashlib
    if not os.path.exists(CLI):
        pytest.skip("frontend not built")
    want = dict(l.split() for l in open(os.path.join(GOLDEN, "s_crop10.yuvmd5")))
    for depth, key, size in ((0, "d0", 202 * 134 * 3 * 5), (8, "d8", 202 * 134 * 3 // 2 * 5)):
        out = str(tmp_path / f"o{depth}.yuv")
        r = subprocess.run([CLI, "-b", os.path.join(GOLDEN, "s_crop10.bin"), "-o", out, "--packed", "-d", str(depth)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        data = open(out, "rb").read()
        assert len(data) == size
        assert hashlib.md5(data).hexdigest() == want[key], f"-d {depth}"
