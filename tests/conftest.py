import os
import re
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
STREAMS = ["c1_intra8_240p", "s_ra8_240p", "s_ra8_240p_q22", "s_ra10_240p", "s_ld10_240p", "s_ldp8_240p",
           "s_intra10_240p_q22", "s_rext444_240p", "s_ra8_odd", "s_ra422_240p", "s_sl8_240p", "s_pcm_240p", "s_lossless_240p", "s_wpp_240p", "s_wpb_240p", "s_tiles_240p", "s_wavefront_240p", "s_cip_240p", "s_ctu32_240p", "s_ctu16_240p", "s_dqp_240p", "s_ra444_240p", "s_nolf_240p", "s_tiles2_240p",
           "s_crc_240p", "s_cksum_240p", "s_slseg_240p", "s_switch_240p", "s_gray400_240p", "s_cra_240p", "s_seek_240p"]
# Partition units other than 4x4 (log2_min_luma_transform_block_size 3 / 4 / 5: HM's per-partition arrays are per 8 / 16 / 32 samples).
# Pinned like every other stream: oracle == HM at all three stages, records byte-identical on the fast path, internals == the reference
# wrapper (CPU); engine == HM's stage MD5s and drop-in == TAppDecoder's digests on a B200 (profiles/r05_units_gpu.log).
STREAMS_UNITS = ["s_mintu8_240p", "s_ramintu8_240p", "s_mintu16_240p", "s_mintu32_240p"]
ALL_STREAMS = STREAMS + STREAMS_UNITS
GPU_STREAMS = ALL_STREAMS


_HASH_LINE = re.compile(r"POC\s+(-?\d+).*?\[(MD5|CRC|Checksum):([0-9a-f]+)(?:,([0-9a-f]+),([0-9a-f]+))?,\(OK\)\]")


def hm_digests(text):
    """[(poc, kind, (hexY, hexCb, hexCr))] from TAppDecoder-style status lines; kind is "MD5", "CRC" (SEI method 2) or "Checksum" (3).
    A 4:0:0 picture has one digest: (hexY,)."""
    return [(int(m.group(1)), m.group(2), tuple(g for g in (m.group(3), m.group(4), m.group(5)) if g is not None)) for m in _HASH_LINE.finditer(text)]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
