"""CPU test of the host side of the drop-in: HM's parser running in the product's fast configuration (picture buffers
reset instead of re-created, per-CTU coefficient zero fill skipped, emitter-side coefficient hygiene: frontend/hm_fast.cpp)
must emit records that are byte-identical to the committed golden records, which were produced with HM's stock picture
turnover and verified against HM's own reconstruction."""
import os
import subprocess
import numpy as np
import pytest
from conftest import GOLDEN, ALL_STREAMS as STREAMS, ROOT
from libhm_b200 import records

CLI = os.path.join(ROOT, "frontend", "_build", "hmdec_cli")


@pytest.mark.parametrize("name", STREAMS)
def test_fast_parse_path_emits_golden_records(name, tmp_path):
    if not os.path.exists(CLI):
        pytest.skip("frontend/_build/hmdec_cli not built (needs the reference sources at build time)")
    out = str(tmp_path / "fast.hmr")
    env = dict(os.environ, HMDUMP_RECORDS_ONLY="1", HMDEC_B200_QUIET="1")
    r = subprocess.run([CLI, "-b", os.path.join(GOLDEN, name + ".bin"), "--dump", out, "--no-hash"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode in (0, 1), r.stderr[-2000:]      # 1 = hash "mismatch": nothing was reconstructed in this mode
    got = records.read_dump(out)
    ref = records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz"))
    assert len(got) == len(ref)
    for g, f in zip(got, ref):
        for field in records.Frame.FIELDS:
            a, b = getattr(g, field), getattr(f, field)
            if a is None or b is None:
                assert (a is None or a.size == 0) and (b is None or b.size == 0), field
            else:
                assert np.array_equal(a, b), (name, int(f.h["poc"]), field)


@pytest.mark.parametrize("switch", ["HMDEC_B200_HM_BS", "HMDEC_B200_BS_FLAGS", "HMDEC_B200_BS_AT_END", "HMDEC_B200_HM_AVAIL", "HMDEC_B200_HM_COEFF", "HMDEC_B200_NO_PREFETCH", "HMDEC_B200_NO_PIC_POOL"])
@pytest.mark.parametrize("name", ["s_ra8_240p", "s_ld10_240p", "s_tiles_240p", "s_wpb_240p", "s_ramintu8_240p", "s_mintu16_240p", "s_mintu32_240p"])
def test_host_side_switches_do_not_change_the_records(name, switch, tmp_path):
    """The A/B switches of the host-side optimisations (HM's own boundary-strength routine / HM's edge-flag arrays instead of the direct rule,
    deblocking side info per picture instead of per CTU, HM's per-unit neighbour look-ups for the intra availability, HM's coefficient
    parser, no CTU prefetch, no picture pool) select other code paths
    for the SAME result: the emitted records stay byte-identical to the goldens."""
    if not os.path.exists(CLI):
        pytest.skip("frontend/_build/hmdec_cli not built (needs the reference sources at build time)")
    out = str(tmp_path / "sw.hmr")
    env = dict(os.environ, HMDUMP_RECORDS_ONLY="1", HMDEC_B200_QUIET="1")
    env[switch] = "1"
    r = subprocess.run([CLI, "-b", os.path.join(GOLDEN, name + ".bin"), "--dump", out, "--no-hash", "--repeat", "2"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode in (0, 1), r.stderr[-2000:]
    got = records.read_dump(out)
    ref = records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz"))
    assert len(got) >= len(ref)
    for g, f in zip(got, ref):
        for field in records.Frame.FIELDS:
            a, b = getattr(g, field), getattr(f, field)
            if a is None or b is None:
                assert (a is None or a.size == 0) and (b is None or b.size == 0), field
            else:
                assert np.array_equal(a, b), (name, switch, int(f.h["poc"]), field)
