"""CPU tests of bench.py's host-side helpers: picture counting on Annex-B bytes (several slice segments per picture, several coded
video sequences) and the on-the-spot record dump for workloads shipped as bitstream only."""
import os
import sys
import numpy as np
import pytest
from conftest import GOLDEN, ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("name", ["s_ra8_240p", "s_slseg_240p", "s_switch_240p", "s_tiles_240p", "c1_intra8_240p"])
def test_count_pictures_matches_the_reference_decoder(name):
    lines = [l for l in open(os.path.join(GOLDEN, name + ".md5")) if l.strip()]
    assert bench.count_pictures(os.path.join(GOLDEN, name + ".bin")) == len(lines)


def test_on_the_spot_dump_equals_the_committed_records(monkeypatch):
    if not os.path.exists(bench.DUMP_CLI):
        pytest.skip("frontend not built")
    from libhm_b200 import records
    monkeypatch.setattr(bench, "DATA", GOLDEN)
    monkeypatch.setattr(bench, "_paths", lambda n: (os.path.join(GOLDEN, n + ".nonexistent"), os.path.join(GOLDEN, n + ".bin")))
    bench._FRAMES.clear()
    got = bench.load_frames("s_ld10_240p")
    ref = records.read_dump(os.path.join(GOLDEN, "s_ld10_240p.hmr.gz"))
    assert len(got) == len(ref) and all(g.gold is None for g in got)
    for g, f in zip(got, ref):
        for field in records.Frame.FIELDS:
            a, b = getattr(g, field), getattr(f, field)
            if a is None or b is None:
                assert (a is None or a.size == 0) and (b is None or b.size == 0), field
            else:
                assert np.array_equal(a, b), field
    bench._FRAMES.clear()
