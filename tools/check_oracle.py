#!/usr/bin/env python
"""Development tool: run the CPU oracle over a record dump and compare every stage with the golden
MD5s (and, if the dump holds planes, locate the first differing sample)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from libhm_b200 import records
from oracle import oracle


def check(path, verbose=False):
    frames = records.read_dump(path)
    dec = oracle.Decoder()
    bad_frames = 0
    for i, fr in enumerate(frames):
        bds = [fr.bit_depth(c) for c in range(3)]
        dec.frame(fr, oracle.STAGE_MC | oracle.STAGE_RESID | oracle.STAGE_INTRA)
        m0 = records.picture_md5(dec.work.planes, bds)
        got0 = [p.copy() for p in dec.work.planes]
        dec.frame(fr, oracle.STAGE_MC | oracle.STAGE_RESID | oracle.STAGE_INTRA | oracle.STAGE_DBV | oracle.STAGE_DBH)
        m1 = records.picture_md5(dec.work.planes, bds)
        got1 = [p.copy() for p in dec.work.planes]
        out = dec.frame(fr)
        m2 = records.picture_md5(out.planes, bds)
        oks = [(m == fr.gold[s]).all() for s, m in enumerate((m0, m1, m2))]
        if verbose or not all(oks):
            print(f"  frame {i} poc {int(fr.h['poc'])} stages ok: {oks}")
        if not all(oks):
            bad_frames += 1
            for st, pl in ((0, got0), (1, got1), (2, out.planes)):
                for c in range(3):
                    if (st, c) in fr.planes:
                        ref, got = fr.planes[(st, c)], pl[c]
                        bad = np.argwhere(ref != got)
                        if len(bad):
                            y, x = bad[0]
                            print(f"    stage {st} comp {c}: {len(bad)} bad, first at x={x} y={y} ref={ref[y, x]} got={got[y, x]}")
            break
    return len(frames), bad_frames


if __name__ == "__main__":
    rc = 0
    for p in sys.argv[1:]:
        n, bad = check(p)
        print(f"{p}: {n} frames, {'OK' if not bad else 'MISMATCH'}")
        rc |= bad != 0
    sys.exit(rc)
