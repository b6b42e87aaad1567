"""Pins the CPU oracle (oracle/hm_oracle.c) against the reference: for every golden stream the oracle must
reproduce HM's picture after CU reconstruction, after deblocking and after SAO (MD5 per component, SEI
definition), and the final MD5 must equal the one TAppDecoder printed and verified against the SEI."""
import os
import re
import numpy as np
import pytest
import ctypes as C
from conftest import GOLDEN, ALL_STREAMS as STREAMS, hm_digests
from libhm_b200 import records
from oracle import oracle


def _tappdecoder_digests(name):
    """Per picture in decoding order: (poc, kind, [Y, Cb, Cr]) as the unmodified TAppDecoder printed (and verified against the SEI): MD5 digests as bytes,
    CRC (method 2) / checksum (method 3) as integers (TDecGop.cpp:231-289)."""
    lines = [l for l in open(os.path.join(GOLDEN, name + ".md5")) if l.strip()]
    out = [(poc, kind, [bytes.fromhex(h) for h in hx] if kind == "MD5" else [int(h, 16) for h in hx]) for poc, kind, hx in hm_digests("".join(lines))]
    assert len(out) == len(lines), name
    return out


def _oracle_hash(kind, planes, bds):
    fn = oracle.lib().orc_crc_plane if kind == "CRC" else oracle.lib().orc_checksum_plane
    return [int(fn(p.ctypes.data_as(C.c_void_p), C.c_int(p.shape[1]), C.c_int(p.shape[0]), C.c_int(p.shape[1]), C.c_int(bds[c]))) for c, p in enumerate(planes)]


@pytest.mark.parametrize("name", STREAMS)
def test_oracle_matches_hm_all_stages(name):
    frames = records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz"))
    ref = _tappdecoder_digests(name)
    assert len(frames) == len(ref)
    dec = oracle.Decoder()
    pre = oracle.STAGE_MC | oracle.STAGE_RESID | oracle.STAGE_INTRA
    for fr, (poc, kind, want) in zip(frames, ref):
        assert poc == int(fr.h["poc"])
        bds = [fr.bit_depth(c) for c in range(3)]
        dec.frame(fr, pre)
        assert (records.picture_md5(dec.work.planes, bds) == fr.gold[0]).all(), f"CU recon, POC {fr.h['poc']}"
        dec.frame(fr, pre | oracle.STAGE_DBV | oracle.STAGE_DBH)
        assert (records.picture_md5(dec.work.planes, bds) == fr.gold[1]).all(), f"deblock, POC {fr.h['poc']}"
        out = dec.frame(fr)
        md5 = records.picture_md5(out.planes, bds)
        assert (md5 == fr.gold[2]).all(), f"SAO, POC {fr.h['poc']}"
        if kind == "MD5":
            assert [bytes(md5[c]) for c in range(len(want))] == want, "final picture vs TAppDecoder/SEI MD5"      # (4:0:0: one digest)
        else:     # SEI hash methods 2 / 3: pins orc_crc_plane / orc_checksum_plane (TComPicYuvMD5.cpp:87-175) to HM's own digests
            assert _oracle_hash(kind, out.planes, bds) == want, f"final picture vs TAppDecoder/SEI {kind}, POC {fr.h['poc']}"


def test_oracle_matches_hm_on_a_concealed_lost_picture():
    """s_lost_240p = s_ra8_240p without its second coded picture (POC 8).  HM conceals it with a copy of the closest picture
    (TDecTop::xCreateLostPicture, TDecTop.cpp:233-281); the emitter sends that copy to the engine's DPB as a picture of its own (zero-vector
    PUs from the source's slot).  The oracle must reproduce HM's planes — the stand-in's and those of the 16 pictures decoded on top of it —
    at all three stages, and the final MD5s must be the ones the unmodified TAppDecoder printed for this cut (mismatching the SEI by design)."""
    frames = records.read_dump(os.path.join(GOLDEN, "s_lost_240p.hmr.gz"))
    printed = {}
    for l in open(os.path.join(GOLDEN, "s_lost_240p.md5")):
        m = re.match(r"POC\s+(-?\d+).*?\[MD5:([0-9a-f]+),([0-9a-f]+),([0-9a-f]+),", l)
        if m:
            printed[int(m.group(1))] = [bytes.fromhex(m.group(k)) for k in (2, 3, 4)]
    assert len(frames) == 17 and len(printed) == 16 and 8 not in printed
    assert int(frames[1].h["poc"]) == 8 and len(frames[1].tu) == 0 and len(frames[1].pu) > 0          # the stand-in: prediction only
    dec = oracle.Decoder()
    pre = oracle.STAGE_MC | oracle.STAGE_RESID | oracle.STAGE_INTRA
    for fr in frames:
        bds = [fr.bit_depth(c) for c in range(3)]
        dec.frame(fr, pre)
        assert (records.picture_md5(dec.work.planes, bds) == fr.gold[0]).all(), f"CU recon, POC {fr.h['poc']}"
        dec.frame(fr, pre | oracle.STAGE_DBV | oracle.STAGE_DBH)
        assert (records.picture_md5(dec.work.planes, bds) == fr.gold[1]).all(), f"deblock, POC {fr.h['poc']}"
        out = dec.frame(fr)
        md5 = records.picture_md5(out.planes, bds)
        assert (md5 == fr.gold[2]).all(), f"SAO, POC {fr.h['poc']}"
        poc = int(fr.h["poc"])
        if poc in printed:
            assert [bytes(md5[c]) for c in range(3)] == printed[poc], f"final picture vs what TAppDecoder printed, POC {poc}"
    assert (frames[1].gold[2] == frames[0].gold[2]).all()                                              # the stand-in IS picture 0


def test_fixtures_cover_the_tools():
    """The golden set must actually exercise the tools the hot path implements."""
    seen = dict(bi=0, uni=0, frac=0, dst=0, tskip=0, rdpcm=0, rotate=0, ccp=0, n32=0, sao_eo=0, sao_bo=0, bs1=0, bs2=0, strong_flag=0, planar=0, dc=0, ang=0)
    for name in STREAMS:
        for fr in records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz")):
            pu, tu, it = fr.pu, fr.tu, fr.intra
            seen["bi"] += int((pu["lists"] == 3).sum()); seen["uni"] += int((pu["lists"] != 3).sum())
            seen["frac"] += int(((pu["mv"] & 3) != 0).any(axis=(1, 2)).sum()) if len(pu) else 0
            seen["dst"] += int(((tu["flags"] & records.TU_DST) != 0).sum())
            seen["tskip"] += int(((tu["flags"] & records.TU_TSKIP) != 0).sum())
            seen["rdpcm"] += int(((tu["flags"] & (records.TU_RDPCM_H | records.TU_RDPCM_V)) != 0).sum())
            seen["rotate"] += int(((tu["flags"] & records.TU_ROTATE) != 0).sum())
            seen["ccp"] += int((tu["ccp_alpha"] != 0).sum())
            seen["n32"] += int((tu["log2_size"] == 5).sum())
            for c in range(3):
                t = fr.ctu["sao"]["type"][:, c]
                seen["sao_eo"] += int(((t >= 1) & (t <= 4)).sum()); seen["sao_bo"] += int((t == 5).sum())
            if fr.bs is not None:
                seen["bs1"] += int((((fr.bs & 3) == 1) | (((fr.bs >> 2) & 3) == 1)).sum())
                seen["bs2"] += int((((fr.bs & 3) == 2) | (((fr.bs >> 2) & 3) == 2)).sum())
            seen["strong_flag"] += int(fr.h["flags"] & records.FRM_STRONG_INTRA_SMOOTHING != 0)
            seen["planar"] += int((it["mode"] == 0).sum()); seen["dc"] += int((it["mode"] == 1).sum()); seen["ang"] += int((it["mode"] > 1).sum())
    missing = [k for k, v in seen.items() if v == 0]
    assert not missing, f"golden streams never exercise: {missing} ({seen})"


def test_dct_matrix_values():
    """Spot values of the generated HEVC core transform matrix (reference table: TComRom.cpp:335-484)."""
    import ctypes as C
    m = np.zeros((32, 32), np.int16)
    oracle.lib().orc_get_dct_matrix(C.c_int(32), m.ctypes.data_as(C.c_void_p))
    assert (m[0] == 64).all()
    assert m[1].tolist() == [90, 90, 88, 85, 82, 78, 73, 67, 61, 54, 46, 38, 31, 22, 13, 4, -4, -13, -22, -31, -38, -46, -54, -61, -67, -73, -78, -82, -85, -88, -90, -90]
    m4 = np.zeros((4, 4), np.int16)
    oracle.lib().orc_get_dct_matrix(C.c_int(4), m4.ctypes.data_as(C.c_void_p))
    assert m4.tolist() == [[64, 64, 64, 64], [83, 36, -36, -83], [64, -64, -64, 64], [36, -83, 83, -36]]
    m8 = np.zeros((8, 8), np.int16)
    oracle.lib().orc_get_dct_matrix(C.c_int(8), m8.ctypes.data_as(C.c_void_p))
    assert m8[1].tolist() == [89, 75, 50, 18, -18, -50, -75, -89]
    # every N-point matrix is orthogonal up to scaling: M M^T ~ 64*64*N/… on the diagonal, near zero elsewhere
    g = m.astype(np.int64) @ m.astype(np.int64).T
    assert (np.abs(g - np.diag(np.diag(g))) < 500).all() and (np.abs(np.diag(g) - 64 * 64 * 32) < 2000).all()
