"""Host-side view of the flat per-frame records (include/hmr_records.h) and of the record-dump
container written by frontend/dump_sink.cpp.  numpy structured dtypes mirror the C structs byte for
byte; `Frame.desc()` builds the `hmr_frame_desc` that both the engine C-ABI and the oracle take.
"""
import ctypes as C
import gzip
import struct
import numpy as np

HMR_MAGIC = 0x52524D48
HMR_VERSION = 4
HMR_NO_OFFSET = 0xFFFFFFFF

FRM_STRONG_INTRA_SMOOTHING, FRM_DEBLOCK, FRM_SAO, FRM_HAS_NOFILTER, FRM_HAS_CCP, FRM_IS_REFERENCE, FRM_INTRA_ONLY, FRM_SCALING_LIST, FRM_WEIGHTED_PRED = (1 << i for i in range(9))
TU_CODED, TU_INTRA, TU_DST, TU_TSKIP, TU_BYPASS, TU_ROTATE, TU_RDPCM_H, TU_RDPCM_V = (1 << i for i in range(8))

HDR_DT = np.dtype([("magic", "<u4"), ("version", "<u4"), ("width", "<i4"), ("height", "<i4"), ("poc", "<i4"),
                   ("chroma_format", "u1"), ("bit_depth_luma", "u1"), ("bit_depth_chroma", "u1"), ("log2_ctu", "u1"),
                   ("out_slot", "u1"), ("slice_type", "u1"), ("pps_cb_qp_offset", "i1"), ("pps_cr_qp_offset", "i1"),
                   ("flags", "<u4"), ("n_tu", "<u4"), ("n_coef", "<u4"), ("n_intra", "<u4"), ("n_pu", "<u4"),
                   ("n_mc_tiles", "<u4"), ("n_ctu", "<u4"), ("tu_first", "<u4", (5,)), ("reserved", "<u4", (1,))])
TU_DT = np.dtype([("x", "<u2"), ("y", "<u2"), ("comp", "u1"), ("log2_size", "u1"), ("flags", "u1"), ("qp", "u1"),
                  ("ccp_alpha", "i1"), ("pad", "u1", (3,)), ("coef_off", "<u4"), ("luma_off", "<u4")])
INTRA_DT = np.dtype([("x", "<u2"), ("y", "<u2"), ("comp", "u1"), ("log2_size", "u1"), ("mode", "u1"), ("flags", "u1"),
                     ("avail_left", "u1"), ("avail_below_left", "u1"), ("avail_above", "u1"), ("avail_above_right", "u1"),
                     ("resid_off", "<u4")])
IRNG_DT = np.dtype([("first", "<u4", (3,)), ("count", "<u4", (3,))])
PU_DT = np.dtype([("x", "<u2"), ("y", "<u2"), ("w", "u1"), ("h", "u1"), ("lists", "u1"), ("slots", "u1"), ("mv", "<i2", (2, 2))])
SAO_DT = np.dtype([("type", "u1"), ("band", "u1"), ("off", "<i2", (4,))])
WP_DT = np.dtype([("weight", "<i2"), ("offset", "<i2"), ("log2_denom", "u1"), ("pad", "u1")])
CTU_DT = np.dtype([("sao", SAO_DT, (3,)), ("avail", "u1"), ("beta_offset_div2", "i1"), ("tc_offset_div2", "i1"), ("pad", "u1", (3,))])
assert (HDR_DT.itemsize, TU_DT.itemsize, INTRA_DT.itemsize, IRNG_DT.itemsize, PU_DT.itemsize, SAO_DT.itemsize, CTU_DT.itemsize) == (80, 20, 16, 24, 16, 10, 36)


class FrameDesc(C.Structure):
    """struct hmr_frame_desc"""
    _fields_ = [(n, C.c_void_p) for n in ("hdr", "tu", "coef", "intra", "intra_range", "pu", "pu_tile_prefix", "ctu", "bs", "qp", "cu_flags", "scaling", "wp", "pu_refidx")]


_SECTIONS = {b"HDR ": ("hdr", HDR_DT), b"TU  ": ("tu", TU_DT), b"COEF": ("coef", np.dtype("<i2")), b"INTR": ("intra", INTRA_DT),
             b"IRNG": ("intra_range", IRNG_DT), b"PU  ": ("pu", PU_DT), b"PUPF": ("pu_tile_prefix", np.dtype("<u4")),
             b"CTU ": ("ctu", CTU_DT), b"BS  ": ("bs", np.dtype("u1")), b"QP  ": ("qp", np.dtype("i1")), b"CUFL": ("cu_flags", np.dtype("u1")), b"SCAL": ("scaling", np.dtype("u1")), b"WP  ": ("wp", WP_DT), b"PURI": ("pu_refidx", np.dtype("u1"))}


class Frame:
    """One picture's records (+ optional golden data recorded from HM's own CPU reconstruction)."""
    FIELDS = ("hdr", "tu", "coef", "intra", "intra_range", "pu", "pu_tile_prefix", "ctu", "bs", "qp", "cu_flags", "scaling", "wp", "pu_refidx")

    def __init__(self):
        for f in self.FIELDS:
            setattr(self, f, None)
        self.gold = None      # uint8 [3 stages][3 comps][16]  MD5 of HM's planes after CU recon / deblock / SAO
        self.planes = {}      # (stage, comp) -> int16 array, only when dumped with HMDUMP_PLANES=1

    # geometry helpers -----------------------------------------------------------------------
    @property
    def h(self):
        return self.hdr[0]

    def comp_size(self, c):
        fmt = int(self.h["chroma_format"])
        if c and fmt == 0:
            return 0, 0          # 4:0:0: no chroma planes (TComPicYuv allocates none)
        sx = 1 if (c and fmt in (1, 2)) else 0
        sy = 1 if (c and fmt == 1) else 0
        return int(self.h["width"]) >> sx, int(self.h["height"]) >> sy

    def bit_depth(self, c):
        return int(self.h["bit_depth_chroma"] if c else self.h["bit_depth_luma"])

    def desc(self):
        d = FrameDesc()
        keep = []
        for f in self.FIELDS:
            a = getattr(self, f)
            if a is None or (f in ("bs", "cu_flags", "scaling", "wp", "pu_refidx") and a.size == 0):
                setattr(d, f, None)
                continue
            a = np.ascontiguousarray(a)
            keep.append(a)
            setattr(d, f, a.ctypes.data if a.size else None)
        d._keep = keep
        return d

    def nbytes(self):
        return sum(getattr(self, f).nbytes for f in self.FIELDS if getattr(self, f) is not None)


def read_dump(path):
    """Parse a record dump (optionally gzip-compressed) into a list of Frame."""
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rb") as fh:
        blob = fh.read()
    assert blob[:8] == b"HMRDUMP1", "not a record dump"
    pos, frames, cur = 8, [], None
    while pos < len(blob):
        tag = blob[pos:pos + 4]
        (n,) = struct.unpack_from("<Q", blob, pos + 8)
        pos += 16
        payload = blob[pos:pos + n]
        pos += (n + 7) & ~7
        if tag == b"HDR ":
            cur = Frame()
            frames.append(cur)
        if tag in _SECTIONS:
            name, dt = _SECTIONS[tag]
            setattr(cur, name, np.frombuffer(payload, dtype=dt).copy())
        elif tag == b"GOLD":
            cur.gold = np.frombuffer(payload, dtype=np.uint8).reshape(3, 3, 16).copy()
        elif tag[:1] == b"P" and tag[2:3] == b"C":
            stage, comp = tag[1] - 48, tag[3] - 48
            w, h = cur.comp_size(comp)
            cur.planes[(stage, comp)] = np.frombuffer(payload, dtype="<i2").reshape(h, w).copy()
        elif tag == b"END ":
            pass
    for f in frames:
        assert int(f.h["magic"]) == HMR_MAGIC and int(f.h["version"]) == HMR_VERSION
        for name in ("tu", "coef", "intra", "pu"):
            if getattr(f, name) is None:
                setattr(f, name, np.zeros(0, dtype=_SECTIONS[{"tu": b"TU  ", "coef": b"COEF", "intra": b"INTR", "pu": b"PU  "}[name]][1]))
    return frames


def picture_md5(planes, bit_depths):
    """MD5 per component the way the SEI decoded-picture hash defines it (TComPicYuvMD5.cpp:183-205):
    1 byte per sample for bit depth <= 8, else 2 bytes little endian.  Returns uint8 [3][16]."""
    import hashlib
    out = np.zeros((3, 16), np.uint8)
    for c, (p, bd) in enumerate(zip(planes, bit_depths)):
        if p.size == 0:          # an absent component (4:0:0 chroma): all-zero digest, the convention of the dump's GOLD section
            continue
        raw = p.astype(np.uint8).tobytes() if bd <= 8 else p.astype("<u2").tobytes()
        out[c] = np.frombuffer(hashlib.md5(raw).digest(), np.uint8)
    return out
