"""ctypes binding of the CPU oracle (oracle/hm_oracle.c).  TEST INFRASTRUCTURE ONLY: import from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never from libhm_b200/."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhm_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "hm_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "hmr_records.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu99", "-I", os.path.join(_HERE, "..", "include"), src, "-o", _SO])
    return _SO


class OrcPic(C.Structure):
    _fields_ = [("plane", C.c_void_p * 3), ("width", C.c_int * 3), ("height", C.c_int * 3), ("stride", C.c_int * 3)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_reconstruct_frame.restype = C.c_int
        _lib.orc_checksum_plane.restype = C.c_uint32
        _lib.orc_crc_plane.restype = C.c_uint32
    return _lib


class Picture:
    """Three int16 planes (numpy) + the orc_pic view of them."""

    def __init__(self, sizes):
        self.planes = [np.zeros((h, w), np.int16) for (w, h) in sizes]
        self.c = OrcPic()
        for i, p in enumerate(self.planes):
            self.c.plane[i] = p.ctypes.data
            self.c.width[i], self.c.height[i], self.c.stride[i] = p.shape[1], p.shape[0], p.shape[1]


STAGE_MC, STAGE_RESID, STAGE_INTRA, STAGE_DBV, STAGE_DBH, STAGE_SAO = 1, 2, 4, 8, 16, 32
STAGE_ALL = 63


class Decoder:
    """Runs orc_reconstruct_frame over consecutive frames, keeping the DPB (16 slots)."""

    def __init__(self):
        self.dpb = [None] * 16
        self.cdpb = (OrcPic * 16)()
        self.work = None

    def frame(self, fr, stage_mask=STAGE_ALL):
        sizes = [fr.comp_size(c) for c in range(3)]
        slot = int(fr.h["out_slot"])
        if self.dpb[slot] is None or [p.shape[::-1] for p in self.dpb[slot].planes] != [tuple(s) for s in sizes]:
            self.dpb[slot] = Picture(sizes)
            self.cdpb[slot] = self.dpb[slot].c
        if self.work is None or [p.shape[::-1] for p in self.work.planes] != [tuple(s) for s in sizes]:
            self.work = Picture(sizes)
        resid = np.zeros(max(16, int(fr.h["n_coef"])), np.int16)
        d = fr.desc()
        rc = lib().orc_reconstruct_frame(C.byref(d), self.cdpb, C.byref(self.work.c), resid.ctypes.data_as(C.c_void_p), C.c_int(stage_mask))
        assert rc == 0
        self.resid = resid
        return self.dpb[slot]
