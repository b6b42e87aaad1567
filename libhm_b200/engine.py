"""ctypes binding of the B200 reconstruction engine (libhmrecon.so, C ABI in include/hmrecon.h).

There is no CPU fallback: if the shared library is missing or no CUDA device can be opened, constructing an
Engine raises.  numpy is used only to hold host buffers; all compute is in the CUDA kernels under csrc/.
"""
import ctypes as C
import os
import numpy as np
from . import records

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhmrecon.so")

STAGE_MC, STAGE_RESID, STAGE_INTRA, STAGE_DBV, STAGE_DBH, STAGE_SAO = 1, 2, 4, 8, 16, 32
STAGE_ALL = 63
STAGE_NAMES = ("h2d", "mc", "resid", "intra", "deblock_v", "deblock_h", "sao")

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.hmr_version.restype = C.c_char_p
        lib.hmr_error_string.restype = C.c_char_p
        lib.hmr_error_string.argtypes = [C.c_void_p]
        lib.hmr_alloc_pinned.restype = C.c_void_p
        lib.hmr_alloc_pinned.argtypes = [C.c_size_t]
        lib.hmr_free_pinned.argtypes = [C.c_void_p]
        for name, args in (("hmr_engine_create", [C.POINTER(C.c_void_p), C.c_int]), ("hmr_engine_destroy", [C.c_void_p]),
                           ("hmr_submit_frame", [C.c_void_p, C.c_void_p]), ("hmr_sync", [C.c_void_p]),
                           ("hmr_read_plane", [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
                           ("hmr_read_plane_async", [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
                           ("hmr_read_work_plane", [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
                           ("hmr_write_plane", [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
                           ("hmr_picture_hash", [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
                           ("hmr_set_stage_mask", [C.c_void_p, C.c_int]), ("hmr_set_validation", [C.c_void_p, C.c_int]), ("hmr_enable_timing", [C.c_void_p, C.c_int]),
                           ("hmr_get_stage_times", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
                           ("hmr_upload_frame", [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
                           ("hmr_run_resident", [C.c_void_p, C.c_void_p]), ("hmr_free_resident", [C.c_void_p, C.c_void_p]),
                           ("hmr_flush_l2", [C.c_void_p, C.c_size_t]),
                           ("hmr_md5_submit", [C.c_void_p, C.c_int, C.c_void_p]), ("hmr_md5_result", [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]),
                           ("hmr_marker_record", [C.c_void_p, C.c_void_p]), ("hmr_marker_wait", [C.c_void_p, C.c_uint64]),
                           ("hmr_run_resident_list", [C.c_void_p, C.c_void_p, C.c_int]), ("hmr_timer_begin", [C.c_void_p]),
                           ("hmr_timer_join", [C.c_void_p, C.c_void_p]), ("hmr_timer_end", [C.c_void_p, C.c_void_p])):
            getattr(lib, name).argtypes = args
        _lib = lib
    return _lib


class EngineError(RuntimeError):
    pass


class Engine:
    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.hmr_engine_create(C.byref(h), device)
        if rc != 0 or not h:
            raise EngineError(f"hmr_engine_create(device={device}) failed with {rc}: no usable CUDA device (no CPU fallback exists)")
        self.h = h
        self.sizes = None
        self.lib.hmr_set_validation(self.h, 2)      # records come from dump files here: every record is bounds-checked before a kernel sees it

    def close(self):
        if getattr(self, "h", None):
            self.lib.hmr_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self.lib.hmr_error_string(self.h).decode()}")

    def submit(self, frame):
        d = frame.desc()
        self._ck(self.lib.hmr_submit_frame(self.h, C.byref(d)), "hmr_submit_frame")
        self.sizes = [frame.comp_size(c) for c in range(3)]

    def submit_desc(self, desc):
        """hmr_submit_frame on a prepared descriptor (records.Frame.desc(); the caller keeps the frame alive)."""
        self._ck(self.lib.hmr_submit_frame(self.h, C.byref(desc)), "hmr_submit_frame")

    def sync(self):
        self._ck(self.lib.hmr_sync(self.h), "hmr_sync")

    def read_plane(self, slot, comp):
        w, h = self.sizes[comp]
        out = np.empty((h, w), np.int16)
        self._ck(self.lib.hmr_read_plane(self.h, slot, comp, out.ctypes.data, w), "hmr_read_plane")
        return out

    def read_picture(self, slot):
        return [self.read_plane(slot, c) for c in range(3)]

    def read_work_picture(self):
        out = []
        for c in range(3):
            w, h = self.sizes[c]
            a = np.empty((h, w), np.int16)
            self._ck(self.lib.hmr_read_work_plane(self.h, c, a.ctypes.data, w), "hmr_read_work_plane")
            out.append(a)
        return out

    def write_plane(self, slot, comp, arr):
        arr = np.ascontiguousarray(arr, np.int16)
        self._ck(self.lib.hmr_write_plane(self.h, slot, comp, arr.ctypes.data, arr.shape[1], arr.shape[1], arr.shape[0]), "hmr_write_plane")

    def picture_hash(self, slot, kind):
        out = (C.c_uint32 * 3)()
        self._ck(self.lib.hmr_picture_hash(self.h, slot, kind, out), "hmr_picture_hash")
        return [int(v) for v in out]

    def md5_submit(self, slot):
        job = C.c_uint64()
        self._ck(self.lib.hmr_md5_submit(self.h, slot, C.byref(job)), "hmr_md5_submit")
        return job.value

    def md5_result(self, job, wait=True):
        """3 x 16 digest bytes (uint8 [3][16]) or None while the chain is still running (wait=False)."""
        out = (C.c_uint8 * 48)()
        rc = self.lib.hmr_md5_result(self.h, job, out, int(wait))
        if rc == 1:
            return None
        self._ck(rc, "hmr_md5_result")
        return np.frombuffer(bytes(out), np.uint8).reshape(3, 16).copy()

    def set_stage_mask(self, mask):
        self._ck(self.lib.hmr_set_stage_mask(self.h, mask), "hmr_set_stage_mask")

    def enable_timing(self, on=True):
        self._ck(self.lib.hmr_enable_timing(self.h, int(on)), "hmr_enable_timing")

    def stage_times(self):
        ms = (C.c_float * 7)()
        nf, nl = C.c_uint32(), C.c_uint32()
        self._ck(self.lib.hmr_get_stage_times(self.h, ms, C.byref(nf), C.byref(nl)), "hmr_get_stage_times")
        return dict(zip(STAGE_NAMES, [float(v) for v in ms])), nf.value, nl.value

    def upload(self, frame):
        d = frame.desc()
        h = C.c_void_p()
        self._ck(self.lib.hmr_upload_frame(self.h, C.byref(d), C.byref(h)), "hmr_upload_frame")
        self.sizes = [frame.comp_size(c) for c in range(3)]
        return h

    def run_resident(self, handle):
        self._ck(self.lib.hmr_run_resident(self.h, handle), "hmr_run_resident")

    def free_resident(self, handle):
        self.lib.hmr_free_resident(self.h, handle)

    def run_resident_list(self, handles):
        arr = (C.c_void_p * len(handles))(*[h.value for h in handles])
        self._ck(self.lib.hmr_run_resident_list(self.h, arr, len(handles)), "hmr_run_resident_list")

    def timer_begin(self):
        self._ck(self.lib.hmr_timer_begin(self.h), "hmr_timer_begin")

    def timer_join(self, other):
        self._ck(self.lib.hmr_timer_join(self.h, other.h), "hmr_timer_join")

    def timer_end(self):
        ms = C.c_float()
        self._ck(self.lib.hmr_timer_end(self.h, C.byref(ms)), "hmr_timer_end")
        return float(ms.value)

    def flush_l2(self, nbytes=256 << 20):
        self._ck(self.lib.hmr_flush_l2(self.h, nbytes), "hmr_flush_l2")
