"""CPU test of the reference-facing wrapper (frontend/_build/libHMDecoder_b200.so): the 16 libHMDec_* entry points of
libHMDecoder.h:111-298 exist with C linkage, and the protocol / error behaviour SURVEY.md §8(b) documents holds:
length <= 0 and short non-final NALs -> READ_ERROR, NULL handles -> NULL / -1, and the push / re-push / drain loop of
libHMDecoder.h:38-77 delivers every picture in output order.  Runs on the record-only back-end (no GPU, nothing is
reconstructed), so it exercises HM's parser, the emitter and the output bumping, not the kernels."""
import ctypes as C
import os
import re
import pytest
from conftest import GOLDEN, ROOT

LIB = os.path.join(ROOT, "frontend", "_build", "libHMDecoder_b200.so")
SYMS = ["libHMDec_get_version", "libHMDec_new_decoder", "libHMDec_free_decoder", "libHMDec_set_SEI_Check", "libHMDec_set_max_temporal_layer",
        "libHMDec_push_nal_unit", "libHMDec_get_picture", "libHMDEC_get_POC", "libHMDEC_get_picture_width", "libHMDEC_get_picture_height",
        "libHMDEC_get_picture_stride", "libHMDEC_get_image_plane", "libHMDEC_get_chroma_format", "libHMDEC_get_internal_bit_depth",
        "libHMDEC_get_internal_info", "libHMDEC_clear_internal_info"]
OK, ERROR, READ_ERROR = 0, 1, 2


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.skip("frontend not built (needs the reference sources at build time)")
    os.environ["HMDEC_B200_QUIET"] = "1"
    l = C.CDLL(LIB)
    l.libHMDec_get_version.restype = C.c_char_p
    l.libHMDecB200_new_decoder_ex.restype = C.c_void_p
    l.libHMDecB200_new_decoder_ex.argtypes = [C.c_int, C.c_char_p]
    l.libHMDec_free_decoder.argtypes = [C.c_void_p]
    l.libHMDec_set_SEI_Check.argtypes = [C.c_void_p, C.c_bool]
    l.libHMDec_push_nal_unit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_bool, C.POINTER(C.c_bool), C.POINTER(C.c_bool)]   # bool& == bool*
    l.libHMDec_get_picture.restype = C.c_void_p
    l.libHMDec_get_picture.argtypes = [C.c_void_p]
    for n in ("libHMDEC_get_POC",):
        getattr(l, n).argtypes = [C.c_void_p]
    for n in ("libHMDEC_get_picture_width", "libHMDEC_get_picture_height", "libHMDEC_get_picture_stride"):
        getattr(l, n).argtypes = [C.c_void_p, C.c_int]
    l.libHMDEC_get_image_plane.restype = C.c_void_p
    l.libHMDEC_get_image_plane.argtypes = [C.c_void_p, C.c_int]
    l.libHMDEC_get_chroma_format.argtypes = [C.c_void_p]
    l.libHMDEC_clear_internal_info.argtypes = [C.c_void_p]
    return l


def _nals(path):
    data = open(path, "rb").read()
    starts = [m.start() for m in re.finditer(b"\x00\x00\x01", data)]
    out = []
    for i, s in enumerate(starts):
        e = starts[i + 1] if i + 1 < len(starts) else len(data)
        nal = data[s + 3:e]
        while i + 1 < len(starts) and nal.endswith(b"\x00"):      # trailing zero of a 4-byte start code / trailing_zero_8bits
            nal = nal[:-1]
        out.append(nal)
    return out


def test_all_sixteen_entry_points_are_exported(lib):
    for s in SYMS:
        assert hasattr(lib, s), s
    assert lib.libHMDec_get_version() == b"16.0"          # NV_VERSION of the reference (libHMDecoder.cpp:71-74)


def test_null_handles_and_bad_lengths(lib):
    assert lib.libHMDec_free_decoder(None) == ERROR
    assert lib.libHMDec_get_picture(None) is None
    assert lib.libHMDEC_get_POC(None) == -1
    assert lib.libHMDEC_get_picture_width(None, 0) == -1 and lib.libHMDEC_get_picture_height(None, 0) == -1 and lib.libHMDEC_get_picture_stride(None, 0) == -1
    assert lib.libHMDEC_get_image_plane(None, 0) is None
    assert lib.libHMDEC_get_chroma_format(None) == 4      # LIBHMDEC_CHROMA_UNKNOWN
    assert lib.libHMDEC_clear_internal_info(None) == ERROR
    dec = lib.libHMDecB200_new_decoder_ex(1, b"null")
    assert dec
    new, chk = C.c_bool(False), C.c_bool(False)
    buf = (C.c_ubyte * 8)(0x40, 0x01, 0x0c, 0x01, 0xff, 0xff, 0x01, 0x60)
    assert lib.libHMDec_push_nal_unit(dec, buf, 0, False, C.byref(new), C.byref(chk)) == READ_ERROR      # libHMDecoder.cpp:118-124
    assert lib.libHMDec_push_nal_unit(dec, buf, -3, False, C.byref(new), C.byref(chk)) == READ_ERROR
    assert lib.libHMDec_push_nal_unit(dec, buf, 3, False, C.byref(new), C.byref(chk)) == READ_ERROR      # < 4 bytes and not the last NAL
    assert lib.libHMDec_push_nal_unit(None, buf, 8, False, C.byref(new), C.byref(chk)) == ERROR
    assert lib.libHMDec_get_picture(dec) is None          # nothing decoded yet
    assert lib.libHMDec_free_decoder(dec) == OK


@pytest.mark.parametrize("name,expected", [("s_ra8_240p", 17), ("s_ld10_240p", 9), ("c1_intra8_240p", 16)])
def test_push_repush_drain_loop_delivers_all_pictures_in_output_order(lib, name, expected):
    nals = _nals(os.path.join(GOLDEN, name + ".bin"))
    dec = lib.libHMDecB200_new_decoder_ex(1, b"null")
    lib.libHMDec_set_SEI_Check(dec, False)
    pocs, k, repushed = [], 0, 0
    while k < len(nals):
        new, chk = C.c_bool(False), C.c_bool(False)
        buf = (C.c_ubyte * len(nals[k])).from_buffer_copy(nals[k])
        assert lib.libHMDec_push_nal_unit(dec, buf, len(nals[k]), k + 1 == len(nals), C.byref(new), C.byref(chk)) == OK
        if chk.value:
            while True:
                pic = lib.libHMDec_get_picture(dec)
                if not pic:
                    break
                pocs.append(lib.libHMDEC_get_POC(pic))
                assert lib.libHMDEC_get_picture_width(pic, 0) == 416 and lib.libHMDEC_get_picture_height(pic, 0) == 240
                assert lib.libHMDEC_get_picture_width(pic, 1) == 208 and lib.libHMDEC_get_chroma_format(pic) == 1
                assert lib.libHMDEC_get_picture_stride(pic, 0) >= 416 and lib.libHMDEC_get_image_plane(pic, 0)
        if new.value:
            repushed += 1                                 # the NAL was not consumed: same NAL again (libHMDecoder.h:151)
        else:
            k += 1
    assert lib.libHMDec_free_decoder(dec) == OK
    assert len(pocs) == expected and pocs == sorted(pocs) and len(set(pocs)) == expected
    assert repushed >= expected - 1


def _blocks(lib, dec, pic, t):
    """libHMDEC_get_internal_info returns a std::vector<libHMDec_BlockValue>* (libHMDecoder.h:291): begin / end pointers, 16-byte entries."""
    lib.libHMDEC_get_internal_info.restype = C.c_void_p
    lib.libHMDEC_get_internal_info.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    v = lib.libHMDEC_get_internal_info(dec, pic, t)
    assert v
    begin, end = (C.c_void_p * 2).from_address(v)
    n = ((end or 0) - (begin or 0))
    assert n % 16 == 0
    return C.string_at(begin, n) if n else b""


def test_internal_info_index_is_order_independent_and_per_picture(lib):
    """The block lists come from a per-picture index built by the first query (frontend/internals.cpp): whatever the order
    of the queries, with libHMDEC_clear_internal_info in between, and with queries for two pictures interleaved, every
    (picture, type) pair must answer the same bytes; the index must not survive into the next push."""
    nals = _nals(os.path.join(GOLDEN, "s_ra8_240p.bin"))
    dec = lib.libHMDecB200_new_decoder_ex(1, b"null")
    lib.libHMDec_set_SEI_Check(dec, False)
    order = [21, 3, 12, 0, 15, 22, 8, 1, 11, 14, 23, 5, 18, 7, 10, 2, 9, 16, 4, 13, 17, 6, 19, 20]
    k, seen, interleaved = 0, 0, 0
    while k < len(nals):
        new, chk = C.c_bool(False), C.c_bool(False)
        buf = (C.c_ubyte * len(nals[k])).from_buffer_copy(nals[k])
        assert lib.libHMDec_push_nal_unit(dec, buf, len(nals[k]), k + 1 == len(nals), C.byref(new), C.byref(chk)) == OK
        if chk.value:
            pics = []
            while True:
                pic = lib.libHMDec_get_picture(dec)
                if not pic:
                    break
                pics.append(pic)
            first = {}
            for p in pics:                                 # ascending types, one picture after the other
                for t in range(24):
                    first[(p, t)] = _blocks(lib, dec, p, t)
            assert lib.libHMDEC_clear_internal_info(dec) == OK
            for t in order:                                # scrambled types, pictures interleaved (the index is rebuilt on every switch)
                for p in reversed(pics):
                    assert _blocks(lib, dec, p, t) == first[(p, t)], (lib.libHMDEC_get_POC(p), t)
            seen += len(pics)
            interleaved += len(pics) > 1
            if pics:
                assert len(first[(pics[0], 0)]) // 16 == 28       # one entry per CTU: 7 x 4 CTUs of 64x64 at 416x240
                assert len(first[(pics[0], 9)]) == 0              # QUIRK: PU_MERGE_INDEX reports nothing (libHMDecoder.cpp:663)
        if not new.value:
            k += 1
    assert lib.libHMDec_free_decoder(dec) == OK
    assert seen == 17 and interleaved >= 1
