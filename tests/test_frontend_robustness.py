"""CPU tests of the drop-in library's process-level behaviour (frontend/_build/libHMDecoder_b200.so, record-only back-end):
  * the shipped library exports exactly the reference wrapper's 16 entry points + the 4 libHMDecB200_* extensions and
    contains no CPU reconstruction (HM's prediction / transform / loop-filter sample code is not linked in);
  * decoders for streams of DIFFERENT SPS geometry (bit depth 8/10/12, CTU 64/32/16, 4:2:0/4:2:2/4:4:4) run concurrently as
    threads of one process and still emit byte-identical records (hm_threadsafe.cpp: geometry gate around HM's globals);
  * an unsupported bitstream feature is reported through the reference ABI: libHMDec_push_nal_unit returns LIBHMDEC_ERROR
    (and keeps returning it), nothing aborts."""
import ctypes as C
import os
import re
import subprocess
import threading
import numpy as np
import pytest
from conftest import GOLDEN, ROOT
from libhm_b200 import records

LIB = os.path.join(ROOT, "frontend", "_build", "libHMDecoder_b200.so")
REF_SYMS = ["libHMDec_get_version", "libHMDec_new_decoder", "libHMDec_free_decoder", "libHMDec_set_SEI_Check", "libHMDec_set_max_temporal_layer",
            "libHMDec_push_nal_unit", "libHMDec_get_picture", "libHMDEC_get_POC", "libHMDEC_get_picture_width", "libHMDEC_get_picture_height",
            "libHMDEC_get_picture_stride", "libHMDEC_get_image_plane", "libHMDEC_get_chroma_format", "libHMDEC_get_internal_bit_depth",
            "libHMDEC_get_internal_info", "libHMDEC_clear_internal_info"]
EXT_SYMS = ["libHMDecB200_new_decoder_ex", "libHMDecB200_hash_mismatch", "libHMDecB200_pack_picture", "libHMDecB200_unsupported"]
OK, ERROR = 0, 1

pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="frontend not built (needs the reference sources at build time)")


def _lib():
    os.environ["HMDEC_B200_QUIET"] = "1"
    l = C.CDLL(LIB)
    l.libHMDecB200_new_decoder_ex.restype = C.c_void_p
    l.libHMDecB200_new_decoder_ex.argtypes = [C.c_int, C.c_char_p]
    l.libHMDecB200_unsupported.restype = C.c_char_p
    l.libHMDecB200_unsupported.argtypes = [C.c_void_p]
    l.libHMDec_free_decoder.argtypes = [C.c_void_p]
    l.libHMDec_set_SEI_Check.argtypes = [C.c_void_p, C.c_bool]
    l.libHMDec_push_nal_unit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_bool, C.POINTER(C.c_bool), C.POINTER(C.c_bool)]
    l.libHMDec_get_picture.restype = C.c_void_p
    l.libHMDec_get_picture.argtypes = [C.c_void_p]
    l.libHMDEC_get_internal_bit_depth.argtypes = [C.c_int]
    return l


def _nals(path):
    data = open(path, "rb").read()
    starts = [m.start() for m in re.finditer(b"\x00\x00\x01", data)]
    out = []
    for i, s in enumerate(starts):
        e = starts[i + 1] if i + 1 < len(starts) else len(data)
        nal = data[s + 3:e]
        while i + 1 < len(starts) and nal.endswith(b"\x00"):
            nal = nal[:-1]
        out.append(nal)
    return out


def _decode(lib, dec, nals):
    """The push / re-push / drain loop of libHMDecoder.h:38-77; returns (pictures, first error code or OK)."""
    pics = 0
    newpic, check = C.c_bool(False), C.c_bool(False)
    for i, nal in enumerate(nals):
        buf = C.create_string_buffer(nal, len(nal))
        while True:
            rc = lib.libHMDec_push_nal_unit(dec, buf, len(nal), i == len(nals) - 1, C.byref(newpic), C.byref(check))
            if rc != OK:
                return pics, rc
            if check.value:
                while lib.libHMDec_get_picture(dec):
                    pics += 1
            if not newpic.value:
                break
    return pics, OK


def test_shipped_library_exports_only_the_wrapper_abi_and_holds_no_cpu_reconstruction():
    dyn = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    exported = sorted(l.split()[-1] for l in dyn.splitlines() if l.split()[1] in "TtDdBbVvWw" and not l.split()[-1].startswith("_"))
    assert exported == sorted(REF_SYMS + EXT_SYMS), exported
    allsyms = subprocess.check_output(["nm", "-C", LIB], text=True)
    for cpu_recon in ("TComPrediction::xPredInterBlk", "TComPrediction::motionCompensation", "TComLoopFilter::loopFilterPic", "TComLoopFilter::xEdgeFilterLuma",
                      "partialButterflyInverse", "TComTrQuant::invTransformNxN", "TComSampleAdaptiveOffset::offsetBlock", "TComSampleAdaptiveOffset::SAOProcess",
                      "TDecCu::xReconInter", "TDecCu::xIntraRecBlk", "TComInterpolationFilter::filterHor"):
        assert cpu_recon not in allsyms, f"the shipped drop-in still links {cpu_recon}"


MIXED = ["s_ra8_240p", "s_ra10_240p", "s_ctu16_240p", "s_ctu32_240p", "s_rext444_240p", "s_ra422_240p", "s_ld10_240p", "s_switch_240p",
         "s_ramintu8_240p", "s_mintu16_240p", "s_mintu32_240p"]       # the last three: HM partitions of 8 / 16 / 32 samples (other g_uiMaxCUDepth / g_uiAddCUDepth)


def test_decoders_of_different_geometry_run_concurrently_as_threads(tmp_path):
    lib = _lib()
    errors = []

    def run(name, rep):
        try:
            out = str(tmp_path / f"{name}.{rep}.hmr")
            dec = lib.libHMDecB200_new_decoder_ex(1, out.encode())
            assert dec
            lib.libHMDec_set_SEI_Check(dec, False)
            pics, rc = _decode(lib, dec, _nals(os.path.join(GOLDEN, name + ".bin")))
            depth = lib.libHMDEC_get_internal_bit_depth(0)          # the bit depth of the decoder THIS thread drove last
            lib.libHMDec_free_decoder(dec)
            assert rc == OK
            got = records.read_dump(out)
            ref = records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz"))
            assert pics == len(ref) == len(got), (name, pics, len(ref), len(got))
            assert depth == int(ref[-1].h["bit_depth_luma"]), (name, depth)
            for g, f in zip(got, ref):
                for field in records.Frame.FIELDS:
                    a, b = getattr(g, field), getattr(f, field)
                    if a is None or b is None:
                        assert (a is None or a.size == 0) and (b is None or b.size == 0), field
                    else:
                        assert np.array_equal(a, b), (name, int(f.h["poc"]), field)
        except BaseException as ex:          # noqa: BLE001 — report from the worker thread
            errors.append((name, rep, repr(ex)))

    threads = [threading.Thread(target=run, args=(n, r)) for r in range(2) for n in MIXED]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_pooled_buffers_pass_cleanly_from_one_bitstream_to_the_next(tmp_path):
    """One thread opens one bitstream after the other (different geometries, bit depths, chroma formats, intra-only and inter): the record
    vectors of a finished decoder's emitter and its parked picture buffers are adopted by the next decoder (frontend/hm_emit.cpp:
    RecordStorage, frontend/hm_fast.cpp: g_picPool) — every record dump must still equal the golden one, twice over."""
    lib = _lib()
    order = ["s_ra8_240p", "c1_intra8_240p", "s_ra8_odd", "s_ra10_240p", "s_rext444_240p", "s_ra8_240p", "s_ra422_240p", "s_gray400_240p", "s_cra_240p", "s_seek_240p", "c1_intra8_240p",
             "s_mintu8_240p", "s_ra8_240p", "s_mintu32_240p", "s_ramintu8_240p", "s_mintu16_240p"]
    for rep, name in enumerate(order + order):
        out = str(tmp_path / f"{rep}.hmr")
        dec = lib.libHMDecB200_new_decoder_ex(1, out.encode())
        assert dec
        lib.libHMDec_set_SEI_Check(dec, False)
        pics, rc = _decode(lib, dec, _nals(os.path.join(GOLDEN, name + ".bin")))
        lib.libHMDec_free_decoder(dec)
        assert rc == OK
        got = records.read_dump(out)
        ref = records.read_dump(os.path.join(GOLDEN, name + ".hmr.gz"))
        assert pics == len(ref) == len(got), (name, pics, len(ref), len(got))
        for g, f in zip(got, ref):
            for field in records.Frame.FIELDS:
                a, b = getattr(g, field), getattr(f, field)
                if a is None or b is None:
                    assert (a is None or a.size == 0) and (b is None or b.size == 0), (name, field)
                else:
                    assert np.array_equal(a, b), (name, rep, int(f.h["poc"]), field)
        os.remove(out)


def test_unsupported_feature_is_reported_through_the_reference_abi(tmp_path):
    """A picture the emitter refuses (every tool an HM-built stream can carry is on the GPU path by now, so the refusal is forced after the
    first picture with HMDEC_B200_REFUSE_AFTER — the same HmEmitter::fail() an unsupported SPS tool takes): the decoder must stop with
    LIBHMDEC_ERROR from libHMDec_push_nal_unit (sticky), name the reason through libHMDecB200_unsupported, and leave the process alive."""
    lib = _lib()
    dec = lib.libHMDecB200_new_decoder_ex(1, str(tmp_path / "g.hmr").encode())
    assert dec
    nals = _nals(os.path.join(GOLDEN, "s_ra8_240p.bin"))
    os.environ["HMDEC_B200_REFUSE_AFTER"] = "1"
    try:
        pics, rc = _decode(lib, dec, nals)
    finally:
        del os.environ["HMDEC_B200_REFUSE_AFTER"]
    assert rc == ERROR
    assert b"HMDEC_B200_REFUSE_AFTER" in lib.libHMDecB200_unsupported(dec)
    newpic, check = C.c_bool(False), C.c_bool(False)
    buf = C.create_string_buffer(nals[-1], len(nals[-1]))
    assert lib.libHMDec_push_nal_unit(dec, buf, len(nals[-1]), True, C.byref(newpic), C.byref(check)) == ERROR
    lib.libHMDec_free_decoder(dec)
