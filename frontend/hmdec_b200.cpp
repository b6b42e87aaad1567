// hmdec_b200.cpp — drop-in implementation of the libHMDecoder C wrapper
// (reference: source/App/libHMDecoder/libHMDecoder.{h,cpp}) on top of the B200 reconstruction engine.
//
// Parsing (NAL, parameter sets, CABAC, motion derivation, DPB management, output bumping) is HM's
// TDecTop, unmodified, exactly as in the reference wrapper.  Reconstruction is NOT: every parsed CTU
// is turned into flat records (hm_emit.cpp) and every finished picture is handed to an HmFrameSink —
// by default the GPU engine.  Sample planes come back from the device only when somebody looks at
// them (libHMDEC_get_image_plane, SEI hash check).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <list>
#include <map>
#include <mutex>
#include <iostream>
#include <sstream>
#include <fstream>
#include <algorithm>
#include <limits>
#include <iomanip>
#include <cmath>
#include <cassert>
#include <malloc.h>
#define private public
#include "TLibCommon/TComSlice.h"
#include "TLibCommon/CommonDef.h"
#include "TLibCommon/TComTU.h"
#include "TLibDecoder/TDecTop.h"
#undef private
#include "TLibDecoder/NALread.h"
#include "libHMDecoder_api.h"
#include "hm_emit.h"
#include "hm_fast.h"
#include "hm_threadsafe.h"
#include "hm_waitstats.h"

// HM keeps the "hash mismatch seen" flag in a global that the application must define (TDecGop.cpp:48); thread_local in this
// build (frontend/Makefile patches the declaration), saved and restored per decoder around every HM call.
thread_local bool g_md5_mismatch = false;

HmFrameSink* hm_new_dump_sink(const char* path);
HmFrameSink* hm_new_null_sink();
HmFrameSink* hm_new_gpu_sink();           // gpu_sink.cpp
// internals.cpp: flat block index of one picture, built by the first query and kept until the next push
struct HmInternalsCache;
HmInternalsCache* hm_internals_cache_new();
void hm_internals_cache_free(HmInternalsCache* c);
void hm_internals_cache_invalidate(HmInternalsCache* c);
std::vector<libHMDec_BlockValue>* hm_collect_internals(HmInternalsCache* cache, std::vector<libHMDec_BlockValue>& out, TComPic* pic, libHMDec_info_type type);

namespace {

// picture -> owning decoder, for the ctx-less plane accessor (libHMDEC_get_image_plane has no ctx argument)
std::mutex g_picOwnerLock;
std::map<const void*, struct Decoder*> g_picOwner;

struct Decoder
{
  TDecTop      top;
  HmFrameSink* sink;
  HmEmitter*   emitter;

  // output state, cf. libHMDecoder.cpp:19-61 / TAppDecTop::xWriteOutput
  int   maxTemporalLayer;
  int   lastDisplayedPoc;
  int   skipFrames;
  TComList<TComPic*>* dpb;
  int   cursor;              // index into *dpb of the next candidate for libHMDec_get_picture
  int   pendingOutput;       // pictures marked for output and not yet displayed
  int   dpbFullness;
  unsigned reorderLimit, bufferingLimit;
  bool  loopFilterDone;
  bool  flushing;            // output everything that is marked, regardless of the bumping rule
  bool  flushAfterThisPass;  // eof: one normal pass, then a flush pass
  bool  hashMismatch;
  bool  failed;              // sticky: an unsupported feature or an engine error was reported through LIBHMDEC_ERROR
  Int   prevTid0POC;         // per-decoder copy of TComSlice::m_prevTid0POC (thread_local in this build)
  HmGeomKey geom;            // SPS-dependent HM globals this decoder runs under (hm_threadsafe.cpp); invalid before the first activation
  bool  statsActive;         // HMDEC_B200_STATS counts this decoder (hm_waitstats.h)
  std::vector<libHMDec_BlockValue> internals;
  HmInternalsCache* internalsIndex;

  Decoder(HmFrameSink* s)
    : sink(s), emitter(new HmEmitter(s)), maxTemporalLayer(-1), lastDisplayedPoc(-MAX_INT), skipFrames(0), dpb(NULL),
      cursor(0), pendingOutput(0), dpbFullness(0), reorderLimit(0), bufferingLimit(0), loopFilterDone(false),
      flushing(false), flushAfterThisPass(false), hashMismatch(false), failed(false), prevTid0POC(0), internalsIndex(hm_internals_cache_new())
  {
    top.create();
    top.init();
    top.setDecodedPictureHashSEIEnabled(true);
  }
  ~Decoder()
  {
    {
      std::lock_guard<std::mutex> g(g_picOwnerLock);
      for (std::map<const void*, Decoder*>::iterator it = g_picOwner.begin(); it != g_picOwner.end();)
        if (it->second == this) g_picOwner.erase(it++); else ++it;
    }
    if (geom.valid) hm_geom_enter(geom);        // HM's teardown walks structures sized by the CTU geometry
    hm_emit_set_current(emitter);
    sink->drainHashes(true);
    sink->releaseHostBuffers();          // nothing may still be DMA-ing into HM's planes
    hm_fast_release_decoder(&top);       // pooled planes go back; buffers dropped by a flush are freed
    top.destroy();
    hm_emit_set_current(NULL);
    if (geom.valid) hm_geom_leave(geom);
    delete emitter;
    delete sink;
    hm_internals_cache_free(internalsIndex);
  }

  void claimPictures()
  {
    if (!dpb) return;
    std::lock_guard<std::mutex> g(g_picOwnerLock);
    for (TComList<TComPic*>::iterator it = dpb->begin(); it != dpb->end(); ++it) g_picOwner[*it] = this;
  }

  // C.5.2.2 bumping inputs, recomputed whenever a picture completes (TAppDecTop.cpp:324-380)
  void refreshOutputCounters()
  {
    TComSPS* sps = top.getActiveSPS();
    const unsigned layers = sps->getMaxTLayers();
    const unsigned t = (maxTemporalLayer == -1 || maxTemporalLayer >= (int)layers) ? layers - 1 : (unsigned)maxTemporalLayer;
    reorderLimit   = sps->getNumReorderPics(t);
    bufferingLimit = sps->getMaxDecPicBuffering(t);
    pendingOutput = dpbFullness = 0;
    for (TComList<TComPic*>::iterator it = dpb->begin(); it != dpb->end(); ++it)
    {
      TComPic* p = *it;
      if (p->getOutputMark() && p->getPOC() > lastDisplayedPoc) { pendingOutput++; dpbFullness++; }
      else if (p->getSlice(0)->isReferenced()) dpbFullness++;
    }
  }
};

inline Decoder* D(libHMDec_context* c) { return (Decoder*)c; }

// The geometry key of the last decoder this thread drove into HM: what the ctx-less libHMDEC_get_internal_bit_depth answers.
thread_local HmGeomKey t_lastGeom;

// The key a slice NAL is going to run under: first_slice_segment_in_pic_flag u(1), no_output_of_prior_pics_flag u(1) for IRAP
// types, slice_pic_parameter_set_id ue(v) (7.3.6.1; at most 15 bits: no emulation-prevention byte can occur that early because
// the second NAL header byte is never 0) -> PPS -> SPS, looked up where TDecTop keeps what it has parsed so far.
HmGeomKey peekSliceGeometry(Decoder* d, const InputNALUnit& nalu, const std::vector<uint8_t>& bytes)
{
  if (bytes.size() < 4) return d->geom;
  const unsigned w = ((unsigned)bytes[2] << 8) | bytes[3];
  int pos = 15 - 1;                                                     // bit 15 = first_slice_segment_in_pic_flag
  if (nalu.m_nalUnitType >= NAL_UNIT_CODED_SLICE_BLA_W_LP && nalu.m_nalUnitType <= NAL_UNIT_RESERVED_IRAP_VCL23) pos--;
  int zeros = 0;
  while (pos >= 0 && !((w >> pos) & 1)) { zeros++; pos--; }
  if (pos < zeros) return d->geom;                                      // longer than 16 bits: not a legal PPS id
  const unsigned id = ((w >> (pos - zeros)) & ((1u << (zeros + 1)) - 1)) - 1;
  TComPPS* pps = id < 64 ? d->top.m_parameterSetManagerDecoder.getPrefetchedPPS((Int)id) : NULL;
  TComSPS* sps = pps ? d->top.m_parameterSetManagerDecoder.getPrefetchedSPS(pps->getSPSId()) : NULL;
  return sps ? hm_geom_key_of(sps) : d->geom;
}

struct GeomScope
{
  Decoder* d; bool entered;
  GeomScope(Decoder* dec, const HmGeomKey& k, bool needed) : d(dec), entered(needed) { if (entered) hm_geom_enter(k); }
  ~GeomScope()
  {
    if (!entered) return;
    TComSPS* sps = d->top.getActiveSPS();
    if (sps) d->geom = hm_geom_key_of(sps);
    t_lastGeom = d->geom;
    hm_geom_leave(d->geom);
  }
};
inline bool isIrapFlushType(NalUnitType t)
{
  return t == NAL_UNIT_CODED_SLICE_IDR_W_RADL || t == NAL_UNIT_CODED_SLICE_IDR_N_LP || t == NAL_UNIT_CODED_SLICE_BLA_N_LP ||
         t == NAL_UNIT_CODED_SLICE_BLA_W_RADL || t == NAL_UNIT_CODED_SLICE_BLA_W_LP;
}
inline ComponentID toComp(libHMDec_ColorComponent c, bool& ok)
{
  ok = (c == LIBHMDEC_LUMA || c == LIBHMDEC_CHROMA_U || c == LIBHMDEC_CHROMA_V);
  return c == LIBHMDEC_LUMA ? COMPONENT_Y : (c == LIBHMDEC_CHROMA_U ? COMPONENT_Cb : COMPONENT_Cr);
}

} // namespace

extern "C" {

const char* libHMDec_get_version(void) { return NV_VERSION; }

// HM allocates and frees every picture buffer and ~50 KB of per-CTU arrays for EVERY picture (TDecTop::xGetNewPicBuffer:
// rpcPic->destroy(); rpcPic->create(), TDecTop.cpp:187-189).  With glibc's defaults those blocks are mmap'ed, so each
// picture costs ~100 MB of page faults and, with several decoder threads in one process, serialises on the mm lock.
// Keeping freed blocks in the malloc arenas removes both (set HMDEC_B200_KEEP_MALLOC=1 to leave malloc untouched).
static void tuneMallocOnce()
{
  static std::once_flag once;
  std::call_once(once, []() {
    if (getenv("HMDEC_B200_KEEP_MALLOC")) return;
    mallopt(M_MMAP_THRESHOLD, 32 << 20);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
  });
}

extern "C" int hm_cpu_is_x86_64_v3(void);      // hm_cpu_guard.cpp

libHMDec_context* libHMDecB200_new_decoder_ex(int backend, const char* arg)
{
#ifdef __AVX2__
  if (!hm_cpu_is_x86_64_v3())
  {
    fprintf(stderr, "libHMDecoder_b200: this build needs an x86-64-v3 CPU (AVX2, BMI2); rebuild frontend/ with ARCHFLAGS=\n");
    return NULL;
  }
#endif
  tuneMallocOnce();
  const bool counted = hm_wait_stats().on && ++hm_wait_stats().decoders > hm_wait_stats().skip;
  t_hmwActive = counted;
  HmWaitScope ws(HMW_NEW_DECODER);
  HmFrameSink* sink = NULL;
  if (backend == 0) sink = hm_new_gpu_sink();
  else if (backend == 1 && arg && !strcmp(arg, "null")) sink = hm_new_null_sink();
  else if (backend == 1 && arg) sink = hm_new_dump_sink(arg);
  if (!sink) return NULL;
  Decoder* d = new Decoder(sink);
  d->statsActive = counted;
  return (libHMDec_context*)d;
}

libHMDec_context* libHMDec_new_decoder(void)
{
  // tools may redirect the default back-end; there is deliberately no CPU reconstruction back-end
  const char* dump = getenv("HMDEC_B200_DUMP");
  return dump ? libHMDecB200_new_decoder_ex(1, dump) : libHMDecB200_new_decoder_ex(0, NULL);
}

libHMDec_error libHMDec_free_decoder(libHMDec_context* decCtx)
{
  if (!decCtx) return LIBHMDEC_ERROR;
  t_hmwActive = D(decCtx)->statsActive;
  HmWaitScope ws(HMW_FREE_DECODER);
  delete D(decCtx);
  return LIBHMDEC_OK;
}

void libHMDec_set_SEI_Check(libHMDec_context* decCtx, bool check_hash)
{
  if (decCtx) D(decCtx)->top.setDecodedPictureHashSEIEnabled(check_hash);
}

void libHMDec_set_max_temporal_layer(libHMDec_context* decCtx, int max_layer)
{
  if (decCtx) D(decCtx)->maxTemporalLayer = max_layer;
}

bool libHMDecB200_hash_mismatch(libHMDec_context* decCtx)
{
  if (!decCtx) return false;
  Decoder* d = D(decCtx);
  t_hmwActive = d->statsActive;
  { HmWaitScope ws(HMW_HASH_WAIT); d->sink->drainHashes(true); }            // digests still in flight on the device
  if (d->sink->hashMismatchSeen()) d->hashMismatch = true;
  return d->hashMismatch;
}
long libHMDecB200_pack_picture(libHMDec_context* decCtx, libHMDec_picture* pic, int outBitDepthLuma, int outBitDepthChroma, void* dst, size_t capacity)
{
  Decoder* d = D(decCtx);
  if (!d || !pic) return -1;
  TComPic* p = (TComPic*)pic;
  // the window TAppDecoder crops to when writing `-o` (TAppDecTop.cpp:476-487): conformance window (+ default display window: not respected by default)
  const Window& conf = p->getConformanceWindow();
  const int crop[4] = { conf.getWindowLeftOffset(), conf.getWindowRightOffset(), conf.getWindowTopOffset(), conf.getWindowBottomOffset() };
  const int bd[2] = { outBitDepthLuma, outBitDepthChroma };
  size_t bytes = 0;
  if (!d->sink->readPacked(p, bd, crop, dst, capacity, &bytes)) return -1;
  return (long)bytes;
}

const char* libHMDecB200_unsupported(libHMDec_context* decCtx) { return decCtx ? D(decCtx)->emitter->unsupported() : NULL; }

libHMDec_error libHMDec_push_nal_unit(libHMDec_context* decCtx, const void* data8, int length, bool eof, bool& bNewPicture, bool& checkOutputPictures)
{
  Decoder* d = D(decCtx);
  if (!d) return LIBHMDEC_ERROR;
  t_hmwActive = d->statsActive;
  HmWaitScope wsTotal(HMW_PUSH_TOTAL);
  if (length <= 0) return LIBHMDEC_ERROR_READ_ERROR;
  if (length < 4 && !eof) return LIBHMDEC_ERROR_READ_ERROR;
  hm_internals_cache_invalidate(d->internalsIndex);      // picture buffers only change inside a push

  // tolerate a leading Annex-B start code (00 00 01 / 00 00 00 01); the payload starts at the 2-byte NAL header.
  // (The reference tests bytes 0,1,1 for the 3-byte form, libHMDecoder.cpp:128, which never matches a real start code.)
  const uint8_t* p = (const uint8_t*)data8;
  int skip = 0;
  if (length >= 3 && p[0] == 0 && p[1] == 0 && p[2] == 1) skip = 3;
  else if (length >= 4 && p[0] == 0 && p[1] == 0 && p[2] == 0 && p[3] == 1) skip = 4;
  std::vector<uint8_t> bytes(p + skip, p + length);
  if (bytes.size() < 2) return LIBHMDEC_ERROR_READ_ERROR;

  // an unsupported stream feature or an engine error is reported through the reference ABI and is final for this decoder
  if (d->failed || d->emitter->unsupported() || d->sink->error()) { d->failed = true; return LIBHMDEC_ERROR; }

  InputNALUnit nalu;
  read(nalu, bytes);                                   // NALread.cpp:144-154

  const bool vcl = nalu.m_nalUnitType <= NAL_UNIT_RESERVED_VCL31;
  const bool completes = eof || nalu.m_nalUnitType == NAL_UNIT_EOS;                  // may run the loop-filter hook without parsing a slice
  // Parameter sets are activated by the first slice of a picture only (TDecTop.cpp:505); a slice that arrives while a picture is
  // open either belongs to it or merely ends it (bNewPicture, nothing activated, the finished picture is filtered under ITS key).
  const bool activates = vcl && d->top.m_bFirstSliceInPicture;
  GeomScope gate(d, activates ? peekSliceGeometry(d, nalu, bytes) : d->geom, vcl || (completes && d->geom.valid));
  // parseSPS writes the global g_bitDepthInStream and reads it back a few lines later (TDecCAVLC.cpp:617-643; nothing else on the product
  // path reads it): SPS NALs of different decoders exclude EACH OTHER — not the slice parsing of the other decoders, which a pass through
  // the geometry gate with an unknown key would stall for the length of a picture at the start of every bitstream.
  static std::mutex spsParseLock;
  std::unique_lock<std::mutex> spsGuard(spsParseLock, std::defer_lock);
  if (nalu.m_nalUnitType == NAL_UNIT_SPS) spsGuard.lock();

  hm_emit_set_current(d->emitter);
  hm_fast_set_skip_coeff_fill(d->emitter->cleanCoeffs());
  bNewPicture = false;
  if (!(d->maxTemporalLayer >= 0 && (int)nalu.m_temporalId > d->maxTemporalLayer))
  {
    g_md5_mismatch = d->hashMismatch;
    TComSlice::m_prevTid0POC = d->prevTid0POC;
    bNewPicture = d->top.decode(nalu, d->skipFrames, d->lastDisplayedPoc);   // TDecTop.cpp:729
    d->prevTid0POC = TComSlice::m_prevTid0POC;
    d->hashMismatch = g_md5_mismatch;
  }

  // a picture is complete when the first slice of the next one shows up, at EOS, or at end of stream
  if (bNewPicture || eof || nalu.m_nalUnitType == NAL_UNIT_EOS)
  {
    if (!d->loopFilterDone || !eof)
    {
      int poc;
      g_md5_mismatch = d->hashMismatch;
      { HmWaitScope ws(HMW_PICTURE_DONE); d->top.executeLoopFilters(poc, d->dpb); }          // -> TDecGop::filterPicture hook -> engine
      d->hashMismatch = g_md5_mismatch;
      d->claimPictures();
    }
    d->loopFilterDone = (nalu.m_nalUnitType == NAL_UNIT_EOS);
  }
  hm_emit_set_current(NULL);
  hm_fast_set_skip_coeff_fill(false);

  checkOutputPictures = false;
  d->flushing = false;
  if (bNewPicture && isIrapFlushType(nalu.m_nalUnitType)) { checkOutputPictures = true; d->flushing = true; }
  if (nalu.m_nalUnitType == NAL_UNIT_EOS) checkOutputPictures = true;

  const bool vclConsumed = !bNewPicture && nalu.m_nalUnitType >= NAL_UNIT_CODED_SLICE_TRAIL_N && nalu.m_nalUnitType <= NAL_UNIT_RESERVED_VCL31;
  if ((bNewPicture || vclConsumed) && d->dpb != NULL)
  {
    checkOutputPictures = true;
    d->refreshOutputCounters();
  }
  if (eof) { checkOutputPictures = true; d->flushAfterThisPass = true; }
  if (checkOutputPictures) d->cursor = 0;
  if (d->emitter->unsupported() || d->sink->error()) { d->failed = true; return LIBHMDEC_ERROR; }
  return LIBHMDEC_OK;
}

libHMDec_picture* libHMDec_get_picture(libHMDec_context* decCtx)
{
  Decoder* d = D(decCtx);
  if (!d || !d->dpb || d->dpb->size() == 0) return NULL;
  if (d->cursor < 0 || d->cursor > (int)d->dpb->size()) return NULL;

  TComList<TComPic*>::iterator it = d->dpb->begin();
  for (int i = 0; i < d->cursor; i++) ++it;
  if (it != d->dpb->end() && (*it)->isField()) return NULL;      // field output unsupported, as in the reference

  for (; it != d->dpb->end(); ++it, d->cursor++)
  {
    TComPic* pic = *it;
    const bool bump = pic->getOutputMark() && pic->getPOC() > d->lastDisplayedPoc &&
                      (d->pendingOutput > (int)d->reorderLimit || d->dpbFullness > (int)d->bufferingLimit);
    if (!((d->flushing && pic->getOutputMark()) || bump)) continue;

    if (!d->flushing) d->pendingOutput--;
    if (!pic->getSlice(0)->isReferenced()) d->dpbFullness--;
    d->lastDisplayedPoc = pic->getPOC();
    if (!pic->getSlice(0)->isReferenced() && pic->getReconMark())
    {
      pic->setReconMark(false);
      pic->getPicYuvRec()->setBorderExtension(false);
    }
    pic->setOutputMark(false);
    return (libHMDec_picture*)pic;
  }

  if (d->flushing)
  {
    // NOTE: the planes of already returned pictures stay valid: TComList::clear() drops the pointers only
    d->dpb->clear();
    d->lastDisplayedPoc = -MAX_INT;
    d->flushing = false;
  }
  if (d->flushAfterThisPass)
  {
    d->flushAfterThisPass = false;
    d->flushing = true;
    d->cursor = 0;
    return libHMDec_get_picture(decCtx);
  }
  return NULL;
}

int libHMDEC_get_POC(libHMDec_picture* pic) { return pic ? ((TComPic*)pic)->getPOC() : -1; }

int libHMDEC_get_picture_width(libHMDec_picture* pic, libHMDec_ColorComponent c)
{
  bool ok; ComponentID id = toComp(c, ok);
  return (pic && ok) ? ((TComPic*)pic)->getPicYuvRec()->getWidth(id) : -1;
}
int libHMDEC_get_picture_height(libHMDec_picture* pic, libHMDec_ColorComponent c)
{
  bool ok; ComponentID id = toComp(c, ok);
  return (pic && ok) ? ((TComPic*)pic)->getPicYuvRec()->getHeight(id) : -1;
}
int libHMDEC_get_picture_stride(libHMDec_picture* pic, libHMDec_ColorComponent c)
{
  bool ok; ComponentID id = toComp(c, ok);
  return (pic && ok) ? ((TComPic*)pic)->getPicYuvRec()->getStride(id) : -1;
}

short* libHMDEC_get_image_plane(libHMDec_picture* pic, libHMDec_ColorComponent c)
{
  bool ok; ComponentID id = toComp(c, ok);
  if (!pic || !ok) return NULL;
  Decoder* owner = NULL;
  {
    std::lock_guard<std::mutex> g(g_picOwnerLock);
    std::map<const void*, Decoder*>::iterator it = g_picOwner.find(pic);
    if (it != g_picOwner.end()) owner = it->second;
  }
  if (owner)
  {
    t_hmwActive = owner->statsActive;
    { HmWaitScope ws(HMW_PLANE_WAIT); owner->sink->fetchPicture((TComPic*)pic); }            // device -> HM's padded host plane, once per picture
    if (owner->sink->error()) return NULL;               // the samples never arrived
  }
  return ((TComPic*)pic)->getPicYuvRec()->getAddr(id);
}

libHMDec_ChromaFormat libHMDEC_get_chroma_format(libHMDec_picture* pic)
{
  if (!pic) return LIBHMDEC_CHROMA_UNKNOWN;
  switch (((TComPic*)pic)->getChromaFormat())
  {
    case CHROMA_400: return LIBHMDEC_CHROMA_400;
    case CHROMA_420: return LIBHMDEC_CHROMA_420;
    case CHROMA_422: return LIBHMDEC_CHROMA_422;
    case CHROMA_444: return LIBHMDEC_CHROMA_444;
    default:         return LIBHMDEC_CHROMA_UNKNOWN;
  }
}

// The reference reads HM's global (libHMDecoder.cpp: no ctx argument).  With several decoders in the process the global belongs to
// whoever ran last, so the answer is the bit depth of the decoder THIS thread drove last; before that, the global like the reference.
int libHMDEC_get_internal_bit_depth(libHMDec_ColorComponent c)
{
  if (c != LIBHMDEC_LUMA && c != LIBHMDEC_CHROMA_U && c != LIBHMDEC_CHROMA_V) return -1;
  const int ch = c == LIBHMDEC_LUMA ? CHANNEL_TYPE_LUMA : CHANNEL_TYPE_CHROMA;
  return t_lastGeom.valid ? t_lastGeom.bitDepth[ch] : g_bitDepth[ch];
}

std::vector<libHMDec_BlockValue>* libHMDEC_get_internal_info(libHMDec_context* decCtx, libHMDec_picture* pic, libHMDec_info_type type)
{
  Decoder* d = D(decCtx);
  if (!d) return NULL;
  d->internals.clear();
  if (!pic) return NULL;
  GeomScope gate(d, d->geom, d->geom.valid);           // the walk indexes the z-scan tables of this decoder's CTU geometry
  return hm_collect_internals(d->internalsIndex, d->internals, (TComPic*)pic, type);
}

libHMDec_error libHMDEC_clear_internal_info(libHMDec_context* decCtx)
{
  if (!decCtx) return LIBHMDEC_ERROR;
  D(decCtx)->internals.clear();
  return LIBHMDEC_OK;
}

} // extern "C"
