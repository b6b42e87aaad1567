// hm_emit.h — reference-side binding: turns HM's parsed per-CTU data (TComDataCU) into the flat
// per-frame records of include/hmr_records.h.  This file and hm_emit.cpp are compiled against the
// HM headers under /root/reference (they are the glue a libHM maintainer would add, see
// INTEGRATION.md); the reconstruction engine itself (libhm_b200/csrc) never sees an HM type.
#ifndef HM_EMIT_H
#define HM_EMIT_H

#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <vector>
#include <memory>
#include <new>
#include <map>
#include "hmr_records.h"

class TComPic;
class TComDataCU;
class TComSlice;
class TComLoopFilter;
class TComSampleAdaptiveOffset;

// Receives one finished picture's records.  Implementations: file dumper (tools), GPU engine
// (hmdec_b200.cpp).  `pic` is HM's DPB entry the picture belongs to (slot bookkeeping / D2H target).
struct HmFrameSink
{
  virtual ~HmFrameSink() {}
  // Called once all CTUs of the picture are parsed and the loop-filter side info is known.
  // false = the sink could not take the picture (error() says why); the decoder then stops and reports LIBHMDEC_ERROR.
  virtual bool frameReady(const hmr_frame_desc& desc, TComPic* pic) = 0;
  // NULL, or why the sink stopped working (device error, rejected picture).  Sticky.
  virtual const char* error() const { return NULL; }
  // Make HM's own TComPicYuv of `pic` hold the final reconstruction (D2H for the GPU sink; no-op when
  // HM reconstructed on the CPU).  Needed before the SEI hash check / plane access.
  virtual void fetchPicture(TComPic* pic) = 0;
  // true: HM's CPU reconstruction runs as well (verification / golden generation); false: GPU only.
  virtual bool wantHmRecon() const = 0;
  // wantHmRecon() only: HM's CPU planes after stage 0 = CU reconstruction, 1 = deblocking, 2 = SAO (final).
  virtual void hmStage(int stage, TComPic* pic) { (void)stage; (void)pic; }
  // SEI hash methods that parallelise (2 = CRC, 3 = checksum) computed where the picture lives; false = not available.
  virtual bool deviceHash(TComPic* pic, int method, uint32_t out[3]) { (void)pic; (void)method; (void)out; return false; }
  // SEI MD5 checked asynchronously where the picture lives: `expected` = 16 bytes per component, `line` = the status line
  // up to the hash field.  The sink prints the completed line (unless quiet) and records a mismatch when the digest
  // arrives; lines keep their order.  false = not available (the caller hashes on the host).
  virtual bool asyncMd5(TComPic* pic, const unsigned char* expected, int ncomp, const std::string& line, bool quiet) { (void)pic; (void)expected; (void)ncomp; (void)line; (void)quiet; return false; }
  // A finished status line that must not overtake lines still waiting for their digest.
  virtual void orderedPrint(const std::string& line) { fputs(line.c_str(), stdout); }
  // Deliver digests that have arrived (wait = true: all of them).
  virtual void drainHashes(bool wait) { (void)wait; }
  virtual bool hashMismatchSeen() const { return false; }
  // The picture in TAppDecoder's `-o` wire format, packed where it lives (crop = left, right, top, bottom luma samples;
  // outBitDepth = luma, chroma; 0 = internal).  *bytes = packed size; dst == NULL only queries it.  false = not available.
  virtual bool readPacked(TComPic* pic, const int outBitDepth[2], const int crop[4], void* dst, size_t capacity, size_t* bytes)
  { (void)pic; (void)outBitDepth; (void)crop; (void)dst; (void)capacity; (void)bytes; return false; }
  // Called before HM's picture buffers are freed (the sink may hold page-locks on them).
  virtual void releaseHostBuffers() {}
};

#include "hm_fast.h"

// The level arena of a picture: int16 levels of every coded TU back to back.  Grown a whole TU at a time and written at once, so
// neither a value-initialising fill nor per-element construction is wanted (std::vector::resize did the first, and with a
// default-initialising allocator still walked every element: 4 % of an intra picture's parse) — a pointer bump, geometric growth.
struct HmLevelArena
{
  int16_t* p; size_t n, cap;
  HmLevelArena() : p(NULL), n(0), cap(0) {}
  ~HmLevelArena() { free(p); }
  void clear() { n = 0; }
  size_t size() const { return n; }
  const int16_t* data() const { return p; }
  void swap(HmLevelArena& o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(cap, o.cap); }
  // `count` more entries (uninitialised); returns the first of them.  Earlier pointers into the arena are invalid afterwards.
  inline int16_t* grow(size_t count)
  {
    if (n + count > cap) reserve(n + count);
    int16_t* r = p + n;
    n += count;
    return r;
  }
  void reserve(size_t want)
  {
    size_t c = cap ? cap : 4096;
    while (c < want) c *= 2;
    int16_t* q = (int16_t*)realloc(p, c * sizeof(int16_t));
    if (!q) throw std::bad_alloc();
    p = q; cap = c;
  }
private:
  HmLevelArena(const HmLevelArena&);
  HmLevelArena& operator=(const HmLevelArena&);
};

class HmEmitter
{
public:
  explicit HmEmitter(HmFrameSink* sink);
  ~HmEmitter();

  void onCtuParsed(TComDataCU* ctu);                                      // splice at TDecSlice.cpp:334
  void onPictureParsed(TComPic* pic, TComLoopFilter* lf, TComSampleAdaptiveOffset* sao, bool lfCrossTiles); // splice at TDecGop.cpp:157
  HmFrameSink* sink() { return m_sink; }
  int  slotOf(TComPic* pic);
  void onLostPicture(TComPic* fill, TComPic* src, int poc);               // TDecTop::xCreateLostPicture: `fill` stands in for a missing reference as a copy of `src`
  void releaseSlot(TComPic* pic);                                         // the picture buffer is gone (hm_fast.cpp): its DPB slot is free again
  const char* unsupported() const { return m_unsupported; }
  bool cleanCoeffs() const { return m_cleanCoeffs; }
  HmPrefetchCursor m_prefetch;                                            // CTU metadata prefetch (hm_fast.h)

private:
  struct CuCtx;
  void beginFrame(TComPic* pic, TComDataCU* ctu);
  void walkCU(TComDataCU* ctu, unsigned absPartIdx, unsigned depth);
  void emitInterCU(TComDataCU* ctu, unsigned absPartIdx, unsigned depth, int cuX, int cuY, int cuSize);
  void emitIntraCU(TComDataCU* ctu, unsigned absPartIdx, unsigned depth, int cuX, int cuY);
  void emitPcmCU(TComDataCU* ctu, unsigned absPartIdx, unsigned depth, int cuX, int cuY);
  void interResidual(CuCtx& c, int compID, void* rTu);
  void intraQT(CuCtx& c, int chType, void* rTu);
  void intraBlk(CuCtx& c, int compID, void* rTu);
  uint32_t emitResidualTU(CuCtx& c, int compID, void* rTu, bool intra, bool coded, int alpha);
  void deblockCtu(TComDataCU* ctu);
  void bsWalk(TComDataCU* ctu, unsigned absPartIdx, unsigned depth, TComLoopFilter* lf);
  void bsDirect(TComDataCU* ctu, unsigned absPartIdx, unsigned depth, bool lfCrossTiles);   // edges enumerated from the CU / PU / TU geometry, no flag arrays
  void saoInfo(TComPic* pic, TComSampleAdaptiveOffset* sao);
  void fail(const char* what);
  void swapStorage(void* recordStorage);   // hm_emit.cpp: record vectors are parked process-wide between decoders

  HmFrameSink* m_sink;
  TComLoopFilter* m_lf;              // own instance: HM's edge-flag machinery, run per CTU while the CTU is still in cache
  unsigned     m_lfDepth;
  bool         m_anyDeblock;
  TComPic*     m_curPic;
  bool         m_open;
  const char*  m_unsupported;
  std::map<TComPic*, int> m_slots;
  std::vector<int> m_freeSlots;      // slot numbers whose picture buffer was destroyed
  std::string  m_failText;

  hmr_frame_hdr                    m_hdr;
  std::vector<hmr_tu>              m_tu, m_tuSorted;
  HmLevelArena                     m_coef;
  std::vector<hmr_intra>           m_intra;
  std::vector<hmr_intra>           m_intraTmp[3];
  std::vector<hmr_ctu_intra_range> m_range;
  std::vector<hmr_pu>              m_pu;
  std::vector<uint32_t>            m_puPrefix;
  std::vector<hmr_ctu>             m_ctu;
  std::vector<uint8_t>             m_bs;
  std::vector<int8_t>              m_qp;
  std::vector<uint8_t>             m_cuFlags;
  std::vector<uint8_t>             m_scaling;
  std::vector<hmr_wp>              m_wp;
  std::vector<uint8_t>             m_puRefIdx;
  int m_bsStride, m_qpStride;
  int m_lgU;                      // log2 of (HM's partition unit / 4 samples): 0 = 4x4 partitions, 1 = 8x8 (min TU 8)
  bool m_in422SubTu;                       // inside the two square halves of a 4:2:2 chroma TU
  bool m_cleanCoeffs;                      // hm_fast.cpp: HM's whole-CTU coefficient zero fills are skipped for this decoder
  double m_tCtu, m_tBs, m_tPic, m_tSink;   // HMDEC_B200_STATS: host time spent emitting records
  int m_nPic;
};

// The emitter the hooks (TDecCu::decompressCU / TDecGop::filterPicture replacements) talk to.
// Set by the wrapper around every TDecTop call; one decoder per thread.
void       hm_emit_set_current(HmEmitter* e);
void       hm_emit_release_slot(TComPic* pic);   // no-op without a current emitter
HmEmitter* hm_emit_current();

#endif
