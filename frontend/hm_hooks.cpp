// hm_hooks.cpp — the two splice points of the GPU reconstruction path inside HM's decoder:
//   * TDecCu::decompressCU   (called from TDecSlice.cpp:334 right after each CTU is parsed)
//   * TDecGop::filterPicture (called from TDecTop::executeLoopFilters, TDecTop.cpp:202)
// frontend/Makefile compiles HM's TDecCu.cpp / TDecGop.cpp with these two member functions renamed
// (-DdecompressCU=decompressCU_hm, -DfilterPicture=filterPicture_hm); the definitions below take their
// place at link time.  HM's sources are not modified.
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <list>
#include <iostream>
#include "TLibDecoder/TDecCu.h"
#include "TLibDecoder/TDecGop.h"
#include "TLibCommon/TComLoopFilter.h"
#include "TLibCommon/TComSampleAdaptiveOffset.h"
#include "TLibCommon/SEI.h"
#include "hm_emit.h"
#include "hm_fast.h"

extern thread_local Bool g_md5_mismatch;
#ifdef HMDEC_WITH_HM_RECON
void hm_call_original_decompressCU(TDecCu* dec, TComDataCU* ctu);     // hm_shim.cpp (libHMDecoder_b200_verify.so only)
#endif

static HmEmitter* requireEmitter()
{
  HmEmitter* e = hm_emit_current();
  if (!e)
  {
    fprintf(stderr, "hm_hooks: no HmEmitter bound to this thread (decoder must be driven through libHMDec_*)\n");
    abort();
  }
  return e;
}

Void TDecCu::decompressCU(TComDataCU* pcCU)
{
  HmEmitter* e = requireEmitter();
  e->onCtuParsed(pcCU);
#ifdef HMDEC_WITH_HM_RECON
  if (e->sink()->wantHmRecon()) hm_call_original_decompressCU(this, pcCU);
#endif
}

// Status line + SEI hash check, same text as TDecGop::filterPicture / calcAndPrintHashStatus (TDecGop.cpp:176-289).
// MD5 (the default SEI method) is checked asynchronously on the device when the sink offers it: the line is completed
// and printed by the sink when the digest arrives (lines keep their order).
static void appendf(std::string& s, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  s += buf;
}

static void printStatusAndHash(TComPic* pic, TComSlice* slice, Int hashEnabled, HmEmitter* e, bool quiet)
{
  std::string line;
  Char c = (slice->isIntra() ? 'I' : slice->isInterP() ? 'P' : 'B');
  if (!slice->isReferenced()) c += 32;
  appendf(line, "POC %4d TId: %1d ( %c-SLICE, QP%3d ) ", slice->getPOC(), slice->getTLayer(), c, slice->getSliceQp());
  appendf(line, "[DT %6.3f] ", 0.0);
  for (Int l = 0; l < 2; l++)
  {
    appendf(line, "[L%d ", l);
    for (Int i = 0; i < slice->getNumRefIdx(RefPicList(l)); i++) appendf(line, "%d ", slice->getRefPOC(RefPicList(l), i));
    appendf(line, "] ");
  }
  if (hashEnabled)
  {
    SEIMessages hashes = getSeisByType(pic->getSEIs(), SEI::DECODED_PICTURE_HASH);
    const SEIDecodedPictureHash* hash = hashes.size() > 0 ? (SEIDecodedPictureHash*)*(hashes.begin()) : NULL;
    if (hashes.size() > 1) appendf(line, "Warning: Got multiple decoded picture hash SEI messages. Using first.");
    TComDigest digest; Int numChar = 0; const Char* type = "\0";
    if (hash)
    {
      uint32_t dv[3];
      const int ncomp = pic->getNumberValidComponents();
      const int method = hash->method == SEIDecodedPictureHash::CRC ? 2 : (hash->method == SEIDecodedPictureHash::CHECKSUM ? 3 : 1);
      if (method == 1 && (int)hash->m_digest.hash.size() == 16 * ncomp && e->sink()->asyncMd5(pic, hash->m_digest.hash.data(), ncomp, line, quiet))
        return;                                                    // the sink finishes the line
      if (method != 1 && e->sink()->deviceHash(pic, method, dv))
      {
        // CRC / checksum are computed on the device (no plane transfer); digest bytes are big-endian (TComPicYuvMD5.cpp:120-121,157-160)
        type = method == 2 ? "CRC" : "Checksum";
        numChar = method == 2 ? 2 : 4;
        digest.hash.clear();
        for (int c = 0; c < ncomp; c++)
          for (int b = numChar - 1; b >= 0; b--) digest.hash.push_back((UChar)((dv[c] >> (8 * b)) & 0xff));
      }
      else
      {
        e->sink()->fetchPicture(pic);      // D2H of the reconstructed planes into HM's TComPicYuv
        TComPicYuv& rec = *pic->getPicYuvRec();
        switch (hash->method)
        {
          case SEIDecodedPictureHash::MD5:      type = "MD5";      numChar = calcMD5(rec, digest); break;
          case SEIDecodedPictureHash::CRC:      type = "CRC";      numChar = calcCRC(rec, digest); break;
          case SEIDecodedPictureHash::CHECKSUM: type = "Checksum"; numChar = calcChecksum(rec, digest); break;
          default: break;
        }
      }
    }
    const Char* ok = "(unk)"; Bool mismatch = false;
    if (hash) { ok = "(OK)"; if (digest != hash->m_digest) { ok = "(***ERROR***)"; mismatch = true; } }
    appendf(line, "[%s:%s,%s] ", type, digestToString(digest, numChar).c_str(), ok);
    if (mismatch)
    {
      g_md5_mismatch = true;
      appendf(line, "[rx%s:%s] ", type, digestToString(hash->m_digest, numChar).c_str());
    }
  }
  line += "\n";
  if (!quiet) e->sink()->orderedPrint(line);
}

// TDecTop::xCreateLostPicture (TDecTop.cpp:233-281, patched at build time to call this right after its host-side copy)
void hm_hook_lost_picture(TComPic* fill, TComPic* src, int poc)
{
  HmEmitter* e = requireEmitter();
  e->onLostPicture(fill, src, poc);
  if (e->sink()->wantHmRecon()) for (int stage = 0; stage < 3; stage++) e->sink()->hmStage(stage, fill);   // golden dumps: HM's copy is the picture at every stage
}

Void TDecGop::filterPicture(TComPic*& rpcPic)
{
  HmEmitter* e = requireEmitter();
  TComSlice* slice = rpcPic->getSlice(rpcPic->getCurrSliceIdx());
  const Bool lfCrossTiles = slice->getPPS()->getLoopFilterAcrossTilesEnabledFlag();

  // boundary strengths need the full-resolution motion field => before compressMotion()
  e->onPictureParsed(rpcPic, m_pcLoopFilter, m_pcSAO, lfCrossTiles);

#ifdef HMDEC_WITH_HM_RECON
  if (e->sink()->wantHmRecon())
  {
    // verification / golden generation: HM's own CPU filters, stage by stage (TDecGop.cpp:165-174)
    e->sink()->hmStage(0, rpcPic);
    m_pcLoopFilter->setCfg(lfCrossTiles);
    m_pcLoopFilter->loopFilterPic(rpcPic);
    e->sink()->hmStage(1, rpcPic);
    if (slice->getSPS()->getUseSAO())
    {
      // SAO parameters were already merged/scaled by the emitter (reconstructBlkSAOParams is not idempotent)
      m_pcSAO->SAOProcess(rpcPic);
      m_pcSAO->PCMLFDisableProcess(rpcPic);
    }
    e->sink()->hmStage(2, rpcPic);
  }
#endif

  // TMVP storage (TDecGop.cpp:176, TComPic::compressMotion): never rewritten — xGetColMVP and libHMDEC_get_internal_info read
  // the uncompressed field at the 16x16 run start (hm_fast_memset.h: hm_fast_col_part; internals.cpp)
  static const bool quiet = getenv("HMDEC_B200_QUIET") != NULL;   // the hash is still verified
  printStatusAndHash(rpcPic, slice, m_decodedPictureHashSEIEnabled, e, quiet);
  e->sink()->drainHashes(false);
  if (e->sink()->hashMismatchSeen()) g_md5_mismatch = true;

#if SETTING_PIC_OUTPUT_MARK
  rpcPic->setOutputMark(rpcPic->getSlice(0)->getPicOutputFlag() ? true : false);
#else
  rpcPic->setOutputMark(true);
#endif
  rpcPic->setReconMark(true);
}
