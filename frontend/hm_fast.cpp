// hm_fast.cpp — allocation- and memset-free picture turnover for HM's parser ("allocation-free slice setup").
//
// Stock HM destroys and re-creates the whole TComPic for EVERY picture (TDecTop::xGetNewPicBuffer, TDecTop.cpp:187-189:
// ~70 000 malloc/free pairs and ~28 MB of sample planes at 2160p) and zeroes ~49 KB of coefficient storage per CTU twice
// (TComDataCU::create, TComDataCU.cpp:173, and TComDataCU::initCU, :453) — together ~100 MB of cold memset per 2160p
// picture, about 45 % of the parser's run time.  Neither is needed once reconstruction lives on the GPU:
//
//  * hm_fast_memset: frontend/Makefile compiles TComDataCU.cpp with -Dmemset=hm_fast_memset.  While the calling thread
//    has "coefficient hygiene" on (hm_fast_set_clean_coeffs), zero-fills of >= 4 KB — only the coefficient arrays are
//    that large — are skipped; the invariant "coefficient storage is all zero between pictures" is kept instead by the
//    record emitter, which clears each coded TU's levels right after copying them (hm_emit.cpp, cache-warm).
//  * TDecTop::xGetNewPicBuffer (frontend/Makefile renames HM's own definition): same buffer selection rule as HM, but
//    a picture buffer of unchanged geometry is RESET (slices, SEIs, flags) instead of destroyed and re-created.  HM's
//    per-CTU state is fully re-initialised by TComDataCU::initCU for every CTU of every picture anyway.
//
// Plane pointers of a DPB entry therefore stay stable for the life of the decoder, which is what lets gpu_sink.cpp
// page-lock them once and have the GPU DMA finished pictures straight into them.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <list>
#include <iostream>
#include <sstream>
#include <fstream>
#include <algorithm>
#include <limits>
#include <iomanip>
#include <cmath>
#include <cassert>
#include <map>
#include <set>
#include <deque>
#define private public
#define protected public
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicSym.h"
#include "TLibDecoder/TDecTop.h"
#undef private
#undef protected
#include "hm_fast.h"

static thread_local bool t_cleanCoeffs = false;    // emitter keeps coefficient storage zero between pictures
static thread_local bool t_creating = false;       // inside TComPic::create: HM's initial zero-fill must happen

void hm_fast_set_clean_coeffs(bool on) { t_cleanCoeffs = on; }
bool hm_fast_clean_coeffs() { return t_cleanCoeffs; }

extern "C" void* hm_fast_memset(void* p, int v, size_t n)
{
  if (v == 0 && n >= 4096 && t_cleanCoeffs && !t_creating) return p;
  return memset(p, v, n);
}

static void createPicture(TComPic* pic, TComSPS* sps, Window& conf, Window& disp, Int* reorder)
{
  t_creating = true;
  pic->create(sps->getPicWidthInLumaSamples(), sps->getPicHeightInLumaSamples(), sps->getChromaFormatIdc(),
              g_uiMaxCUWidth, g_uiMaxCUHeight, g_uiMaxCUDepth, conf, disp, reorder, true);
  t_creating = false;
}

static bool sameGeometry(TComPic* pic, TComSPS* sps)
{
  TComPicSym* sym = pic->m_apcPicSym;
  TComPicYuv* rec = pic->getPicYuvRec();
  if (!sym || !rec) return false;
  return rec->getWidth(COMPONENT_Y) == sps->getPicWidthInLumaSamples() && rec->getHeight(COMPONENT_Y) == sps->getPicHeightInLumaSamples() &&
         rec->getChromaFormat() == sps->getChromaFormatIdc() && sym->m_uiMaxCUWidth == g_uiMaxCUWidth && sym->m_uiMaxCUHeight == g_uiMaxCUHeight &&
         sym->m_uhTotalDepth == g_uiMaxCUDepth;
}

// What TComPic::create / TComPicSym::create leave behind, minus the allocations (TComPic.cpp:70-98, TComPicSym.cpp:84-127)
static void resetPicture(TComPic* pic, Window& conf, Window& disp, Int* reorder)
{
  TComPicSym* sym = pic->m_apcPicSym;
  sym->clearSliceBuffer();
  delete sym->getSlice(0);
  sym->setSlice(new TComSlice, 0);
  for (UInt i = 0; i < sym->m_uiNumCUsInFrame; i++) { sym->m_puiCUOrderMap[i] = i; sym->m_puiInverseCUOrderMap[i] = i; }
  if (pic->m_SEIs.size() > 0) deleteSEIs(pic->m_SEIs);
  pic->m_bUsedByCurr = false;
  pic->m_conformanceWindow = conf;
  pic->m_defaultDisplayWindow = disp;
  memcpy(pic->m_numReorderPics, reorder, MAX_TLAYER * sizeof(Int));
  pic->getPicYuvRec()->setBorderExtension(false);
}

Void TDecTop::xGetNewPicBuffer(TComSlice* pcSlice, TComPic*& rpcPic)
{
  TComSPS* sps = pcSlice->getSPS();
  Int reorder[MAX_TLAYER];
  for (Int t = 0; t < MAX_TLAYER; t++) reorder[t] = sps->getNumReorderPics(t);
  Window& conf = sps->getConformanceWindow();
  Window disp = sps->getVuiParametersPresentFlag() ? sps->getVuiParameters()->getDefaultDisplayWindow() : Window();

  m_iMaxRefPicNum = sps->getMaxDecPicBuffering(pcSlice->getTLayer());   // includes the picture being decoded
  if (m_cListPic.size() < (UInt)m_iMaxRefPicNum)
  {
    rpcPic = new TComPic();
    createPicture(rpcPic, sps, conf, disp, reorder);
    m_cListPic.pushBack(rpcPic);
    return;
  }

  // HM's rule (TDecTop.cpp:158-178): first entry that is neither awaiting output nor (reconstructed and referenced)
  TComPic* found = NULL;
  for (TComList<TComPic*>::iterator it = m_cListPic.begin(); it != m_cListPic.end() && !found; ++it)
  {
    TComPic* p = *it;
    if (!p->getReconMark() && !p->getOutputMark()) { p->setOutputMark(false); found = p; }
    else if (!p->getSlice(0)->isReferenced() && !p->getOutputMark()) { p->setOutputMark(false); p->setReconMark(false); found = p; }
  }
  if (!found)
  {
    // no room (faulty encoder or dropped NAL): extend the buffer
    m_iMaxRefPicNum++;
    rpcPic = new TComPic();
    m_cListPic.pushBack(rpcPic);
    createPicture(rpcPic, sps, conf, disp, reorder);
    return;
  }
  rpcPic = found;
  if (sameGeometry(rpcPic, sps)) resetPicture(rpcPic, conf, disp, reorder);
  else
  {
    rpcPic->destroy();
    createPicture(rpcPic, sps, conf, disp, reorder);
  }
}
