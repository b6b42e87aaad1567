// hm_fast.cpp — allocation- and memset-free picture turnover for HM's parser ("allocation-free slice setup").
//
// Stock HM destroys and re-creates the whole TComPic for EVERY picture (TDecTop::xGetNewPicBuffer, TDecTop.cpp:187-189:
// ~70 000 malloc/free pairs and ~28 MB of sample planes at 2160p) and zeroes ~49 KB of coefficient storage per CTU twice
// (TComDataCU::create, TComDataCU.cpp:173, and TComDataCU::initCU, :453) — together ~100 MB of cold memset per 2160p
// picture, about 45 % of the parser's run time.  Neither is needed once reconstruction lives on the GPU:
//
//  * hm_fast_memset: frontend/Makefile force-includes hm_fast_memset.h into TComDataCU.cpp.  While the calling thread
//    runs the product path (hm_fast_set_skip_coeff_fill), zero-fills of >= 4 KB — only the coefficient arrays are that
//    large — are skipped, at creation and per CTU: the pages are never even touched.  What the parser really needs,
//    "the block I am about to parse is zero" (TDecSbac::parseCoeffNxN writes only the scanned positions), is
//    provided by a one-line build-time patch of parseCoeffNxN (frontend/Makefile) that zeroes exactly that block.
//    Coefficient blocks of TUs without coded levels are therefore undefined; every consumer checks the cbf first.
//  * TDecTop::xGetNewPicBuffer (frontend/Makefile renames HM's own definition): same buffer selection rule as HM, but
//    a picture buffer of unchanged geometry is RESET (slices, SEIs, flags) instead of destroyed and re-created.  HM's
//    per-CTU state is fully re-initialised by TComDataCU::initCU for every CTU of every picture anyway.
//
//  * sample planes: a sink may install an allocator (hm_fast_set_plane_allocator); the planes of every picture buffer
//    then come from it instead of malloc.  gpu_sink.cpp hands out page-locked memory from a process-wide pool, so the
//    GPU DMAs finished pictures straight into the planes libHMDEC_get_image_plane returns, with no per-decoder
//    registration cost.  Plane pointers of a DPB entry stay stable for the life of the decoder.
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
#include <thread>
#include <deque>
#include <mutex>
#define private public
#define protected public
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicSym.h"
#include "TLibDecoder/TDecTop.h"
#undef private
#undef protected
#include "hm_emit.h"
#include "hm_fast.h"

static thread_local bool t_skipCoeffFill = false;

void hm_fast_set_skip_coeff_fill(bool on) { t_skipCoeffFill = on; }

extern "C" void* hm_fast_memset(void* p, int v, size_t n)
{
  if (v == 0 && n >= 4096 && t_skipCoeffFill) return p;
  return memset(p, v, n);
}

// ---- sample planes from a caller-provided (page-locked) allocator -------------------------------------------------
static HmPlaneAlloc g_planeAlloc = NULL;
static HmPlaneFree  g_planeFree = NULL;
static std::mutex g_planeLock;
static std::set<void*> g_planeOwned;       // plane buffers that came from g_planeAlloc (any decoder of the process)

void hm_fast_set_plane_allocator(HmPlaneAlloc a, HmPlaneFree f)
{
  std::lock_guard<std::mutex> g(g_planeLock);
  g_planeAlloc = a; g_planeFree = f;
}

bool hm_fast_plane_is_pinned(const void* buf)
{
  std::lock_guard<std::mutex> g(g_planeLock);
  return g_planeOwned.count((void*)buf) != 0;
}

// Swap the malloc'ed planes of a freshly created TComPicYuv for allocator-owned ones (same size, same origin offset).
static void adoptPlanes(TComPicYuv* yuv)
{
  if (!g_planeAlloc || !yuv) return;
  for (UInt c = 0; c < yuv->getNumberValidComponents(); c++)
  {
    const ComponentID id = ComponentID(c);
    Pel* old = yuv->m_apiPicBuf[c];
    if (!old) continue;
    const size_t bytes = (size_t)yuv->getStride(id) * yuv->getTotalHeight(id) * sizeof(Pel);
    Pel* fresh = (Pel*)g_planeAlloc(bytes);
    if (!fresh) continue;
    const ptrdiff_t org = yuv->m_piPicOrg[c] - old;
    xFree(old);
    yuv->m_apiPicBuf[c] = fresh;
    yuv->m_piPicOrg[c] = fresh + org;
    std::lock_guard<std::mutex> g(g_planeLock);
    g_planeOwned.insert(fresh);
  }
}

// Give allocator-owned planes back before HM frees the picture (TComPicYuv::destroy would free() them).
static void releasePlanes(TComPicYuv* yuv)
{
  if (!yuv) return;
  for (UInt c = 0; c < MAX_NUM_COMPONENT; c++)
  {
    Pel* p = yuv->m_apiPicBuf[c];
    if (!p) continue;
    bool owned;
    { std::lock_guard<std::mutex> g(g_planeLock); owned = g_planeOwned.erase(p) != 0; }
    if (!owned) continue;
    if (g_planeFree) g_planeFree(p);
    yuv->m_apiPicBuf[c] = NULL;
    yuv->m_piPicOrg[c] = NULL;
  }
}

// ---- process-wide pool of complete picture buffers ------------------------------------------------------------
// A TComPic of a 2160p stream is ~70 000 allocations (2040 CTUs x ~35 arrays).  A decoder that ends parks its pictures here
// (planes attached) and the next decoder of the same geometry adopts them: opening a new bitstream costs no allocation.
// HMDEC_B200_NO_PIC_POOL=1 disables the pool (pictures are then freed with their decoder, as in stock HM).
struct PicKey
{
  int w, h, fmt; unsigned cuW, cuH, depth;
  bool operator<(const PicKey& o) const
  {
    if (w != o.w) return w < o.w; if (h != o.h) return h < o.h; if (fmt != o.fmt) return fmt < o.fmt;
    if (cuW != o.cuW) return cuW < o.cuW; if (cuH != o.cuH) return cuH < o.cuH; return depth < o.depth;
  }
};
// Entries remember the thread that parked them, and a thread takes its own back first, oldest first: a caller that opens one
// bitstream after the other on the same thread then finds every picture buffer in the role it had before (the buffer that held
// the intra picture holds it again), so no page of the ~49 MB of per-CTU level arrays of a 2160p buffer is touched for the
// first time after the first bitstream.  With buffers wandering between threads every intra picture kept landing in a
// buffer that had only held B pictures: 1 650 page faults per intra picture, i.e. the memory-map lock, in what should be
// steady state (32 decoder threads: 190 faults per picture, the harness idle for up to a quarter of its time).
struct ParkedPic { TComPic* pic; std::thread::id owner; };
static std::mutex g_picPoolLock;
static std::multimap<PicKey, ParkedPic> g_picPool;

static PicKey keyOfSps(TComSPS* sps)
{
  PicKey k = { (int)sps->getPicWidthInLumaSamples(), (int)sps->getPicHeightInLumaSamples(), (int)sps->getChromaFormatIdc(), g_uiMaxCUWidth, g_uiMaxCUHeight, g_uiMaxCUDepth };
  return k;
}
static PicKey keyOfPic(TComPic* pic)
{
  TComPicYuv* rec = pic->getPicYuvRec();
  TComPicSym* sym = pic->m_apcPicSym;
  PicKey k = { rec->getWidth(COMPONENT_Y), rec->getHeight(COMPONENT_Y), (int)rec->getChromaFormat(), sym->m_uiMaxCUWidth, sym->m_uiMaxCUHeight, (unsigned)sym->m_uhTotalDepth };
  return k;
}

static bool poolEnabled() { static const bool on = getenv("HMDEC_B200_NO_PIC_POOL") == NULL; return on; }

// A parked picture of this geometry, made safe for the calling thread, or NULL.
static TComPic* takeFromPool(TComSPS* sps)
{
  if (!poolEnabled()) return NULL;
  TComPic* pic = NULL;
  {
    std::lock_guard<std::mutex> g(g_picPoolLock);
    typedef std::multimap<PicKey, ParkedPic>::iterator It;
    std::pair<It, It> range = g_picPool.equal_range(keyOfSps(sps));
    if (range.first == range.second) return NULL;
    It it = range.first;                                      // oldest entry of this geometry ...
    const std::thread::id me = std::this_thread::get_id();
    for (It k = range.first; k != range.second; ++k) if (k->second.owner == me) { it = k; break; }   // ... or this thread's own oldest
    pic = it->second.pic;
    g_picPool.erase(it);
  }
  // The CTUs' ARL coefficient pointers alias a per-THREAD global buffer of the thread that created them
  // (TComDataCU.cpp:178-181, made thread_local by frontend/Makefile): re-point them at this thread's buffer.
  TComPicSym* sym = pic->m_apcPicSym;
  for (UInt a = 0; a < sym->m_uiNumCUsInFrame; a++)
  {
    TComDataCU* cu = sym->m_apcTComDataCU[a];
    if (!cu->m_ArlCoeffIsAliasedAllocation) continue;
    for (UInt c = 0; c < MAX_NUM_COMPONENT; c++)
    {
      if (!cu->m_pcArlCoeff[c]) continue;
      if (!TComDataCU::m_pcGlbArlCoeff[c])
      {
        const UInt shift = getComponentScaleX(ComponentID(c), sps->getChromaFormatIdc()) + getComponentScaleY(ComponentID(c), sps->getChromaFormatIdc());
        TComDataCU::m_pcGlbArlCoeff[c] = (TCoeff*)xMalloc(TCoeff, (g_uiMaxCUWidth * g_uiMaxCUHeight) >> shift);
      }
      cu->m_pcArlCoeff[c] = TComDataCU::m_pcGlbArlCoeff[c];
    }
  }
  return pic;
}

// Every picture buffer a decoder ever created.  The wrapper's flush (like the reference's, libHMDecoder.cpp:329-336) only
// drops the pointers from the DPB list; without this registry those pictures would be lost and never reused.
static std::mutex g_createdLock;
static std::map<TDecTop*, std::vector<TComPic*> > g_created;

static void remember(TDecTop* dec, TComPic* pic)
{
  std::lock_guard<std::mutex> g(g_createdLock);
  g_created[dec].push_back(pic);
}

static bool sameGeometry(TComPic* pic, TComSPS* sps);

// A picture of this decoder that is no longer in its DPB list (dropped by a flush) and has the geometry of `sps`, or NULL.
// Orphans of ANOTHER geometry (the stream switched resolution at an IRAP) can never be reused by this decoder again: they are
// destroyed here, their planes go back to the pool and their DPB slot is released (HmEmitter::releaseSlot), so a stream that
// keeps switching neither leaks picture buffers nor runs out of slots.
static TComPic* findOrphan(TDecTop* dec, TComList<TComPic*>& list, TComSPS* sps)
{
  std::vector<TComPic*> stale;
  TComPic* hit = NULL;
  {
    std::lock_guard<std::mutex> g(g_createdLock);
    std::vector<TComPic*>& v = g_created[dec];
    for (size_t i = 0; i < v.size() && !hit;)
    {
      bool listed = false;
      for (TComList<TComPic*>::iterator it = list.begin(); it != list.end() && !listed; ++it) listed = (*it == v[i]);
      if (listed) { i++; continue; }
      if (sameGeometry(v[i], sps)) { hit = v[i]; break; }
      stale.push_back(v[i]);
      v.erase(v.begin() + i);
    }
  }
  if (!stale.empty())
    if (HmEmitter* e = hm_emit_current()) e->sink()->releaseHostBuffers();   // no DMA may still target the planes given back below
  for (size_t i = 0; i < stale.size(); i++)
  {
    TComPic* pic = stale[i];
    hm_emit_release_slot(pic);
    if (pic->getPicYuvRec()) releasePlanes(pic->getPicYuvRec());
    pic->destroy();
    delete pic;
  }
  return hit;
}

void hm_fast_release_decoder(TDecTop* dec)
{
  std::vector<TComPic*> mine;
  {
    std::lock_guard<std::mutex> g(g_createdLock);
    std::map<TDecTop*, std::vector<TComPic*> >::iterator it = g_created.find(dec);
    if (it == g_created.end()) return;
    mine.swap(it->second);
    g_created.erase(it);
  }
  for (size_t i = 0; i < mine.size(); i++)
  {
    TComPic* pic = mine[i];
    if (poolEnabled() && pic->m_apcPicSym && pic->getPicYuvRec())
    {
      // park the complete buffer (planes stay attached); HM's teardown must not see it any more
      for (TComList<TComPic*>::iterator it = dec->m_cListPic.begin(); it != dec->m_cListPic.end();)
        if (*it == pic) it = dec->m_cListPic.erase(it); else ++it;
      if (pic->m_SEIs.size() > 0) deleteSEIs(pic->m_SEIs);
      pic->m_apcPicSym->clearSliceBuffer();                  // slices point into the dying decoder's parameter sets
      std::lock_guard<std::mutex> g(g_picPoolLock);
      ParkedPic parked = { pic, std::this_thread::get_id() };
      g_picPool.insert(std::make_pair(keyOfPic(pic), parked));
      continue;
    }
    releasePlanes(pic->getPicYuvRec());
    bool listed = false;
    for (TComList<TComPic*>::iterator it = dec->m_cListPic.begin(); it != dec->m_cListPic.end() && !listed; ++it) listed = (*it == pic);
    if (!listed) { pic->destroy(); delete pic; }           // HM's own teardown only knows the pictures still in its list
  }
}

static void createPicture(TComPic* pic, TComSPS* sps, Window& conf, Window& disp, Int* reorder)
{
  pic->create(sps->getPicWidthInLumaSamples(), sps->getPicHeightInLumaSamples(), sps->getChromaFormatIdc(),
              g_uiMaxCUWidth, g_uiMaxCUHeight, g_uiMaxCUDepth, conf, disp, reorder, true);
  adoptPlanes(pic->getPicYuvRec());
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
    rpcPic = findOrphan(this, m_cListPic, sps);              // a buffer the last flush dropped from the list
    if (rpcPic)
    {
      rpcPic->setOutputMark(false); rpcPic->setReconMark(false);
      resetPicture(rpcPic, conf, disp, reorder);
    }
    else if ((rpcPic = takeFromPool(sps)) != NULL)
    {
      rpcPic->setOutputMark(false); rpcPic->setReconMark(false);
      resetPicture(rpcPic, conf, disp, reorder);
      remember(this, rpcPic);
    }
    else
    {
      rpcPic = new TComPic();
      createPicture(rpcPic, sps, conf, disp, reorder);
      remember(this, rpcPic);
    }
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
    if ((rpcPic = takeFromPool(sps)) != NULL)
    {
      rpcPic->setOutputMark(false); rpcPic->setReconMark(false);
      resetPicture(rpcPic, conf, disp, reorder);
    }
    else
    {
      rpcPic = new TComPic();
      createPicture(rpcPic, sps, conf, disp, reorder);
    }
    m_cListPic.pushBack(rpcPic);
    remember(this, rpcPic);
    return;
  }
  rpcPic = found;
  if (sameGeometry(rpcPic, sps)) resetPicture(rpcPic, conf, disp, reorder);
  else
  {
    releasePlanes(rpcPic->getPicYuvRec());
    rpcPic->destroy();
    createPicture(rpcPic, sps, conf, disp, reorder);
  }
}

// ---- CTU metadata prefetch -------------------------------------------------------------------------------------
// TComDataCU::initCU (TComDataCU.cpp:453) refills ~35 per-partition arrays plus both motion fields of a CTU — about
// 13 KB spread over as many separate allocations, last touched one whole picture ago, i.e. cold in every cache level.
// The sampling profile shows the parser thread stalled on exactly those stores (memset + initCU + clearMvField ≈ 17 %).
// While the emitter walks CTU n it therefore asks for the lines of CTU n+1 (HMDEC_B200_PF_DIST), a few per coding unit so that the
// requests trickle out between real work instead of queueing behind the core's dozen line-fill buffers.
static inline void pfAdd(HmPrefetchCursor& c, const void* p, size_t bytes)
{
  if (!p || c.count >= 48) return;
  c.base[c.count] = (const char*)p; c.lines[c.count] = (unsigned)((bytes + 63) >> 6); c.count++;
}

void hm_fast_prefetch_begin(HmPrefetchCursor& pf, TComPic* pic, unsigned ctuAddr)
{
  pf.count = pf.cur = 0; pf.line = 0;
  if (ctuAddr >= pic->getNumCUsInFrame()) return;
  TComDataCU* cu = pic->getCU(ctuAddr);
  const size_t n = cu->m_uiNumPartition;
  pfAdd(pf, cu->m_puhDepth, n);            pfAdd(pf, cu->m_pePartSize, n);     pfAdd(pf, cu->m_skipFlag, n);      pfAdd(pf, cu->m_pePredMode, n);
  pfAdd(pf, cu->m_CUTransquantBypass, n);  pfAdd(pf, cu->m_puhWidth, n);       pfAdd(pf, cu->m_puhHeight, n);     pfAdd(pf, cu->m_phQP, n);
  pfAdd(pf, cu->m_ChromaQpAdj, n);         pfAdd(pf, cu->m_puhTrIdx, n);       pfAdd(pf, cu->m_pbMergeFlag, n);   pfAdd(pf, cu->m_puhMergeIndex, n);
  pfAdd(pf, cu->m_puhInterDir, n);         pfAdd(pf, cu->m_pbIPCMFlag, n);
  for (int c = 0; c < MAX_NUM_COMPONENT; c++)
  {
    pfAdd(pf, cu->m_crossComponentPredictionAlpha[c], n); pfAdd(pf, cu->m_puhTransformSkip[c], n);
    pfAdd(pf, cu->m_explicitRdpcmMode[c], n);             pfAdd(pf, cu->m_puhCbf[c], n);
  }
  for (int c = 0; c < MAX_NUM_CHANNEL_TYPE; c++) pfAdd(pf, cu->m_puhIntraDir[c], n);
  for (int l = 0; l < NUM_REF_PIC_LIST_01; l++)
  {
    pfAdd(pf, cu->m_apiMVPIdx[l], n); pfAdd(pf, cu->m_apiMVPNum[l], n);
    TComCUMvField& f = cu->m_acCUMvField[l];
    pfAdd(pf, f.m_pcMv, n * sizeof(TComMv)); pfAdd(pf, f.m_piRefIdx, n);      // (not the MVD array: see initCU)
  }
}

// ---- per-CTU initialisation -----------------------------------------------------------------------------------------
// TComDataCU::initCU (TComDataCU.cpp:329-520, its definition is renamed initCU_hm at build time) runs before every CTU is
// parsed.  Two of its loops evaluate "does partition i lie behind the slice start" with half a dozen dependent loads per
// partition (the stores into the UInt arrays keep the compiler from hoisting them: ~5000 loads per CTU), and a CTU that starts
// inside the current slice segment — every CTU except the first of a segment that begins in mid-CTU, which no HEVC stream has —
// is initialised purely by fills.  That case is restated here with the bounds computed once; anything else takes HM's routine.
Void TComDataCU::initCU(TComPic* pcPic, UInt ctuAddr)
{
  TComSlice* slice = pcPic->getSlice(pcPic->getCurrSliceIdx());
  const UInt n = pcPic->getNumPartInCU();
  const UInt firstPart = pcPic->getPicSym()->getInverseCUOrderMap(ctuAddr) * n;        // this CTU's first partition in coding order
  if (slice->getSliceSegmentCurStartCUAddr() > firstPart || slice->getSliceCurStartCUAddr() > firstPart)
  {
    initCU_hm(pcPic, ctuAddr);
    return;
  }
  m_pcPic = pcPic;
  m_pcSlice = slice;
  m_uiCUAddr = ctuAddr;
  m_uiCUPelX = (ctuAddr % pcPic->getFrameWidthInCU()) * g_uiMaxCUWidth;
  m_uiCUPelY = (ctuAddr / pcPic->getFrameWidthInCU()) * g_uiMaxCUHeight;
  m_uiAbsIdxInLCU = 0;
  m_dTotalCost = MAX_DOUBLE;
  m_uiTotalDistortion = 0;
  m_uiTotalBits = 0;
  m_uiTotalBins = 0;
  m_uiNumPartition = n;
  std::fill_n(m_sliceStartCU, n, slice->getSliceCurStartCUAddr());
  std::fill_n(m_sliceSegmentStartCU, n, slice->getSliceSegmentCurStartCUAddr());

  struct Fill { void* p; int v; size_t elem; };
  const Fill fills[] = {
    { m_skipFlag, 0, sizeof(*m_skipFlag) },                         { m_pePartSize, NUMBER_OF_PART_SIZES, sizeof(*m_pePartSize) },
    { m_pePredMode, NUMBER_OF_PREDICTION_MODES, sizeof(*m_pePredMode) }, { m_CUTransquantBypass, 0, sizeof(*m_CUTransquantBypass) },
    { m_puhDepth, 0, sizeof(*m_puhDepth) },                         { m_puhTrIdx, 0, sizeof(*m_puhTrIdx) },
    { m_puhWidth, (int)g_uiMaxCUWidth, sizeof(*m_puhWidth) },       { m_puhHeight, (int)g_uiMaxCUHeight, sizeof(*m_puhHeight) },
    { m_apiMVPIdx[0], -1, sizeof(*m_apiMVPIdx[0]) },                { m_apiMVPNum[0], -1, sizeof(*m_apiMVPNum[0]) },
    { m_apiMVPIdx[1], -1, sizeof(*m_apiMVPIdx[1]) },                { m_apiMVPNum[1], -1, sizeof(*m_apiMVPNum[1]) },
    { m_phQP, slice->getSliceQp(), sizeof(*m_phQP) },               { m_ChromaQpAdj, 0, sizeof(*m_ChromaQpAdj) },
    { m_pbMergeFlag, 0, sizeof(*m_pbMergeFlag) },                   { m_puhMergeIndex, 0, sizeof(*m_puhMergeIndex) },
    { m_puhIntraDir[0], DC_IDX, sizeof(*m_puhIntraDir[0]) },        { m_puhIntraDir[1], 0, sizeof(*m_puhIntraDir[1]) },
    { m_puhInterDir, 0, sizeof(*m_puhInterDir) },                   { m_pbIPCMFlag, 0, sizeof(*m_pbIPCMFlag) } };
  for (size_t i = 0; i < sizeof(fills) / sizeof(fills[0]); i++) ::memset(fills[i].p, fills[i].v, n * fills[i].elem);
  for (UInt c = 0; c < MAX_NUM_COMPONENT; c++)
  {
    ::memset(m_crossComponentPredictionAlpha[c], 0, n * sizeof(*m_crossComponentPredictionAlpha[c]));
    ::memset(m_puhTransformSkip[c], 0, n * sizeof(*m_puhTransformSkip[c]));
    ::memset(m_puhCbf[c], 0, n * sizeof(*m_puhCbf[c]));
    ::memset(m_explicitRdpcmMode[c], NUMBER_OF_RDPCM_MODES, n * sizeof(*m_explicitRdpcmMode[c]));
  }
  // coefficient storage: HM zeroes all of it here (TComDataCU.cpp:453); same rule as hm_fast_memset
  {
    const UInt numCoeffY = g_uiMaxCUWidth * g_uiMaxCUHeight;
    for (UInt c = 0; c < MAX_NUM_COMPONENT; c++)
    {
      const UInt shift = m_pcPic->getComponentScaleX(ComponentID(c)) + m_pcPic->getComponentScaleY(ComponentID(c));
      hm_fast_memset(m_pcTrCoeff[c], 0, sizeof(TCoeff) * numCoeffY >> shift);
#if ADAPTIVE_QP_SELECTION
      hm_fast_memset(m_pcArlCoeff[c], 0, sizeof(TCoeff) * numCoeffY >> shift);
#endif
    }
  }
  for (UInt l = 0; l < NUM_REF_PIC_LIST_01; l++)
  {
    TComCUMvField& f = m_acCUMvField[l];                      // clearMvField (TComMotionInfo.cpp:88-97): zero vectors, refIdx NOT_VALID
    ::memset(f.m_pcMv, 0, sizeof(TComMv) * f.m_uiNumPartition);
    // m_pcMvd: not cleared — the decoder writes an MVD at a PU's first partition and reads it back from there at once, nothing else (frontend/Makefile)
    ::memset(f.m_piRefIdx, NOT_VALID, sizeof(*f.m_piRefIdx) * f.m_uiNumPartition);
  }
  // neighbours (TComDataCU.cpp:497-540)
  const UInt widthInCtus = pcPic->getFrameWidthInCU();
  const UInt col = ctuAddr % widthInCtus;
  m_pcCULeft = col ? pcPic->getCU(ctuAddr - 1) : NULL;
  m_pcCUAbove = ctuAddr >= widthInCtus ? pcPic->getCU(ctuAddr - widthInCtus) : NULL;
  m_pcCUAboveLeft = (m_pcCULeft && m_pcCUAbove) ? pcPic->getCU(ctuAddr - widthInCtus - 1) : NULL;
  m_pcCUAboveRight = (m_pcCUAbove && col < widthInCtus - 1) ? pcPic->getCU(ctuAddr - widthInCtus + 1) : NULL;
  for (UInt l = 0; l < NUM_REF_PIC_LIST_01; l++)
  {
    const RefPicList rpl = RefPicList(l);
    m_apcCUColocated[rpl] = slice->getNumRefIdx(rpl) > 0 ? slice->getRefPic(rpl, 0)->getCU(ctuAddr) : NULL;
  }
}
