// hm_emit.cpp — frame-record emitter (host side of the drop-in boundary).
//
// Walks HM's per-CTU data exactly the way HM's own reconstruction walks it, but instead of
// computing samples it appends plain records (include/hmr_records.h).  Each function names the HM
// routine whose traversal it mirrors.  Compiled against the HM headers in /root/reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cassert>
#include <cmath>
#include <string>
#include <vector>
#include <mutex>
#include <list>
#include <map>
#include <iostream>
#include <sstream>
#include <fstream>
#include <algorithm>
#include <limits>
#include <iomanip>

// The boundary-strength derivation is reused from HM (TComLoopFilter::xGetBoundaryStrengthSingle and
// friends are protected, their result arrays private).  We only *read* them.
#define private public
#define protected public
#include "TLibCommon/TComLoopFilter.h"
#undef private
#undef protected
#include "TLibCommon/TComSampleAdaptiveOffset.h"
#include "TLibCommon/TComPrediction.h"
#include "TLibCommon/TComTrQuant.h"
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComTU.h"
#include "TLibCommon/TComRom.h"

#include "hm_emit.h"
#include "hm_fast.h"
#include <chrono>
static inline double nowSec() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// free functions of TComPattern.cpp:558-735 (defined there, not declared in any header)
Bool isAboveLeftAvailable ( TComDataCU* pcCU, UInt uiPartIdxLT );
Int  isAboveAvailable     ( TComDataCU* pcCU, UInt uiPartIdxLT, UInt uiPartIdxRT, Bool* bValidFlags );
Int  isLeftAvailable      ( TComDataCU* pcCU, UInt uiPartIdxLT, UInt uiPartIdxLB, Bool* bValidFlags );
Int  isAboveRightAvailable( TComDataCU* pcCU, UInt uiPartIdxLT, UInt uiPartIdxRT, Bool* bValidFlags );
Int  isBelowLeftAvailable ( TComDataCU* pcCU, UInt uiPartIdxLT, UInt uiPartIdxLB, Bool* bValidFlags );

static thread_local HmEmitter* t_current = NULL;
void       hm_emit_set_current(HmEmitter* e) { t_current = e; }
HmEmitter* hm_emit_current()                 { return t_current; }
void       hm_emit_release_slot(TComPic* pic) { if (t_current) t_current->releaseSlot(pic); }

struct HmEmitter::CuCtx
{
  TComDataCU* ctu;
  unsigned    absPartIdx;   // z-order index of the CU inside the CTU
  unsigned    depth;
  int         cuX, cuY;     // luma position of the CU
  uint32_t    lumaOff[256]; // coef/resid offset of the luma TU starting at partition i (CCP), HMR_NO_OFFSET if none
};

// ---- record storage survives the decoder -------------------------------------------------------------------------------------
// A new bitstream means a new decoder and a new emitter.  Its record vectors would grow from nothing again — the level arena of a
// 2160p intra picture alone doubles twelve times up to 24 MB: 48 MB of fresh pages, copies and mmap / munmap calls per bitstream,
// 190 page faults per picture in steady state, and every one of them takes the process's memory-map lock, which is what made the
// 32-thread harness idle for up to a quarter of its time (profiles/r04d_e2e_thp.log).  An emitter that ends parks its (emptied)
// vectors here with their capacity; the next one adopts them.
namespace {
struct RecordStorage
{
  std::vector<hmr_tu> tu, tuSorted; HmLevelArena coef; std::vector<hmr_intra> intra, intraTmp[3];
  std::vector<hmr_ctu_intra_range> range; std::vector<hmr_pu> pu; std::vector<uint32_t> puPrefix; std::vector<hmr_ctu> ctu;
  std::vector<uint8_t> bs, cuFlags, puRefIdx; std::vector<int8_t> qp;
};
// never destroyed: a decoder freed from a static destructor of the host program must still find them
std::mutex& g_storageLock = *new std::mutex;
std::vector<RecordStorage*>& g_storage = *new std::vector<RecordStorage*>;
bool storagePooled() { static const bool on = getenv("HMDEC_B200_NO_RECORD_POOL") == NULL; return on; }
}

// emptied vectors <-> the emitter's members (swap in both directions)
void HmEmitter::swapStorage(void* p)
{
  RecordStorage& s = *(RecordStorage*)p;
  m_tu.swap(s.tu); m_tuSorted.swap(s.tuSorted); m_coef.swap(s.coef); m_intra.swap(s.intra);
  for (int c = 0; c < 3; c++) m_intraTmp[c].swap(s.intraTmp[c]);
  m_range.swap(s.range); m_pu.swap(s.pu); m_puPrefix.swap(s.puPrefix); m_ctu.swap(s.ctu);
  m_bs.swap(s.bs); m_cuFlags.swap(s.cuFlags); m_puRefIdx.swap(s.puRefIdx); m_qp.swap(s.qp);
}

HmEmitter::HmEmitter(HmFrameSink* sink)
  : m_sink(sink), m_lf(NULL), m_lfDepth(0), m_anyDeblock(false), m_curPic(NULL), m_open(false), m_unsupported(NULL), m_bsStride(0), m_qpStride(0), m_lgU(0), m_tCtu(0), m_tBs(0), m_tPic(0), m_tSink(0), m_nPic(0)
{
  m_in422SubTu = false;
  // product path: HM's whole-CTU coefficient zero fills are skipped (hm_fast.cpp); verification / golden generation keep them
  m_cleanCoeffs = !sink->wantHmRecon();
  memset(&m_hdr, 0, sizeof(m_hdr));
  RecordStorage* st = NULL;
  if (storagePooled())
  {
    std::lock_guard<std::mutex> g(g_storageLock);
    if (!g_storage.empty()) { st = g_storage.back(); g_storage.pop_back(); }
  }
  if (st) { swapStorage(st); delete st; }
}

HmEmitter::~HmEmitter()
{
  if (storagePooled())
  {
    m_tu.clear(); m_tuSorted.clear(); m_coef.clear(); m_intra.clear(); m_range.clear(); m_pu.clear(); m_puPrefix.clear(); m_ctu.clear();
    m_bs.clear(); m_cuFlags.clear(); m_puRefIdx.clear(); m_qp.clear();
    for (int c = 0; c < 3; c++) m_intraTmp[c].clear();
    RecordStorage* st = new RecordStorage;
    swapStorage(st);
    std::lock_guard<std::mutex> g(g_storageLock);
    if (g_storage.size() < 128) { g_storage.push_back(st); st = NULL; }
    delete st;
  }
  if (m_lf) { m_lf->destroy(); delete m_lf; }
  if (getenv("HMDEC_B200_STATS") && m_nPic)
    fprintf(stderr, "hm_emit stats: %d pictures; per picture: CTU record emission %.2f ms, BS/QP maps %.2f ms, SAO+pack %.2f ms, sink submit %.2f ms\n",
            m_nPic, 1e3 * m_tCtu / m_nPic, 1e3 * m_tBs / m_nPic, 1e3 * m_tPic / m_nPic, 1e3 * m_tSink / m_nPic);
}

void HmEmitter::fail(const char* what)
{
  if (!m_unsupported)
  {
    m_unsupported = what;
    fprintf(stderr, "hm_emit: UNSUPPORTED bitstream feature on the GPU reconstruction path: %s\n", what);
  }
}

// DPB slot of a picture buffer.  A slot lives as long as the TComPic it was handed to: hm_fast.cpp reports every buffer it
// destroys (stale geometry after a resolution switch) through releaseSlot(), and the number goes back on a free list.
// More than HMR_MAX_SLOTS live buffers is a hard decode error (the decoder stops; LIBHMDEC_ERROR), never an alias.
int HmEmitter::slotOf(TComPic* pic)
{
  std::map<TComPic*, int>::iterator it = m_slots.find(pic);
  if (it != m_slots.end()) return it->second;
  int s;
  if (!m_freeSlots.empty()) { s = m_freeSlots.back(); m_freeSlots.pop_back(); }
  else s = (int)m_slots.size();
  if (s >= HMR_MAX_SLOTS) { fail("more than HMR_MAX_SLOTS DPB entries"); return HMR_MAX_SLOTS - 1; }   // not recorded: nothing is submitted any more
  m_slots[pic] = s;
  return s;
}

void HmEmitter::releaseSlot(TComPic* pic)
{
  std::map<TComPic*, int>::iterator it = m_slots.find(pic);
  if (it == m_slots.end()) return;
  m_freeSlots.push_back(it->second);
  m_slots.erase(it);
}

// ---------------------------------------------------------------------------------------------
// picture start: frame header from SPS/PPS/slice (TDecTop::xActivateParameterSets, TDecTop.cpp:283-349)
void HmEmitter::beginFrame(TComPic* pic, TComDataCU* ctu)
{
  TComSlice* slice = ctu->getSlice();
  TComSPS*   sps   = slice->getSPS();
  TComPPS*   pps   = slice->getPPS();
  memset(&m_hdr, 0, sizeof(m_hdr));
  m_anyDeblock = false;
  m_hdr.magic   = HMR_MAGIC;
  m_hdr.version = HMR_VERSION;
  m_hdr.width   = sps->getPicWidthInLumaSamples();
  m_hdr.height  = sps->getPicHeightInLumaSamples();
  m_hdr.poc     = slice->getPOC();
  m_hdr.chroma_format    = (uint8_t)pic->getChromaFormat();
  m_hdr.bit_depth_luma   = (uint8_t)g_bitDepth[CHANNEL_TYPE_LUMA];
  m_hdr.bit_depth_chroma = (uint8_t)g_bitDepth[CHANNEL_TYPE_CHROMA];
  m_hdr.log2_ctu = (uint8_t)(g_aucConvertToBit[g_uiMaxCUWidth] + 2);
  m_hdr.out_slot = (uint8_t)slotOf(pic);
  m_hdr.slice_type = (uint8_t)slice->getSliceType();
  m_hdr.pps_cb_qp_offset = (int8_t)pps->getQpOffset(COMPONENT_Cb);
  m_hdr.pps_cr_qp_offset = (int8_t)pps->getQpOffset(COMPONENT_Cr);
  m_hdr.n_ctu = pic->getNumCUsInFrame();
  if (sps->getUseStrongIntraSmoothing())     m_hdr.flags |= HMR_FRM_STRONG_INTRA_SMOOTHING;
  if (pps->getUseCrossComponentPrediction()) m_hdr.flags |= HMR_FRM_HAS_CCP;
  if (sps->getUseExtendedPrecision())        fail("extended_precision_processing");
  if (sps->getScalingListFlag())
  {
    // the list TDecTop installs for this slice (TDecTop.cpp: SPS list, overridden by the PPS list, else the default one),
    // expanded per transform size exactly like TComTrQuant::xSetScalingListDec / processScalingListDec
    m_hdr.flags |= HMR_FRM_SCALING_LIST;
    TComScalingList* sl = slice->getScalingList();
    m_scaling.assign(HMR_SCALING_BYTES, 16);
    for (UInt sizeId = 0; sizeId < SCALING_LIST_SIZE_NUM; sizeId++)
    {
      const int N = (int)g_scalingListSizeX[sizeId];
      const int stored = N < MAX_MATRIX_SIZE_NUM ? N : MAX_MATRIX_SIZE_NUM, ratio = N / stored;
      for (UInt listId = 0; listId < SCALING_LIST_NUM; listId++)
      {
        const Int* coeff = sl->getScalingListAddress(sizeId, listId);
        uint8_t* dst = &m_scaling[HMR_SCALING_OFFSET(sizeId) + listId * N * N];
        for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) dst[j * N + i] = (uint8_t)coeff[stored * (j / ratio) + i / ratio];
        if (ratio > 1) dst[0] = (uint8_t)sl->getScalingListDC(sizeId, listId);
      }
    }
  }
  if (g_uiMaxCUWidth != g_uiMaxCUHeight)     fail("non-square CTU");
  // HM's per-partition arrays have one entry per (CTU size >> total depth) samples: 4 when the smallest transform block is 4x4,
  // 8 / 16 / 32 with log2_min_luma_transform_block_size 3 / 4 / 5.  The records are per 4 samples either way (availability bits,
  // BS map): every HM value then covers 1 << m_lgU units.  (Prediction and transform edges always lie on partition boundaries.)
  {
    const unsigned unit = g_uiMaxCUWidth >> g_uiMaxCUDepth;
    if (unit != 4 && unit != 8 && unit != 16 && unit != 32) fail("minimum partition size outside 4..32");
    m_lgU = 0;
    while ((4u << m_lgU) < unit && m_lgU < 3) m_lgU++;
  }
  // test hook for the error path of the reference ABI: refuse the (n+1)-th picture as if it used an unsupported tool
  {
    const char* refuse = getenv("HMDEC_B200_REFUSE_AFTER");          // read per picture: a test sets it for one decoder of a long-lived process
    if (refuse && m_nPic >= atoi(refuse)) fail("refused by HMDEC_B200_REFUSE_AFTER");
  }
  // 4:0:0: luma records only (HM walks getNumberValidComponents() components everywhere)

  m_tu.clear(); m_coef.clear(); m_intra.clear(); m_pu.clear(); m_puPrefix.clear(); m_puRefIdx.clear();
  m_puPrefix.push_back(0);
  m_range.assign(m_hdr.n_ctu, hmr_ctu_intra_range());
  memset(m_range.data(), 0, m_range.size() * sizeof(hmr_ctu_intra_range));
  m_ctu.assign(m_hdr.n_ctu, hmr_ctu());
  memset(m_ctu.data(), 0, m_ctu.size() * sizeof(hmr_ctu));
  m_bsStride = (m_hdr.width + 3) >> 2;
  m_qpStride = (m_hdr.width + 7) >> 3;
  m_bs.assign((size_t)m_bsStride * ((m_hdr.height + 3) >> 2), 0);
  m_qp.assign((size_t)m_qpStride * ((m_hdr.height + 7) >> 3), 0);
  m_cuFlags.assign(m_qp.size(), 0);
  m_curPic = pic;
  m_open = true;
}

// ---------------------------------------------------------------------------------------------
// A reference picture that never arrived: TDecTop::xCreateLostPicture (TDecTop.cpp:233-281) takes a fresh buffer, copies the samples
// of the closest picture of the DPB into it and marks it as a reference that is also output.  HM copies host planes; here the samples
// live in the engine's DPB, so the stand-in is sent as a picture of its own: one uni-predicted PU per CTU with a zero vector from
// the source's slot, no residual, no loop filters — integer-sample motion compensation is a bit-exact copy.
void HmEmitter::onLostPicture(TComPic* fill, TComPic* src, int poc)
{
  if (m_open) { fail("a lost reference picture is concealed in the middle of another picture"); return; }
  std::map<TComPic*, int>::iterator from = m_slots.find(src);
  if (from == m_slots.end()) { fail("a lost reference picture is concealed from a picture the engine never saw"); return; }
  TComSPS* sps = fill->getSlice(0)->getSPS();
  const int ctu = (int)g_uiMaxCUWidth;
  memset(&m_hdr, 0, sizeof(m_hdr));
  m_hdr.magic   = HMR_MAGIC;
  m_hdr.version = HMR_VERSION;
  m_hdr.width   = sps->getPicWidthInLumaSamples();
  m_hdr.height  = sps->getPicHeightInLumaSamples();
  m_hdr.poc     = poc;
  m_hdr.chroma_format    = (uint8_t)fill->getChromaFormat();
  m_hdr.bit_depth_luma   = (uint8_t)g_bitDepth[CHANNEL_TYPE_LUMA];
  m_hdr.bit_depth_chroma = (uint8_t)g_bitDepth[CHANNEL_TYPE_CHROMA];
  m_hdr.log2_ctu = (uint8_t)(g_aucConvertToBit[g_uiMaxCUWidth] + 2);
  m_hdr.out_slot = (uint8_t)slotOf(fill);
  m_hdr.slice_type = (uint8_t)P_SLICE;
  m_hdr.n_ctu = fill->getNumCUsInFrame();
  m_hdr.flags = HMR_FRM_IS_REFERENCE;
  if (m_unsupported) return;                                   // slotOf ran out of slots
  if (g_uiMaxCUWidth != g_uiMaxCUHeight || ctu > 64) { fail("non-square CTU"); return; }

  m_tu.clear(); m_coef.clear(); m_intra.clear(); m_pu.clear(); m_puPrefix.clear(); m_puRefIdx.clear();
  m_puPrefix.push_back(0);
  m_range.assign(m_hdr.n_ctu, hmr_ctu_intra_range());
  memset(m_range.data(), 0, m_range.size() * sizeof(hmr_ctu_intra_range));
  m_ctu.assign(m_hdr.n_ctu, hmr_ctu());
  memset(m_ctu.data(), 0, m_ctu.size() * sizeof(hmr_ctu));
  m_bsStride = (m_hdr.width + 3) >> 2;
  m_qpStride = (m_hdr.width + 7) >> 3;
  m_bs.assign((size_t)m_bsStride * ((m_hdr.height + 3) >> 2), 0);
  m_qp.assign((size_t)m_qpStride * ((m_hdr.height + 7) >> 3), 0);
  m_cuFlags.assign(m_qp.size(), 0);
  for (int y = 0; y < m_hdr.height; y += ctu)
    for (int x = 0; x < m_hdr.width; x += ctu)
    {
      hmr_pu p;
      memset(&p, 0, sizeof(p));
      p.x = (uint16_t)x; p.y = (uint16_t)y;
      p.w = (uint8_t)std::min(ctu, m_hdr.width - x); p.h = (uint8_t)std::min(ctu, m_hdr.height - y);
      p.lists = HMR_PU_L0;
      p.slots = (uint8_t)from->second;
      m_pu.push_back(p);
      m_puPrefix.push_back(m_puPrefix.back() + (uint32_t)(((p.w + 15) >> 4) * ((p.h + 15) >> 4)));
    }
  for (int k = 0; k <= 4; k++) m_hdr.tu_first[k] = 0;
  m_hdr.n_pu = (uint32_t)m_pu.size();
  m_hdr.n_mc_tiles = m_puPrefix.back();

  hmr_frame_desc d;
  memset(&d, 0, sizeof(d));
  d.hdr = &m_hdr;
  d.tu = m_tu.data(); d.coef = m_coef.data();
  d.intra = m_intra.data(); d.intra_range = m_range.data();
  d.pu = m_pu.data(); d.pu_tile_prefix = m_puPrefix.data();
  d.ctu = m_ctu.data();
  d.qp = m_qp.data();
  if (!m_sink->frameReady(d, fill))
  {
    m_failText = std::string("the reconstruction engine rejected the stand-in for a lost picture: ") + (m_sink->error() ? m_sink->error() : "unknown error");
    fail(m_failText.c_str());
  }
  m_nPic++;
}

// ---------------------------------------------------------------------------------------------
// one CTU: mirrors TDecCu::decompressCU -> xDecompressCU (TDecCu.cpp:142-145, 373-447)
void HmEmitter::onCtuParsed(TComDataCU* ctu)
{
  static const bool stats = getenv("HMDEC_B200_STATS") != NULL;     // two clock reads per CTU are 1 % of the parser thread: only when asked for
  const double t0 = stats ? nowSec() : 0.0;
  TComPic* pic = ctu->getPic();
  if (!m_open || pic != m_curPic) beginFrame(pic, ctu);
  for (int c = 0; c < 3; c++) m_intraTmp[c].clear();
  static const bool prefetch = getenv("HMDEC_B200_NO_PREFETCH") == NULL;
  // which CTU: the one the parser initialises and parses next.  (Two ahead — the first choice — is 2-3 % slower alone and with 8 or 16
  // threads: the lines arrive in time either way, but a whole CTU's parse then lies between the request and the use and pushes them out of L1.)
  static const int pfDist = getenv("HMDEC_B200_PF_DIST") ? atoi(getenv("HMDEC_B200_PF_DIST")) : 1;
  if (prefetch) hm_fast_prefetch_begin(m_prefetch, pic, ctu->getAddr() + pfDist);   // the parser is about to init + parse CTU addr+1
  walkCU(ctu, 0, 0);
  static const bool bsAtEnd = getenv("HMDEC_B200_BS_AT_END") != NULL;   // A/B switch: derive the deblocking side info per picture, like HM
  if (!bsAtEnd) deblockCtu(ctu);                                   // edge flags + boundary strengths while the CTU's arrays are hot
  m_prefetch.step(1 << 20);                                        // whatever the walk did not get to
  hmr_ctu_intra_range& r = m_range[ctu->getAddr()];
  for (int c = 0; c < 3; c++)
  {
    r.first[c] = (uint32_t)m_intra.size();
    r.count[c] = (uint32_t)m_intraTmp[c].size();
    m_intra.insert(m_intra.end(), m_intraTmp[c].begin(), m_intraTmp[c].end());
  }
  if (stats) m_tCtu += nowSec() - t0;
}

void HmEmitter::walkCU(TComDataCU* ctu, unsigned absPartIdx, unsigned depth)
{
  TComPic*   pic   = ctu->getPic();
  TComSlice* slice = pic->getSlice(pic->getCurrSliceIdx());
  const unsigned picW = slice->getSPS()->getPicWidthInLumaSamples();
  const unsigned picH = slice->getSPS()->getPicHeightInLumaSamples();
  unsigned lx = ctu->getCUPelX() + g_auiRasterToPelX[g_auiZscanToRaster[absPartIdx]];
  unsigned ty = ctu->getCUPelY() + g_auiRasterToPelY[g_auiZscanToRaster[absPartIdx]];
  unsigned rx = lx + (g_uiMaxCUWidth >> depth) - 1;
  unsigned by = ty + (g_uiMaxCUHeight >> depth) - 1;
  unsigned curNumParts = pic->getNumPartInCU() >> (depth << 1);
  const unsigned scu = ctu->getSCUAddr();
  bool startInCU = scu + absPartIdx + curNumParts > slice->getSliceSegmentCurStartCUAddr() && scu + absPartIdx < slice->getSliceSegmentCurStartCUAddr();
  bool boundary = startInCU || rx >= picW || by >= picH;

  if ((depth < ctu->getDepth(absPartIdx) && depth < g_uiMaxCUDepth - g_uiAddCUDepth) || boundary)
  {
    unsigned qNumParts = ctu->getTotalNumPart() >> ((depth + 1) << 1);
    unsigned idx = absPartIdx;
    for (unsigned q = 0; q < 4; q++, idx += qNumParts)
    {
      unsigned qx = ctu->getCUPelX() + g_auiRasterToPelX[g_auiZscanToRaster[idx]];
      unsigned qy = ctu->getCUPelY() + g_auiRasterToPelY[g_auiZscanToRaster[idx]];
      bool inSlice = (scu + idx + qNumParts > slice->getSliceSegmentCurStartCUAddr()) && (scu + idx < slice->getSliceSegmentCurEndCUAddr());
      if (inSlice && qx < picW && qy < picH) walkCU(ctu, idx, depth + 1);
    }
    return;
  }

  m_prefetch.step(8);
  const int cuSize = g_uiMaxCUWidth >> depth;
  switch (ctu->getPredictionMode(absPartIdx))
  {
    case MODE_INTER: emitInterCU(ctu, absPartIdx, depth, lx, ty, cuSize); break;
    case MODE_INTRA: emitIntraCU(ctu, absPartIdx, depth, lx, ty); break;
    default: fail("CU with no prediction mode"); break;
  }
}

// ---------------------------------------------------------------------------------------------
// residual record for one square TU of one component: TComTrQuant::invTransformNxN (TComTrQuant.cpp:1423-1548)
uint32_t HmEmitter::emitResidualTU(CuCtx& c, int compIdx, void* pTu, bool intra, bool coded, int alpha)
{
  TComTU& rTu = *(TComTU*)pTu;
  const ComponentID compID = ComponentID(compIdx);
  TComDataCU* ctu = c.ctu;
  TComPic* pic = ctu->getPic();
  const TComRectangle& rect = rTu.getRect(compID);
  const unsigned absPartIdxTU = rTu.GetAbsPartIdxTU();
  const int csx = pic->getComponentScaleX(compID), csy = pic->getComponentScaleY(compID);
  const int N = rect.width;
  assert(rect.width == rect.height);

  hmr_tu t;
  memset(&t, 0, sizeof(t));
  t.x = (uint16_t)((c.cuX >> csx) + rect.x0);
  t.y = (uint16_t)((c.cuY >> csy) + rect.y0);
  t.comp = (uint8_t)compIdx;
  t.log2_size = (uint8_t)(g_aucConvertToBit[N] + 2);
  t.ccp_alpha = (int8_t)alpha;
  t.luma_off = HMR_NO_OFFSET;
  if (coded) t.flags |= HMR_TU_CODED;
  if (intra) t.flags |= HMR_TU_INTRA;

  const bool bypass = ctu->getCUTransquantBypass(absPartIdxTU);
  const bool tskip  = ctu->getTransformSkip(absPartIdxTU, compID) != 0;
  if (bypass) t.flags |= HMR_TU_BYPASS;
  else if (tskip) t.flags |= HMR_TU_TSKIP;
  if ((bypass || tskip) && rTu.isNonTransformedResidualRotated(compID)) t.flags |= HMR_TU_ROTATE;
  if (!bypass && !tskip && N == 4 && rTu.useDST(compID)) t.flags |= HMR_TU_DST;

  // invRdpcmNxN (TComTrQuant.cpp:1737-1792)
  if (ctu->isRDPCMEnabled(absPartIdxTU) && (tskip || bypass))
  {
    int mode = RDPCM_OFF;
    if (ctu->isIntra(absPartIdxTU))
    {
      const ChromaFormat chFmt = pic->getChromaFormat();
      const ChannelType chType = toChannelType(compID);
      const UInt chPredMode  = ctu->getIntraDir(chType, absPartIdxTU);
      const UInt chCodedMode = (chPredMode == DM_CHROMA_IDX && isChroma(compID)) ? ctu->getIntraDir(CHANNEL_TYPE_LUMA, getChromasCorrespondingPULumaIdx(absPartIdxTU, chFmt)) : chPredMode;
      const UInt chFinalMode = ((chFmt == CHROMA_422) && isChroma(compID)) ? g_chroma422IntraAngleMappingTable[chCodedMode] : chCodedMode;
      if (chFinalMode == VER_IDX) mode = RDPCM_VER; else if (chFinalMode == HOR_IDX) mode = RDPCM_HOR;
    }
    else mode = ctu->getExplicitRdpcmMode(compID, absPartIdxTU);
    if (mode == RDPCM_HOR) t.flags |= HMR_TU_RDPCM_H;
    if (mode == RDPCM_VER) t.flags |= HMR_TU_RDPCM_V;
  }

  // QpParam::QpParam(const TComDataCU&, ComponentID) evaluated at this CU (TComTrQuant.cpp:99-119)
  {
    TComSlice* slice = ctu->getSlice();
    Int chromaQpOffset = 0;
    if (isChroma(compID))
    {
      chromaQpOffset += slice->getPPS()->getQpOffset(compID);
      chromaQpOffset += slice->getSliceChromaQpDelta(compID);
      chromaQpOffset += slice->getPPS()->getChromaQpAdjTableAt(ctu->getChromaQpAdj(c.absPartIdx)).u.offset[Int(compID) - 1];
    }
    QpParam qp(ctu->getQP(c.absPartIdx), toChannelType(compID), slice->getSPS()->getQpBDOffset(toChannelType(compID)), chromaQpOffset, pic->getChromaFormat());
    t.qp = (uint8_t)qp.Qp;
  }

  // coefficient levels: int32 raster N x N at getCoeff()+offset (TDecSbac.cpp:1260); the dequantiser clips its
  // input to 16 bits for every legal QP/bit-depth (TComTrQuant.cpp:1284-1286), so int16 transport is exact
  t.coef_off = (uint32_t)m_coef.size();
  int16_t* dst = m_coef.grow((size_t)N * N);
  if (coded)
  {
    const TCoeff* src = ctu->getCoeff(compID) + rTu.getCoefficientOffset(compID);
    for (int i = 0; i < N * N; i++)
    {
      TCoeff v = src[i];
      dst[i] = (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
    }
  }
  else memset(dst, 0, sizeof(int16_t) * N * N);

  if (compIdx == 0) c.lumaOff[absPartIdxTU - c.absPartIdx] = t.coef_off;
  else if (alpha != 0) t.luma_off = c.lumaOff[absPartIdxTU - c.absPartIdx];

  m_tu.push_back(t);
  return t.coef_off;
}

// ---------------------------------------------------------------------------------------------
// inter CU: TDecCu::xReconInter (TDecCu.cpp:449-480) = motionCompensation + xDecodeInterTexture
void HmEmitter::emitInterCU(TComDataCU* ctu, unsigned absPartIdx, unsigned depth, int cuX, int cuY, int cuSize)
{
  TComSlice* slice = ctu->getSlice();
  TComPic*   pic   = ctu->getPic();
  // explicit weighted prediction (TComSlice::applyWP): table per (list, refIdx, component), once per picture
  const bool applyWP = (slice->getPPS()->getUseWP() && slice->getSliceType() == P_SLICE) || (slice->getPPS()->getWPBiPred() && slice->getSliceType() == B_SLICE);
  if (applyWP && !(m_hdr.flags & HMR_FRM_WEIGHTED_PRED))
  {
    if (!m_pu.empty()) fail("slices with and without weighted prediction in one picture");
    m_hdr.flags |= HMR_FRM_WEIGHTED_PRED;
    m_wp.assign(HMR_WP_ENTRIES, hmr_wp());
    memset(m_wp.data(), 0, sizeof(hmr_wp) * m_wp.size());
    const bool highPrec = slice->getSPS()->getUseHighPrecisionPredictionWeighting();
    for (int l = 0; l < 2; l++)
      for (int r = 0; r < slice->getNumRefIdx(RefPicList(l)) && r < 16; r++)
      {
        WPScalingParam* wp = NULL;
        slice->getWpScaling(RefPicList(l), r, wp);
        for (UInt c = 0; c < pic->getNumberValidComponents(); c++)
        {
          hmr_wp& o = m_wp[(l * 16 + r) * 3 + c];
          const int bd = g_bitDepth[toChannelType(ComponentID(c))];
          o.weight = (int16_t)wp[c].iWeight;
          o.offset = (int16_t)(wp[c].iOffset * (highPrec ? 1 : (1 << (bd - 8))));
          o.log2_denom = (uint8_t)wp[c].uiLog2WeightDenom;
        }
      }
  }
  else if (!applyWP && (m_hdr.flags & HMR_FRM_WEIGHTED_PRED)) fail("slices with and without weighted prediction in one picture");

  // PU geometry: TComDataCU::getPartIndexAndSize (TComDataCU.cpp:2178-2230)
  int n = 1, px[4] = {0, 0, 0, 0}, py[4] = {0, 0, 0, 0}, pw[4], ph[4];
  const int S = cuSize, H = S >> 1, Q = S >> 2;
  switch (ctu->getPartitionSize(absPartIdx))
  {
    case SIZE_2NxN:  n = 2; pw[0] = pw[1] = S; ph[0] = ph[1] = H; py[1] = H; break;
    case SIZE_Nx2N:  n = 2; pw[0] = pw[1] = H; ph[0] = ph[1] = S; px[1] = H; break;
    case SIZE_NxN:   n = 4; for (int i = 0; i < 4; i++) { pw[i] = ph[i] = H; px[i] = (i & 1) * H; py[i] = (i >> 1) * H; } break;
    case SIZE_2NxnU: n = 2; pw[0] = pw[1] = S; ph[0] = Q; ph[1] = S - Q; py[1] = Q; break;
    case SIZE_2NxnD: n = 2; pw[0] = pw[1] = S; ph[0] = S - Q; ph[1] = Q; py[1] = S - Q; break;
    case SIZE_nLx2N: n = 2; ph[0] = ph[1] = S; pw[0] = Q; pw[1] = S - Q; px[1] = Q; break;
    case SIZE_nRx2N: n = 2; ph[0] = ph[1] = S; pw[0] = S - Q; pw[1] = Q; px[1] = S - Q; break;
    default:         n = 1; pw[0] = ph[0] = S; break;
  }
  const unsigned partStride = pic->getNumPartInWidth();
  const unsigned cuRaster = g_auiZscanToRaster[absPartIdx];
  const int picW = slice->getSPS()->getPicWidthInLumaSamples(), picH = slice->getSPS()->getPicHeightInLumaSamples();
  for (int i = 0; i < n; i++)
  {
    const unsigned partAddr = g_auiRasterToZscan[cuRaster + (py[i] >> (2 + m_lgU)) * partStride + (px[i] >> (2 + m_lgU))];
    hmr_pu p;
    memset(&p, 0, sizeof(p));
    p.x = (uint16_t)(cuX + px[i]); p.y = (uint16_t)(cuY + py[i]);
    p.w = (uint8_t)pw[i]; p.h = (uint8_t)ph[i];
    int refIdx[2]; TComMv mv[2]; TComPic* ref[2] = {NULL, NULL};
    for (int l = 0; l < 2; l++)
    {
      refIdx[l] = ctu->getCUMvField(RefPicList(l))->getRefIdx(partAddr);
      mv[l]     = ctu->getCUMvField(RefPicList(l))->getMv(partAddr);
      if (refIdx[l] >= 0) ref[l] = slice->getRefPic(RefPicList(l), refIdx[l]);
    }
    bool use0 = refIdx[0] >= 0, use1 = refIdx[1] >= 0;
    // xCheckIdenticalMotion (TComPrediction.cpp:497-512)
    if (use0 && use1 && slice->isInterB() && !slice->getPPS()->getWPBiPred() &&
        ref[0]->getPOC() == ref[1]->getPOC() && mv[0] == mv[1]) use1 = false;
    if (!use0 && !use1) { fail("inter PU without reference"); continue; }
    int slot[2] = {0, 0};
    for (int l = 0; l < 2; l++)
    {
      if (!(l ? use1 : use0)) continue;
      // TComDataCU::clipMv with the CU's own position (TComDataCU.cpp:3102-3114; pcCU is the sub-CU copy there)
      int hmax = (picW + 8 - cuX - 1) << 2, hmin = (-(int)g_uiMaxCUWidth - 8 - cuX + 1) << 2;
      int vmax = (picH + 8 - cuY - 1) << 2, vmin = (-(int)g_uiMaxCUHeight - 8 - cuY + 1) << 2;
      p.mv[l][0] = (int16_t)std::min(hmax, std::max(hmin, (int)mv[l].getHor()));
      p.mv[l][1] = (int16_t)std::min(vmax, std::max(vmin, (int)mv[l].getVer()));
      slot[l] = slotOf(ref[l]);
      p.lists |= (l ? HMR_PU_L1 : HMR_PU_L0);
    }
    p.slots = (uint8_t)(slot[0] | (slot[1] << 4));
    m_pu.push_back(p);
    m_puRefIdx.push_back((uint8_t)((use0 ? (refIdx[0] & 15) : 0) | ((use1 ? (refIdx[1] & 15) : 0) << 4)));
    m_puPrefix.push_back(m_puPrefix.back() + ((pw[i] + 15) >> 4) * ((ph[i] + 15) >> 4));
  }

  // residual: xDecodeInterTexture -> invRecurTransformNxN per component (TDecCu.cpp:743-758, TComTrQuant.cpp:1550-1615)
  CuCtx c;
  c.ctu = ctu; c.absPartIdx = absPartIdx; c.depth = depth; c.cuX = cuX; c.cuY = cuY;
  for (int i = 0; i < 256; i++) c.lumaOff[i] = HMR_NO_OFFSET;
  TComTURecurse tuRecur(ctu, absPartIdx, depth);
  for (UInt ch = 0; ch < pic->getNumberValidComponents(); ch++) interResidual(c, ch, &tuRecur);
}

void HmEmitter::interResidual(CuCtx& c, int compIdx, void* pTu)
{
  TComTU& rTu = *(TComTU*)pTu;
  const ComponentID compID = ComponentID(compIdx);
  if (!rTu.ProcessComponentSection(compID)) return;
  TComDataCU* ctu = c.ctu;
  const UInt absPartIdxTU = rTu.GetAbsPartIdxTU();
  const UInt trMode = rTu.GetTransformDepthRel();
  const bool ccpEnabled = ctu->getSlice()->getPPS()->getUseCrossComponentPrediction();
  if (ctu->getCbf(absPartIdxTU, compID, trMode) == 0 && (isLuma(compID) || !ccpEnabled)) return;

  if (trMode == ctu->getTransformIdx(absPartIdxTU))
  {
    const bool coded = ctu->getCbf(absPartIdxTU, compID, trMode) != 0;
    int alpha = 0;
    if (isChroma(compID) && ctu->getCrossComponentPredictionAlpha(absPartIdxTU, compID) != 0 &&
        ctu->getCbf(absPartIdxTU, COMPONENT_Y, trMode) != 0)
      alpha = ctu->getCrossComponentPredictionAlpha(absPartIdxTU, compID);
    if (!coded && alpha == 0) return;
    const TComRectangle& rect = rTu.getRect(compID);
    if (rect.width != rect.height)
    {
      // 4:2:2 chroma: two square halves, both run through the transform (TComTrQuant.cpp:1437-1464)
      // Each half has its own cbf one level down (TDecSbac::parseQtCbf); HM runs both through invTransformNxN, the one
      // without coded levels on an all-zero block — which only exists if the storage was zero-filled (hm_fast.cpp), so say so.
      TComTURecurse sub(rTu, false, TComTU::VERTICAL_SPLIT, true, compID);
      do
      {
        const bool subCoded = coded && ctu->getCbf(sub.GetAbsPartIdxTU(compID), compID, trMode + 1) != 0;
        emitResidualTU(c, compIdx, &sub, false, subCoded, alpha);
      } while (sub.nextSection(rTu));
    }
    else emitResidualTU(c, compIdx, &rTu, false, coded, alpha);
  }
  else
  {
    TComTURecurse child(rTu, false);
    do { interResidual(c, compIdx, &child); } while (child.nextSection(rTu));
  }
}

// ---------------------------------------------------------------------------------------------
// intra CU: TDecCu::xReconIntraQT / xIntraRecQT / xIntraRecBlk (TDecCu.cpp:662-732, 483-659)
void HmEmitter::emitIntraCU(TComDataCU* ctu, unsigned absPartIdx, unsigned depth, int cuX, int cuY)
{
  if (ctu->getIPCMFlag(absPartIdx)) { emitPcmCU(ctu, absPartIdx, depth, cuX, cuY); return; }
  CuCtx c;
  c.ctu = ctu; c.absPartIdx = absPartIdx; c.depth = depth; c.cuX = cuX; c.cuY = cuY;
  for (int i = 0; i < 256; i++) c.lumaOff[i] = HMR_NO_OFFSET;
  const ChromaFormat chFmt = ctu->getPic()->getChromaFormat();
  const UInt numChType = chFmt != CHROMA_400 ? 2 : 1;
  for (UInt chType = CHANNEL_TYPE_LUMA; chType < numChType; chType++)
  {
    const bool NxNPUHas4Parts = ::isChroma(ChannelType(chType)) ? enable4ChromaPUsInIntraNxNCU(chFmt) : true;
    const UInt initTrDepth = (ctu->getPartitionSize(absPartIdx) != SIZE_2Nx2N && NxNPUHas4Parts) ? 1 : 0;
    TComTURecurse tuCU(ctu, absPartIdx);
    TComTURecurse tuPU(tuCU, false, (initTrDepth == 0) ? TComTU::DONT_SPLIT : TComTU::QUAD_SPLIT);
    do { intraQT(c, chType, &tuPU); } while (tuPU.nextSection(tuCU));
  }
}

// I_PCM CU: TDecCu::xReconPCM / xDecodePCMTexture (TDecCu.cpp:771-842).  Per component one (4:2:2 chroma: two) square
// block(s): a bypass residual record whose "levels" are the PCM samples at the internal bit depth, and an intra record
// with the PCM pseudo mode (prediction 0), so that the samples land in decode order like any other intra block.
void HmEmitter::emitPcmCU(TComDataCU* ctu, unsigned absPartIdx, unsigned depth, int cuX, int cuY)
{
  TComPic* pic = ctu->getPic();
  const int cuSize = g_uiMaxCUWidth >> depth;
  const UInt minArea = (g_uiMaxCUWidth >> g_uiMaxCUDepth) * (g_uiMaxCUHeight >> g_uiMaxCUDepth);
  for (UInt ch = 0; ch < pic->getNumberValidComponents(); ch++)
  {
    const ComponentID compID = ComponentID(ch);
    const int csx = pic->getComponentScaleX(compID), csy = pic->getComponentScaleY(compID);
    const int w = cuSize >> csx, h = cuSize >> csy;
    const Pel* pcm = ctu->getPCMSample(compID) + ((minArea * absPartIdx) >> (csx + csy));
    const int shift = g_bitDepth[toChannelType(compID)] - ctu->getSlice()->getSPS()->getPCMBitDepth(toChannelType(compID));
    for (int part = 0; part < h / w; part++)                 // 4:2:2 chroma: w x 2w -> two squares
    {
      int lg = 0; while ((1 << lg) < w) lg++;
      hmr_tu t; memset(&t, 0, sizeof(t));
      t.x = (uint16_t)(cuX >> csx); t.y = (uint16_t)((cuY >> csy) + part * w);
      t.comp = (uint8_t)ch; t.log2_size = (uint8_t)lg;
      t.flags = HMR_TU_CODED | HMR_TU_INTRA | HMR_TU_BYPASS;
      t.luma_off = HMR_NO_OFFSET;
      t.coef_off = (uint32_t)m_coef.size();
      int16_t* dst = m_coef.grow((size_t)w * w);
      const Pel* src = pcm + (size_t)part * w * w;
      for (int i = 0; i < w * w; i++) dst[i] = (int16_t)(src[i] << shift);
      m_tu.push_back(t);
      hmr_intra r; memset(&r, 0, sizeof(r));
      r.x = t.x; r.y = t.y; r.comp = t.comp; r.log2_size = t.log2_size;
      r.mode = HMR_INTRA_MODE_PCM;
      r.resid_off = t.coef_off;
      m_intraTmp[ch].push_back(r);
    }
  }
}

void HmEmitter::intraQT(CuCtx& c, int chType, void* pTu)
{
  TComTU& rTu = *(TComTU*)pTu;
  TComDataCU* ctu = c.ctu;
  const UInt trDepth = rTu.GetTransformDepthRel();
  const UInt absPartIdx = rTu.GetAbsPartIdxTU();
  if (ctu->getTransformIdx(absPartIdx) == trDepth)
  {
    if (chType == CHANNEL_TYPE_LUMA) intraBlk(c, COMPONENT_Y, &rTu);
    else
    {
      const UInt numValidComp = getNumberValidComponents(rTu.GetChromaFormat());
      for (UInt comp = COMPONENT_Cb; comp < numValidComp; comp++) intraBlk(c, comp, &rTu);
    }
  }
  else
  {
    TComTURecurse child(rTu, false);
    do { intraQT(c, chType, &child); } while (child.nextSection(rTu));
  }
}

void HmEmitter::intraBlk(CuCtx& c, int compIdx, void* pTu)
{
  TComTU& rTu = *(TComTU*)pTu;
  const ComponentID compID = ComponentID(compIdx);
  if (!rTu.ProcessComponentSection(compID)) return;
  TComDataCU* ctu = c.ctu;
  TComPic* pic = ctu->getPic();
  const bool bIsLuma = isLuma(compID);
  const UInt absPartIdx = rTu.GetAbsPartIdxTU();
  const TComRectangle& rect = rTu.getRect(compID);
  const UInt W = rect.width, Hh = rect.height;
  const ChromaFormat chFmt = rTu.GetChromaFormat();
  if (W != Hh)
  {
    TComTURecurse sub(rTu, false, TComTU::VERTICAL_SPLIT, true, compID);
    m_in422SubTu = true;
    do { intraBlk(c, compIdx, &sub); } while (sub.nextSection(rTu));
    m_in422SubTu = false;
    return;
  }
  const UInt chPredMode  = ctu->getIntraDir(toChannelType(compID), absPartIdx);
  const UInt chCodedMode = (chPredMode == DM_CHROMA_IDX && !bIsLuma) ? ctu->getIntraDir(CHANNEL_TYPE_LUMA, getChromasCorrespondingPULumaIdx(absPartIdx, chFmt)) : chPredMode;
  const UInt chFinalMode = ((chFmt == CHROMA_422) && !bIsLuma) ? g_chroma422IntraAngleMappingTable[chCodedMode] : chCodedMode;

  hmr_intra r;
  memset(&r, 0, sizeof(r));
  const int csx = pic->getComponentScaleX(compID), csy = pic->getComponentScaleY(compID);
  r.x = (uint16_t)((c.cuX >> csx) + rect.x0);
  r.y = (uint16_t)((c.cuY >> csy) + rect.y0);
  r.comp = (uint8_t)compIdx;
  r.log2_size = (uint8_t)(g_aucConvertToBit[W] + 2);
  r.mode = (uint8_t)chFinalMode;
  if (TComPrediction::filteringIntraReferenceSamples(compID, chFinalMode, W, Hh, chFmt, ctu->getSlice()->getSPS()->getDisableIntraReferenceSmoothing()))
    r.flags |= HMR_INTRA_FILTER_REFS;
  if (bIsLuma) r.flags |= HMR_INTRA_LUMA_RULES;
  if (ctu->isRDPCMEnabled(absPartIdx) && ctu->getCUTransquantBypass(absPartIdx)) r.flags |= HMR_INTRA_NO_EDGE_FLT;

  // neighbour availability: TComPrediction::initAdiPatternChType (TComPattern.cpp:107-145)
  {
    const Int baseUnit = g_uiMaxCUWidth >> g_uiMaxCUDepth;
    const Int unitW = baseUnit >> csx, unitH = baseUnit >> csy;
    const Int wUnits = W / unitW, hUnits = Hh / unitH;
    const Int leftUnits = hUnits << 1;
    const Int partStride = pic->getNumPartInWidth();
    const UInt idxLT = ctu->getZorderIdxInCU() + absPartIdx;
    const UInt idxRT = g_auiRasterToZscan[g_auiZscanToRaster[idxLT] + wUnits - 1];
    const UInt idxLB = g_auiRasterToZscan[g_auiZscanToRaster[idxLT] + (hUnits - 1) * partStride];
    Bool flags[4 * MAX_NUM_SPU_W + 1];
    memset(flags, 0, sizeof(flags));
    // HM asks getPUAbove / getPULeft / getPUAboveRightAdi / getPUBelowLeftAdi once per 4-sample unit (z-scan conversions, slice,
    // tile and CTU-order tests every time: 4 % of the parser thread).  In a picture of ONE slice and ONE tile without constrained
    // intra prediction all of those rules collapse to geometry: a unit is available iff it lies inside the picture and was
    // decoded before this block — a CTU to the left or in the row above (the above-right one included), or, inside this CTU,
    // a unit that precedes the block in z-order (units straight above / left of it always do).  Everything else: HM's functions.
    static const bool hmAvail = getenv("HMDEC_B200_HM_AVAIL") != NULL;      // A/B switch
    TComSlice* sl = ctu->getSlice();
    if (!hmAvail && !sl->getPPS()->getConstrainedIntraPred() && sl->getSliceCurStartCUAddr() == 0 &&
        sl->getPPS()->getNumTileColumnsMinus1() == 0 && sl->getPPS()->getTileNumRowsMinus1() == 0)
    {
      const int ctuSize = (int)g_uiMaxCUWidth, unitsPerCtu = ctuSize / baseUnit;
      const int picW = (int)sl->getSPS()->getPicWidthInLumaSamples(), picH = (int)sl->getSPS()->getPicHeightInLumaSamples();
      const int ctuX = (int)ctu->getCUPelX(), ctuY = (int)ctu->getCUPelY();
      const int rasterLT = (int)g_auiZscanToRaster[idxLT];
      const int ux0 = rasterLT % partStride, uy0 = rasterLT / partStride;   // the block's first unit inside the CTU
      // is the unit at (ux, uy) — CTU-relative unit coordinates, possibly outside [0, unitsPerCtu) — available to this block?
      auto avail = [&](int ux, int uy) -> Bool
      {
        const int x = ctuX + ux * baseUnit, y = ctuY + uy * baseUnit;
        if (x < 0 || y < 0 || x >= picW || y >= picH) return false;
        if (uy < 0) return ux < 2 * unitsPerCtu;                            // the CTU row above: left, above and above-right CTUs are done
        if (uy >= unitsPerCtu) return false;                                // the CTU row below: not yet
        if (ux < 0) return true;                                            // the CTU to the left
        if (ux >= unitsPerCtu) return false;                                // the CTU to the right
        return g_auiRasterToZscan[uy * partStride + ux] < idxLT;            // inside this CTU: decoded before the block?
      };
      flags[leftUnits] = avail(ux0 - 1, uy0 - 1);
      for (int i = 0; i < wUnits; i++)
      {
        flags[leftUnits + 1 + i] = avail(ux0 + i, uy0 - 1);
        flags[leftUnits + 1 + wUnits + i] = avail(ux0 + wUnits + i, uy0 - 1);
      }
      for (int i = 0; i < hUnits; i++)
      {
        flags[leftUnits - 1 - i] = avail(ux0 - 1, uy0 + i);
        flags[leftUnits - 1 - hUnits - i] = avail(ux0 - 1, uy0 + hUnits + i);
      }
      static const bool check = getenv("HMDEC_B200_CHECK_AVAIL") != NULL;   // development aid: compare with HM's functions
      if (check)
      {
        Bool ref[4 * MAX_NUM_SPU_W + 1];
        memset(ref, 0, sizeof(ref));
        ref[leftUnits] = isAboveLeftAvailable(ctu, idxLT);
        isAboveAvailable     (ctu, idxLT, idxRT, ref + leftUnits + 1);
        isAboveRightAvailable(ctu, idxLT, idxRT, ref + leftUnits + 1 + wUnits);
        isLeftAvailable      (ctu, idxLT, idxLB, ref + leftUnits - 1);
        isBelowLeftAvailable (ctu, idxLT, idxLB, ref + leftUnits - 1 - hUnits);
        for (int i = 0; i < 2 * leftUnits + 1 - 0 && i < leftUnits + 1 + 2 * wUnits; i++)
          if ((flags[i] != 0) != (ref[i] != 0))
          {
            fprintf(stderr, "intra availability mismatch: POC %d CTU %u block (%d,%d) %ux%u comp %d, flag %d (corner at %d): direct %d, HM %d\n",
                    (int)pic->getPOC(), ctu->getAddr(), (int)r.x, (int)r.y, W, Hh, compIdx, i, leftUnits, (int)flags[i], (int)ref[i]);
            abort();
          }
      }
    }
    else
    {
      flags[leftUnits] = isAboveLeftAvailable(ctu, idxLT);
      isAboveAvailable     (ctu, idxLT, idxRT, flags + leftUnits + 1);
      isAboveRightAvailable(ctu, idxLT, idxRT, flags + leftUnits + 1 + wUnits);
      isLeftAvailable      (ctu, idxLT, idxLB, flags + leftUnits - 1);
      isBelowLeftAvailable (ctu, idxLT, idxLB, flags + leftUnits - 1 - hUnits);
    }
    if (flags[leftUnits]) r.flags |= HMR_INTRA_AVAIL_CORNER;
    // record bits are per 4 luma samples: one HM unit is 1 << m_lgU of them
    const unsigned unitBits = (1u << (1 << m_lgU)) - 1;
    for (int i = 0; i < wUnits; i++)
    {
      if (flags[leftUnits + 1 + i])          r.avail_above       |= (uint8_t)(unitBits << (i << m_lgU));
      if (flags[leftUnits + 1 + wUnits + i]) r.avail_above_right |= (uint8_t)(unitBits << (i << m_lgU));
    }
    for (int i = 0; i < hUnits; i++)
    {
      if (flags[leftUnits - 1 - i])          r.avail_left        |= (uint8_t)(unitBits << (i << m_lgU));
      if (flags[leftUnits - 1 - hUnits - i]) r.avail_below_left  |= (uint8_t)(unitBits << (i << m_lgU));
    }
  }

  // residual (TDecCu.cpp:560-586) + cross-component prediction (TDecCu.cpp:600-625)
  // HM tests the cbf at the TU's own level (TDecCu.cpp:560); for the two halves of a 4:2:2 chroma TU that is the COMBINED
  // flag, and the half without levels is inverse-transformed from an all-zero block — which only exists if the storage
  // was zero-filled (hm_fast.cpp).  Each half has its own flag one level down (TDecSbac::parseQtCbf): use it.
  bool coded = ctu->getCbf(absPartIdx, compID, rTu.GetTransformDepthRel()) != 0;
  if (coded && m_in422SubTu) coded = ctu->getCbf(rTu.GetAbsPartIdxTU(compID), compID, rTu.GetTransformDepthRel() + 1) != 0;
  const int alpha = isChroma(compID) ? ctu->getCrossComponentPredictionAlpha(absPartIdx, compID) : 0;
  const bool keepLuma = bIsLuma && (m_hdr.flags & HMR_FRM_HAS_CCP); // chroma CCP may read a zero luma residual
  r.resid_off = HMR_NO_OFFSET;
  if (coded || alpha != 0 || keepLuma)
  {
    bool lumaHasResid = c.lumaOff[absPartIdx - c.absPartIdx] != HMR_NO_OFFSET;
    if (coded || (alpha != 0 && lumaHasResid))
      r.resid_off = emitResidualTU(c, compIdx, &rTu, true, coded, (alpha != 0 && lumaHasResid) ? alpha : 0);
  }
  m_intraTmp[compIdx].push_back(r);
}

// ---------------------------------------------------------------------------------------------
// deblocking side info: TComLoopFilter::loopFilterPic / xDeblockCU (TComLoopFilter.cpp:130-234) without the filtering
// Boundary strength of one 4x4 edge unit: the rule of TComLoopFilter::xGetBoundaryStrengthSingle (TComLoopFilter.cpp:411-537)
// evaluated on the two CTUs' arrays directly.  HM's own routine finds the P side through getPULeft / getPUAbove for every unit;
// here it is the raster neighbour inside the CTU, or the facing unit of the left / above CTU (TComDataCU.cpp:503,508,1243).  The
// slice / tile / picture-border rules of those getters have already decided whether the edge exists at all
// (xSetLoopfilterParam, TComLoopFilter.cpp:352-409): a unit that reaches this point has its P side.  tuEdge = HM's pre-set TU-edge marker.
static inline bool mvFar(const TComMv& a, const TComMv& b) { return abs(a.getHor() - b.getHor()) >= 4 || abs(a.getVer() - b.getVer()) >= 4; }

static unsigned bsOfUnit(TComDataCU* cuP, unsigned partP, TComDataCU* cuQ, unsigned partQ, bool tuEdge)
{
  if (cuP->isIntra(partP) || cuQ->isIntra(partQ)) return 2;
  if (tuEdge && (cuQ->getCbf(partQ, COMPONENT_Y, cuQ->getTransformIdx(partQ)) != 0 || cuP->getCbf(partP, COMPONENT_Y, cuP->getTransformIdx(partP)) != 0)) return 1;
  TComSlice* sliceQ = cuQ->getSlice();
  TComSlice* sliceP = cuP->getSlice();
  TComCUMvField* fP0 = cuP->getCUMvField(REF_PIC_LIST_0);
  TComCUMvField* fQ0 = cuQ->getCUMvField(REF_PIC_LIST_0);
  // identical motion data under the same reference lists: every branch of the rule below ends in 0 (the common case by far)
  if (sliceP == sliceQ && fP0->getRefIdx(partP) == fQ0->getRefIdx(partQ) && fP0->getMv(partP) == fQ0->getMv(partQ))
  {
    TComCUMvField* gP1 = cuP->getCUMvField(REF_PIC_LIST_1);
    TComCUMvField* gQ1 = cuQ->getCUMvField(REF_PIC_LIST_1);
    if (gP1->getRefIdx(partP) == gQ1->getRefIdx(partQ) && gP1->getMv(partP) == gQ1->getMv(partQ)) return 0;
  }
  int r = fP0->getRefIdx(partP);
  const TComPic* refP0 = r < 0 ? NULL : sliceP->getRefPic(REF_PIC_LIST_0, r);
  r = fQ0->getRefIdx(partQ);
  const TComPic* refQ0 = r < 0 ? NULL : sliceQ->getRefPic(REF_PIC_LIST_0, r);
  const TComMv zero;
  const TComMv& mvP0 = refP0 ? fP0->getMv(partP) : zero;
  const TComMv& mvQ0 = refQ0 ? fQ0->getMv(partQ) : zero;
  if (!sliceQ->isInterB() && !sliceP->isInterB()) return (refP0 != refQ0 || mvFar(mvQ0, mvP0)) ? 1 : 0;
  TComCUMvField* fP1 = cuP->getCUMvField(REF_PIC_LIST_1);
  TComCUMvField* fQ1 = cuQ->getCUMvField(REF_PIC_LIST_1);
  r = fP1->getRefIdx(partP);
  const TComPic* refP1 = r < 0 ? NULL : sliceP->getRefPic(REF_PIC_LIST_1, r);
  r = fQ1->getRefIdx(partQ);
  const TComPic* refQ1 = r < 0 ? NULL : sliceQ->getRefPic(REF_PIC_LIST_1, r);
  const TComMv& mvP1 = refP1 ? fP1->getMv(partP) : zero;
  const TComMv& mvQ1 = refQ1 ? fQ1->getMv(partQ) : zero;
  if (!((refP0 == refQ0 && refP1 == refQ1) || (refP0 == refQ1 && refP1 == refQ0))) return 1;     // different reference pictures
  if (refP0 != refP1)                                                                              // two distinct pictures: match the lists up
    return (refP0 == refQ0) ? ((mvFar(mvQ0, mvP0) || mvFar(mvQ1, mvP1)) ? 1 : 0) : ((mvFar(mvQ1, mvP0) || mvFar(mvQ0, mvP1)) ? 1 : 0);
  return ((mvFar(mvQ0, mvP0) || mvFar(mvQ1, mvP1)) && (mvFar(mvQ1, mvP0) || mvFar(mvQ0, mvP1))) ? 1 : 0;   // both lists use the same picture
}

void HmEmitter::bsWalk(TComDataCU* ctu, unsigned absZorderIdx, unsigned depth, TComLoopFilter* lf)
{
  if (ctu->getPic() == 0 || ctu->getPartitionSize(absZorderIdx) == NUMBER_OF_PART_SIZES) return;
  TComPic* pic = ctu->getPic();
  const unsigned curNumParts = pic->getNumPartInCU() >> (depth << 1);
  const unsigned qNumParts = curNumParts >> 2;
  const unsigned picW = ctu->getSlice()->getSPS()->getPicWidthInLumaSamples();
  const unsigned picH = ctu->getSlice()->getSPS()->getPicHeightInLumaSamples();
  if (ctu->getDepth(absZorderIdx) > depth)
  {
    for (unsigned p = 0; p < 4; p++, absZorderIdx += qNumParts)
    {
      unsigned lx = ctu->getCUPelX() + g_auiRasterToPelX[g_auiZscanToRaster[absZorderIdx]];
      unsigned ty = ctu->getCUPelY() + g_auiRasterToPelY[g_auiZscanToRaster[absZorderIdx]];
      if (lx < picW && ty < picH) bsWalk(ctu, absZorderIdx, depth + 1, lf);
    }
    return;
  }
  lf->xSetLoopfilterParam(ctu, absZorderIdx);
  TComTURecurse tuRecurse(ctu, absZorderIdx);
  lf->xSetEdgefilterTU(tuRecurse);
  lf->xSetEdgefilterPU(ctu, absZorderIdx);

  const int cuX = ctu->getCUPelX() + g_auiRasterToPelX[g_auiZscanToRaster[absZorderIdx]];
  const int cuY = ctu->getCUPelY() + g_auiRasterToPelY[g_auiZscanToRaster[absZorderIdx]];
  const int cuSize = g_uiMaxCUWidth >> depth;
  static const bool fastBs = getenv("HMDEC_B200_HM_BS") == NULL;     // HMDEC_B200_HM_BS=1: HM's routine for every unit
  const unsigned numPartInWidth = pic->getNumPartInWidth();
  for (int dir = 0; dir < 2; dir++)
  {
    const Bool* edgeFlag = lf->m_aapbEdgeFilter[dir];
    for (unsigned part = absZorderIdx; part < absZorderIdx + curNumParts; part++)
    {
      // most partitions carry no edge at all (large CUs): skip runs of 8 clear flags at once (CUs start at multiples of 4
      // partitions and the flag arrays are 8-byte aligned allocations)
      if ((part & 7) == 0 && part + 8 <= absZorderIdx + curNumParts)
      {
        uint64_t w;
        memcpy(&w, edgeFlag + part, 8);
        if (w == 0) { part += 7; continue; }
      }
      if (!edgeFlag[part]) continue;
      // only partitions on the 8x8 luma grid carry an edge (uiBSCheck, TComLoopFilter.cpp:199-206)
      const unsigned raster = g_auiZscanToRaster[part];
      const int ux = g_auiRasterToPelX[raster] >> 2, uy = g_auiRasterToPelY[raster] >> 2;
      const bool onGrid = (dir == EDGE_VER) ? ((ux & 1) == 0) : ((uy & 1) == 0);
      if (!onGrid) continue;
      if (lf->m_aapbEdgeFilter[dir][part])
      {
        unsigned bs;
        if (fastBs)
        {
          TComDataCU* cuP = ctu;
          unsigned partP;
          if (dir == EDGE_VER)
          {
            if (ux > 0) partP = g_auiRasterToZscan[raster - 1];
            else { cuP = ctu->getCULeft(); partP = g_auiRasterToZscan[raster + numPartInWidth - 1]; }
          }
          else
          {
            if (uy > 0) partP = g_auiRasterToZscan[raster - numPartInWidth];
            else { cuP = ctu->getCUAbove(); partP = g_auiRasterToZscan[raster + pic->getNumPartInCU() - numPartInWidth]; }
          }
          bs = bsOfUnit(cuP, partP, ctu, part, lf->m_aapucBS[dir][part] != 0);
        }
        else
        {
          lf->xGetBoundaryStrengthSingle(ctu, DeblockEdgeDir(dir), part);
          bs = lf->m_aapucBS[dir][part];
        }
        if (bs)
        {
          const int gx = (ctu->getCUPelX() >> 2) + ux, gy = (ctu->getCUPelY() >> 2) + uy;
          for (int j = 0; j < (1 << m_lgU); j++)                // an 8x8 partition's edge is two 4-sample units long
            m_bs[(size_t)(gy + (dir == EDGE_VER ? j : 0)) * m_bsStride + gx + (dir == EDGE_VER ? 0 : j)] |= (uint8_t)(bs << (dir == EDGE_VER ? 0 : 2));
        }
      }
    }
  }
  // QP / no-filter maps per 8x8 (TComLoopFilter.cpp:587-600,607-617)
  const int qp = ctu->getQP(absZorderIdx);
  const bool pcmFilterOff = ctu->getSlice()->getSPS()->getUsePCM() && ctu->getSlice()->getSPS()->getPCMFilterDisableFlag();
  const bool nofilter = (pcmFilterOff && ctu->getIPCMFlag(absZorderIdx)) || ctu->isLosslessCoded(absZorderIdx);
  if (nofilter) m_hdr.flags |= HMR_FRM_HAS_NOFILTER;
  for (int y = cuY >> 3; y < ((cuY + cuSize) >> 3) && y < ((m_hdr.height + 7) >> 3); y++)
    for (int x = cuX >> 3; x < ((cuX + cuSize) >> 3) && x < m_qpStride; x++)
    {
      m_qp[(size_t)y * m_qpStride + x] = (int8_t)qp;
      m_cuFlags[(size_t)y * m_qpStride + x] = nofilter ? HMR_CU_NOFILTER : 0;
    }
}

// The same side info without HM's flag arrays: the edges of a leaf CU are enumerated from its geometry.  Per 4x4 unit on the
// 8x8 grid (TComLoopFilter.cpp:199-206) an edge exists and counts as a transform edge exactly where xSetLoopfilterParam +
// xSetEdgefilterTU + xSetEdgefilterPU (TComLoopFilter.cpp:270-409) leave their marks:
//   CU border (unit column / row 0)   filter = TU marker = "the neighbour exists for deblocking" (slice / tile crossing rules of
//                                     getPULeft / getPUAbove; a neighbour inside the same CTU always exists)
//   inside the CU                     TU marker = the unit starts a transform block (size = CU size >> stored transform index);
//                                     filter = TU marker, or the unit lies on the CU's prediction-partition boundary
// each under "deblocking not disabled in this slice".  The flag-array version (bsWalk) stays as the A/B reference:
// HMDEC_B200_BS_FLAGS=1 selects it, tests/test_frontend_fast_path.py requires byte-identical BS maps from both.
void HmEmitter::bsDirect(TComDataCU* ctu, unsigned absZorderIdx, unsigned depth, bool lfCrossTiles)
{
  if (ctu->getPic() == 0 || ctu->getPartitionSize(absZorderIdx) == NUMBER_OF_PART_SIZES) return;
  TComPic* pic = ctu->getPic();
  const unsigned curNumParts = pic->getNumPartInCU() >> (depth << 1);
  const unsigned qNumParts = curNumParts >> 2;
  TComSlice* slice = ctu->getSlice();
  const unsigned picW = slice->getSPS()->getPicWidthInLumaSamples();
  const unsigned picH = slice->getSPS()->getPicHeightInLumaSamples();
  if (ctu->getDepth(absZorderIdx) > depth)
  {
    for (unsigned p = 0; p < 4; p++, absZorderIdx += qNumParts)
    {
      unsigned lx = ctu->getCUPelX() + g_auiRasterToPelX[g_auiZscanToRaster[absZorderIdx]];
      unsigned ty = ctu->getCUPelY() + g_auiRasterToPelY[g_auiZscanToRaster[absZorderIdx]];
      if (lx < picW && ty < picH) bsDirect(ctu, absZorderIdx, depth + 1, lfCrossTiles);
    }
    return;
  }
  const unsigned raster0 = g_auiZscanToRaster[absZorderIdx];
  const unsigned stride = pic->getNumPartInWidth();
  const int ux0 = g_auiRasterToPelX[raster0] >> 2, uy0 = g_auiRasterToPelY[raster0] >> 2;
  const int cuX = ctu->getCUPelX() + 4 * ux0, cuY = ctu->getCUPelY() + 4 * uy0;
  const int cuSize = g_uiMaxCUWidth >> depth;
  const int n = cuSize >> 2;
  const bool internal = !slice->getDeblockingFilterDisable();
  if (internal)
  {
    UInt tmp;
    const bool sliceRestr = !slice->getLFCrossSliceBoundaryFlag();
    const bool border[2] = {
      cuX > 0 && (ux0 > 0 || ctu->getPULeft(tmp, absZorderIdx, sliceRestr, !lfCrossTiles) != NULL),
      cuY > 0 && (uy0 > 0 || ctu->getPUAbove(tmp, absZorderIdx, sliceRestr, false, !lfCrossTiles) != NULL) };
    int puEdge[2] = { -1, -1 };                              // unit column (EDGE_VER) / row (EDGE_HOR) of the partition boundary inside the CU
    switch (ctu->getPartitionSize(absZorderIdx))
    {
      case SIZE_2NxN:  puEdge[EDGE_HOR] = n >> 1; break;
      case SIZE_Nx2N:  puEdge[EDGE_VER] = n >> 1; break;
      case SIZE_NxN:   puEdge[EDGE_VER] = puEdge[EDGE_HOR] = n >> 1; break;
      case SIZE_2NxnU: puEdge[EDGE_HOR] = n >> 2; break;
      case SIZE_2NxnD: puEdge[EDGE_HOR] = n - (n >> 2); break;
      case SIZE_nLx2N: puEdge[EDGE_VER] = n >> 2; break;
      case SIZE_nRx2N: puEdge[EDGE_VER] = n - (n >> 2); break;
      default: break;
    }
    const bool oneTU = ctu->getTransformIdx(absZorderIdx) == 0;          // the CU is a single transform block: no transform edge inside
    const unsigned numPartInCtu = pic->getNumPartInCU();
    for (int dir = 0; dir < 2; dir++)
      for (int e = 0; e < n; e += 2)                          // CU origins are multiples of 8 samples: e even = on the 8x8 grid
      {
        if (e == 0 ? !border[dir] : (oneTU && e != puEdge[dir])) continue;
        for (int k = 0; k < n; k += 1 << m_lgU)                // x, y, e, k: 4-sample units; raster: HM partitions (4 or 8 samples)
        {
          const int x = dir == EDGE_VER ? e : k, y = dir == EDGE_VER ? k : e;
          const unsigned raster = raster0 + (y >> m_lgU) * stride + (x >> m_lgU);
          const unsigned part = g_auiRasterToZscan[raster];
          bool tuEdge = true;
          if (e > 0)
          {
            tuEdge = !oneTU && ((4 * e) & ((cuSize >> ctu->getTransformIdx(part)) - 1)) == 0;
            if (!tuEdge && e != puEdge[dir]) continue;
          }
          TComDataCU* cuP = ctu;
          unsigned partP;
          if (dir == EDGE_VER)
          {
            if (ux0 + x > 0) partP = g_auiRasterToZscan[raster - 1];
            else { cuP = ctu->getCULeft(); partP = g_auiRasterToZscan[raster + stride - 1]; }
          }
          else
          {
            if (uy0 + y > 0) partP = g_auiRasterToZscan[raster - stride];
            else { cuP = ctu->getCUAbove(); partP = g_auiRasterToZscan[raster + numPartInCtu - stride]; }
          }
          const unsigned bs = bsOfUnit(cuP, partP, ctu, part, tuEdge);
          if (bs)
          {
            const int gx = (cuX >> 2) + x, gy = (cuY >> 2) + y;
            for (int j = 0; j < (1 << m_lgU); j++)
              m_bs[(size_t)(gy + (dir == EDGE_VER ? j : 0)) * m_bsStride + gx + (dir == EDGE_VER ? 0 : j)] |= (uint8_t)(bs << (dir == EDGE_VER ? 0 : 2));
          }
        }
      }
  }
  // QP / no-filter maps per 8x8 (TComLoopFilter.cpp:587-600,607-617)
  const int qp = ctu->getQP(absZorderIdx);
  const bool pcmFilterOff = slice->getSPS()->getUsePCM() && slice->getSPS()->getPCMFilterDisableFlag();
  const bool nofilter = (pcmFilterOff && ctu->getIPCMFlag(absZorderIdx)) || ctu->isLosslessCoded(absZorderIdx);
  if (nofilter) m_hdr.flags |= HMR_FRM_HAS_NOFILTER;
  for (int y = cuY >> 3; y < ((cuY + cuSize) >> 3) && y < ((m_hdr.height + 7) >> 3); y++)
    for (int x = cuX >> 3; x < ((cuX + cuSize) >> 3) && x < m_qpStride; x++)
    {
      m_qp[(size_t)y * m_qpStride + x] = (int8_t)qp;
      m_cuFlags[(size_t)y * m_qpStride + x] = nofilter ? HMR_CU_NOFILTER : 0;
    }
}

// Deblocking side info of one CTU, right after it was parsed (HM derives it for the whole picture at the end, TComLoopFilter.cpp:
// 130-155, when every CTU's arrays have long left the caches): edge flags with HM's own xSetLoopfilterParam / xSetEdgefilterTU /
// xSetEdgefilterPU on an emitter-owned TComLoopFilter, strengths with bsOfUnit.  Only the left and the above CTU are consulted and
// both precede this one in decoding order (also across tiles: the tile to the left / above is decoded first).
void HmEmitter::deblockCtu(TComDataCU* ctu)
{
  TComSlice* s = ctu->getSlice();
  static const bool flagArrays = getenv("HMDEC_B200_BS_FLAGS") != NULL || getenv("HMDEC_B200_HM_BS") != NULL;   // A/B: HM's edge-flag machinery
  if (!flagArrays) bsDirect(ctu, 0, 0, s->getPPS()->getLoopFilterAcrossTilesEnabledFlag());
  else
  {
    if (!m_lf || m_lfDepth != g_uiMaxCUDepth)
    {
      if (m_lf) { m_lf->destroy(); delete m_lf; }
      m_lf = new TComLoopFilter;
      m_lf->create(g_uiMaxCUDepth);
      m_lfDepth = g_uiMaxCUDepth;
    }
    TComLoopFilter* lf = m_lf;
    lf->setCfg(s->getPPS()->getLoopFilterAcrossTilesEnabledFlag());
    ::memset(lf->m_aapucBS[EDGE_VER], 0, sizeof(UChar) * lf->m_uiNumPartitions);
    ::memset(lf->m_aapbEdgeFilter[EDGE_VER], 0, sizeof(Bool) * lf->m_uiNumPartitions);
    ::memset(lf->m_aapucBS[EDGE_HOR], 0, sizeof(UChar) * lf->m_uiNumPartitions);
    ::memset(lf->m_aapbEdgeFilter[EDGE_HOR], 0, sizeof(Bool) * lf->m_uiNumPartitions);
    bsWalk(ctu, 0, 0, lf);
  }
  const UInt a = ctu->getAddr();
  m_ctu[a].beta_offset_div2 = (int8_t)s->getDeblockingFilterBetaOffsetDiv2();
  m_ctu[a].tc_offset_div2   = (int8_t)s->getDeblockingFilterTcOffsetDiv2();
  if (!s->getDeblockingFilterDisable()) m_anyDeblock = true;
}

// SAO side info: reconstructBlkSAOParams (TComSampleAdaptiveOffset.cpp:348-372) then per-CTU flattening
void HmEmitter::saoInfo(TComPic* pic, TComSampleAdaptiveOffset* sao)
{
  SAOBlkParam* blk = pic->getPicSym()->getSAOBlkParam();
  sao->reconstructBlkSAOParams(pic, blk);
  const int nComp = getNumberValidComponents(pic->getChromaFormat());
  bool any = false;
  for (UInt a = 0; a < pic->getNumCUsInFrame(); a++)
  {
    Bool l, r, ab, be, al, ar, bl, br;
    pic->getPicSym()->deriveLoopFilterBoundaryAvailibility(a, l, r, ab, be, al, ar, bl, br);
    m_ctu[a].avail = (uint8_t)((l ? HMR_AV_L : 0) | (r ? HMR_AV_R : 0) | (ab ? HMR_AV_A : 0) | (be ? HMR_AV_B : 0) |
                               (al ? HMR_AV_AL : 0) | (ar ? HMR_AV_AR : 0) | (bl ? HMR_AV_BL : 0) | (br ? HMR_AV_BR : 0));
    for (int c = 0; c < nComp; c++)
    {
      SAOOffset& o = blk[a][c];
      hmr_sao& d = m_ctu[a].sao[c];
      if (o.modeIdc == SAO_MODE_OFF) { d.type = HMR_SAO_OFF; continue; }
      any = true;
      if (o.typeIdc == SAO_TYPE_START_BO)
      {
        d.type = HMR_SAO_BO;
        d.band = (uint8_t)o.typeAuxInfo;
        for (int i = 0; i < 4; i++) d.off[i] = (int16_t)o.offset[(o.typeAuxInfo + i) % NUM_SAO_BO_CLASSES];
      }
      else
      {
        d.type = (uint8_t)(HMR_SAO_EO_0 + (o.typeIdc - SAO_TYPE_START_EO));
        // offset[] is indexed by edgeIdx+2 with the "plain" class (edgeIdx 0) fixed at 0 (TComSampleAdaptiveOffset.cpp:245-249)
        d.off[0] = (int16_t)o.offset[0]; d.off[1] = (int16_t)o.offset[1];
        d.off[2] = (int16_t)o.offset[3]; d.off[3] = (int16_t)o.offset[4];
      }
    }
  }
  if (any) m_hdr.flags |= HMR_FRM_SAO;
}

// ---------------------------------------------------------------------------------------------
void HmEmitter::onPictureParsed(TComPic* pic, TComLoopFilter* lf, TComSampleAdaptiveOffset* sao, bool lfCrossTiles)
{
  if (!m_open || pic != m_curPic) { fail("filterPicture for a picture with no parsed CTU"); return; }
  const double t0 = nowSec();
  TComSlice* slice = pic->getSlice(pic->getCurrSliceIdx());
  if (getenv("HMDEC_B200_BS_AT_END")) for (UInt a = 0; a < pic->getNumCUsInFrame(); a++) deblockCtu(pic->getCU(a));
  if (m_anyDeblock) m_hdr.flags |= HMR_FRM_DEBLOCK;              // the maps were filled CTU by CTU (deblockCtu)
  const double t1 = nowSec();
  if (slice->getSPS()->getUseSAO()) saoInfo(pic, sao);
  if (slice->getSPS()->getUsePCM()) { /* PCM CUs themselves are rejected in emitIntraCU */ }
  if (slice->isReferenced()) m_hdr.flags |= HMR_FRM_IS_REFERENCE;
  if (m_pu.empty()) m_hdr.flags |= HMR_FRM_INTRA_ONLY;

  // pad the coefficient buffer to a multiple of 16 entries
  while (m_coef.size() & 15) *m_coef.grow(1) = 0;
  // group the residual records by transform size (stable: luma stays ahead of its co-located chroma for CCP)
  {
    std::vector<hmr_tu>& sorted = m_tuSorted;                  // a member: both vectors keep their capacity from picture to picture
    sorted.clear();
    sorted.reserve(m_tu.size());
    for (int k = 0; k < 4; k++)
    {
      m_hdr.tu_first[k] = (uint32_t)sorted.size();
      for (size_t i = 0; i < m_tu.size(); i++) if (m_tu[i].log2_size == k + 2) sorted.push_back(m_tu[i]);
    }
    m_hdr.tu_first[4] = (uint32_t)sorted.size();
    m_tu.swap(sorted);
  }
  m_hdr.n_tu = (uint32_t)m_tu.size();
  m_hdr.n_coef = (uint32_t)m_coef.size();
  m_hdr.n_intra = (uint32_t)m_intra.size();
  m_hdr.n_pu = (uint32_t)m_pu.size();
  m_hdr.n_mc_tiles = m_puPrefix.back();

  hmr_frame_desc d;
  d.hdr = &m_hdr;
  d.tu = m_tu.data(); d.coef = m_coef.data();
  d.intra = m_intra.data(); d.intra_range = m_range.data();
  d.pu = m_pu.data(); d.pu_tile_prefix = m_puPrefix.data();
  d.ctu = m_ctu.data();
  d.scaling = (m_hdr.flags & HMR_FRM_SCALING_LIST) ? m_scaling.data() : NULL;
  d.wp = (m_hdr.flags & HMR_FRM_WEIGHTED_PRED) ? m_wp.data() : NULL;
  d.pu_refidx = (m_hdr.flags & HMR_FRM_WEIGHTED_PRED) ? m_puRefIdx.data() : NULL;
  d.bs = (m_hdr.flags & HMR_FRM_DEBLOCK) ? m_bs.data() : NULL;
  d.qp = m_qp.data();
  d.cu_flags = (m_hdr.flags & HMR_FRM_HAS_NOFILTER) ? m_cuFlags.data() : NULL;
  const double t2 = nowSec();
  // once anything was flagged the records may be wrong: nothing reaches the engine any more, push_nal_unit reports LIBHMDEC_ERROR
  if (!m_unsupported && !m_sink->frameReady(d, pic))
  {
    m_failText = std::string("the reconstruction engine rejected the picture: ") + (m_sink->error() ? m_sink->error() : "unknown error");
    fail(m_failText.c_str());
  }
  const double t3 = nowSec();
  m_tBs += t1 - t0; m_tPic += t2 - t1; m_tSink += t3 - t2; m_nPic++;
  m_open = false;
}
