// hm_cabac.cpp — the product's residual-coefficient parser: TDecSbac::parseCoeffNxN re-implemented on
// an inlined arithmetic decoder, behind HM's own TDecEntropyIf slot.
//
// HM's routine (TLibDecoder/TDecSbac.cpp:1253-1632, renamed parseCoeffNxN_hm at build time: frontend/Makefile) decodes every bin
// through a virtual call into TDecBinCABAC (TDecBinCoderCABAC.cpp:108-200), derives the significance context of every
// coefficient with a function call (TComTrQuant::getSigCtxInc, TComTrQuant.cpp:2552-2650) and finds the scan position of the
// last coefficient by a linear search.  On an intra picture that is 54 % of the parser thread (sampling profile, 2160p QP32).
// This file keeps HM's data model — the same ContextModel objects, the same engine state (range / value / bitsNeeded /
// byte position are loaded from and stored back into HM's objects, so every other syntax element is still parsed by HM) — and
// restates residual_coding() (H.265 7.3.8.11, 9.3.4.2.4-9.3.4.2.7) with
//   * the engine in registers, refilled from the raw byte array with HM's exact schedule (so the position HM sees afterwards
//     is the one it would have reached itself),
//   * bypass bins decoded several at a time (one division instead of a compare/subtract per bin),
//   * per (channel type, block size, scan) tables, built once FROM HM's own functions: significance context per scan position
//     and neighbour pattern, and the inverse scan.
// Streams that use range-extension entropy tools (extended precision, CABAC bypass alignment, persistent Rice adaptation, single
// significance context, RDPCM, transquant bypass) take HM's own routine.  tests/test_frontend_fast_path.py proves the records byte-identical
// to the goldens (which came from HM's routine); HMDEC_B200_HM_COEFF=1 selects HM's routine for A/B runs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <list>
#include <map>
#include <iostream>
#include <sstream>
#include <fstream>
#include <iomanip>
#include <algorithm>
#include <utility>
#include <limits>
#include <deque>
#include <set>
#include <cassert>
#include <cmath>
#include <stdint.h>
#define private public
#define protected public
#include "TLibCommon/TComBitStream.h"
#include "TLibDecoder/TDecBinCoderCABAC.h"
#include "TLibDecoder/TDecSbac.h"
#undef private
#undef protected
#include "TLibCommon/TComTU.h"
#include "TLibCommon/TComTrQuant.h"
#include "TLibCommon/TComCABACTables.h"
#include "TLibCommon/TComChromaFormat.h"
#include "TLibCommon/ContextTables.h"

namespace {

// state transitions of a context model, [0..127] after an MPS, [128..255] after an LPS, index = (state << 1) | mps: read out of
// HM's ContextModel once (buildTables)
static uint8_t g_nextState[256];

// ---- arithmetic decoder, HM's state representation (value is the 16+7-bit window of TDecBinCABAC) ----
struct Engine
{
  uint32_t range, value;
  int      bitsNeeded;
  const uint8_t* p;

  // One context-coded bin (H.265 9.3.4.3.2), the same arithmetic as TDecBinCABAC::decodeBin but without a data-dependent branch:
  // whether the bin took the LPS path is a mask.  The renormalisation shift is the number of leading zeros of the 9-bit
  // range (what sm_aucRenormTable / the one-bit MPS case of HM produce), the refill schedule is HM's.
  inline unsigned bin(ContextModel& cm)
  {
    const unsigned s = cm.m_ucState;
    const uint32_t lps = TComCABACTables::sm_aucLPSTable[s >> 1][(range >> 6) & 3];
    range -= lps;
    const uint32_t scaled = range << 7;
    const uint32_t mask = (uint32_t)((int32_t)(scaled - value - 1) >> 31);      // all ones: value >= scaled, the LPS path
    value -= scaled & mask;
    range += (lps - range) & mask;
    cm.m_ucState = g_nextState[(mask & 128) + s];
    const int shift = __builtin_clz(range) - 23;
    range <<= shift; value <<= shift;
    bitsNeeded += shift;
    if (bitsNeeded >= 0) { value += (uint32_t)*p++ << bitsNeeded; bitsNeeded -= 8; }
    return (s ^ mask) & 1;
  }

  inline unsigned bypass()
  {
    value += value;
    if (++bitsNeeded >= 0) { bitsNeeded = -8; value += *p++; }
    const uint32_t scaled = range << 7;
    if (value >= scaled) { value -= scaled; return 1; }
    return 0;
  }

  // n bypass bins, first decoded = most significant (the long division HM does bit by bit, TDecBinCoderCABAC.cpp:233-282)
  inline uint32_t bypassBins(int n)
  {
    uint32_t bins = 0;
    const uint32_t scaled = range << 7;
    while (n > 8)
    {
      value = (value << 8) + ((uint32_t)*p++ << (8 + bitsNeeded));
      const uint32_t q = value / scaled;
      value -= q * scaled;
      bins = (bins << 8) | q;
      n -= 8;
    }
    bitsNeeded += n;
    value <<= n;
    if (bitsNeeded >= 0) { value += (uint32_t)*p++ << bitsNeeded; bitsNeeded -= 8; }
    const uint32_t q = value / scaled;
    value -= q * scaled;
    return (bins << n) | q;
  }
};

// ---- tables derived from HM's own functions, per (channel type, log2 size, scan) ----
struct ScanTables
{
  std::vector<UInt> scan;         // HM's grouped 4x4 scan: scan position -> raster position (a COPY: HM's ROM arrays are freed
  std::vector<UInt> scanCG;       // scan order of the coefficient groups                  when the last decoder of the process goes)
  std::vector<uint16_t> inverse;  // raster position -> scan position
  std::vector<uint8_t>  sigCtx;   // [pattern 0..3][scan position] -> context increment (TComTrQuant::getSigCtxInc)
};

static ScanTables g_tables[2][4][3];      // [channel type][log2 size - 2][scan type]

static void buildTables()
{
  for (int v = 0; v < 128; v++)
  {
    ContextModel cm;
    cm.setStateAndMps((UChar)(v >> 1), (UChar)(v & 1)); cm.updateMPS(); g_nextState[v] = (uint8_t)((cm.getState() << 1) | cm.getMps());
    cm.setStateAndMps((UChar)(v >> 1), (UChar)(v & 1)); cm.updateLPS(); g_nextState[128 + v] = (uint8_t)((cm.getState() << 1) | cm.getMps());
  }
  for (int ch = 0; ch < 2; ch++)
    for (int lg = 2; lg <= 5; lg++)
      for (int st = 0; st < 3; st++)
      {
        ScanTables& t = g_tables[ch][lg - 2][st];
        const int n = 1 << (2 * lg);
        TUEntropyCodingParameters cp;
        cp.scanType = COEFF_SCAN_TYPE(st);
        cp.widthInGroups = cp.heightInGroups = (1u << lg) >> 2;
        const UInt lgGroups = g_aucConvertToBit[cp.widthInGroups * 4];
        cp.scan   = g_scanOrder[SCAN_GROUPED_4x4][st][lg][lg];
        cp.scanCG = g_scanOrder[SCAN_UNGROUPED][st][lgGroups][lgGroups];
        // (TComChromaFormat.cpp:110-131, without the single-context mode, which takes HM's routine)
        if (lg == 2) cp.firstSignificanceMapContext = significanceMapContextSetStart[ch][CONTEXT_TYPE_4x4];
        else if (lg == 3) cp.firstSignificanceMapContext = significanceMapContextSetStart[ch][CONTEXT_TYPE_8x8] + (st != SCAN_DIAG ? nonDiagonalScan8x8ContextOffset[ch] : 0);
        else cp.firstSignificanceMapContext = significanceMapContextSetStart[ch][CONTEXT_TYPE_NxN];
        t.scan.assign(cp.scan, cp.scan + n); t.scanCG.assign(cp.scanCG, cp.scanCG + (n >> 4));
        t.inverse.resize(n);
        for (int s = 0; s < n; s++) t.inverse[cp.scan[s]] = (uint16_t)s;
        t.sigCtx.resize((size_t)4 * n);
        for (int pat = 0; pat < 4; pat++)
          for (int s = 0; s < n; s++)
            t.sigCtx[(size_t)pat * n + s] = (uint8_t)TComTrQuant::getSigCtxInc(pat, cp, s, lg, lg, ChannelType(ch));
      }
}

static bool useHmRoutine() { static const bool on = getenv("HMDEC_B200_HM_COEFF") != NULL; return on; }

// The byte array behind a TComInputBitstream.  m_fifo is the first data member of a class without virtual functions
// (TComBitStream.h:164-166), private by default access; fifoOf() reads it at offset 0 and fifoLayoutOk() checks that reading once
// against the class's own peekPreviousByte() — if it ever disagreed, every block would take HM's routine.
static inline const std::vector<uint8_t>* fifoOf(const TComInputBitstream* bs) { return *reinterpret_cast<const std::vector<uint8_t>* const*>(bs); }
static bool fifoLayoutOk(TComInputBitstream* bs)
{
  const std::vector<uint8_t>* f = fifoOf(bs);
  const UInt idx = bs->getByteLocation();
  if (!f || idx == 0 || idx > f->size() || 8 * (UInt)(f->size() - idx) != bs->getNumBitsLeft() - bs->getNumBitsUntilByteAligned()) return false;
  UInt prev = 0;
  bs->peekPreviousByte(prev);
  return prev == (*f)[idx - 1];
}

}  // namespace

Void TDecSbac::parseCoeffNxN(TComTU& rTu, ComponentID compID)
{
  TComDataCU* cu = rTu.getCU();
  const UInt absPartIdx = rTu.GetAbsPartIdxTU(compID);
  const TComRectangle& rect = rTu.getRect(compID);
  const TComSPS* sps = cu->getSlice()->getSPS();
  TComPPS* pps = const_cast<TComPPS*>(cu->getSlice()->getPPS());
  const UInt N = rect.width;
  if (useHmRoutine() || rect.width != rect.height || N > 32 || N < 4 ||
      sps->getUseExtendedPrecision() || sps->getAlignCABACBeforeBypass() || sps->getUseGolombRiceParameterAdaptation() ||
      sps->getUseSingleSignificanceMapContext() || cu->getCUTransquantBypass(absPartIdx) || cu->isRDPCMEnabled(absPartIdx))
  {
    parseCoeffNxN_hm(rTu, compID);
    return;
  }
  TDecBinCABAC* hm = m_pcTDecBinIf->getTDecBinCABAC();
  TComInputBitstream* bs = hm ? hm->m_pcTComBitstream : NULL;
  static const bool tablesReady = (buildTables(), true);     // once, thread-safe; the ROM tables exist by the time a block is parsed
  static const bool layoutOk = bs && fifoLayoutOk(bs);       // (checked on the first block of the process)
  if (!tablesReady || !layoutOk || !bs) { parseCoeffNxN_hm(rTu, compID); return; }

  TCoeff* coef = cu->getCoeff(compID) + rTu.getCoefficientOffset(compID);
  ::memset(coef, 0, sizeof(TCoeff) * N * N);                 // HM's per-CTU zero fill of the whole coefficient storage is skipped (hm_fast.cpp)
  if (pps->getUseTransformSkip()) parseTransformSkipFlags(rTu, compID);     // HM's own (one bin on its own engine, before ours is loaded)
  const bool signHiding = pps->getSignHideFlag() > 0;

  const int lg = g_aucConvertToBit[N] + 2;
  const ChannelType chType = toChannelType(compID);
  const int ch = chType == CHANNEL_TYPE_LUMA ? 0 : 1;
  const int scanType = (int)cu->getCoefScanIdx(absPartIdx, N, N, compID);
  const ScanTables& T = g_tables[ch][lg - 2][scanType];

  const uint8_t* const base = fifoOf(bs)->data();
  Engine e;
  e.range = hm->m_uiRange; e.value = hm->m_uiValue; e.bitsNeeded = hm->m_bitsNeeded; e.p = base + bs->m_fifo_idx;

  // ---- last significant coefficient (prefix: context-coded unary per axis, suffix: bypass) ----
  int lastX, lastY;
  {
    ContextModel* ctxX = m_cCuCtxLastX.get(0, chType);
    ContextModel* ctxY = m_cCuCtxLastY.get(0, chType);
    Int offX, offY, shX, shY;
    getLastSignificantContextParameters(compID, N, N, offX, offY, shX, shY);
    const int maxPrefix = (int)g_uiGroupIdx[N - 1];
    int px = 0, py = 0;
    while (px < maxPrefix && e.bin(ctxX[offX + (px >> shX)])) px++;
    while (py < maxPrefix && e.bin(ctxY[offY + (py >> shY)])) py++;
    if (px > 3) px = (int)g_uiMinInGroup[px] + (int)e.bypassBins((px - 2) >> 1);
    if (py > 3) py = (int)g_uiMinInGroup[py] + (int)e.bypassBins((py - 2) >> 1);
    if (scanType == SCAN_VER) { lastX = py; lastY = px; } else { lastX = px; lastY = py; }
  }
  const int lastRaster = lastX + (lastY << lg);
  const int lastScan = T.inverse[lastRaster];
  const int lastGroup = lastScan >> 4;
  const int lgGroups = lg - 2, groupsW = 1 << lgGroups;

  ContextModel* const ctxGroup = m_cCUSigCoeffGroupSCModel.get(0, ch);
  ContextModel* const ctxSig   = m_cCUSigSCModel.get(0, 0) + getSignificanceMapContextOffset(compID);
  const uint8_t* const sigTab  = T.sigCtx.data();
  const int NN = 1 << (2 * lg);

  uint64_t groupCoded = 0;                // bit = coefficient group (raster index in the block) has a coefficient
  bool greater1Seen = false;              // the previous group ended with c1 == 0 (selects the context set of the next)
  int scanPos = lastScan;

  for (int g = lastGroup; g >= 0; g--)
  {
    const int groupStart = g << 4;
    const int gRaster = (int)T.scanCG[g];
    const int gy = gRaster >> lgGroups, gx = gRaster & (groupsW - 1);
    const unsigned right = (gx + 1 < groupsW) ? (unsigned)((groupCoded >> (gRaster + 1)) & 1) : 0;
    const unsigned below = (gy + 1 < groupsW) ? (unsigned)((groupCoded >> (gRaster + groupsW)) & 1) : 0;

    int pos[17], level[16];
    int count = 0;
    unsigned sigMask = 0;                                   // bit = position inside the group (scan order) holds a coefficient
    if (scanPos == lastScan)
    {
      pos[0] = lastRaster; count = 1;
      sigMask = 1u << (scanPos - groupStart);
      scanPos--;
    }
    bool coded = true;
    if (g != lastGroup && g != 0) coded = e.bin(ctxGroup[right | below]) != 0;
    if (coded)
    {
      groupCoded |= 1ull << gRaster;
      const uint8_t* tab = sigTab + (size_t)(groupsW > 1 ? (right | (below << 1)) : 0) * NN;
      for (; scanPos >= groupStart; scanPos--)
      {
        unsigned sig;
        if (scanPos > groupStart || g == 0 || count) sig = e.bin(ctxSig[tab[scanPos]]);
        else sig = 1;                                       // the group is coded and nothing else in it was: its first coefficient must be
        pos[count] = (int)T.scan[scanPos];                  // (overwritten by the next one if this one is zero)
        count += (int)sig;
        sigMask |= sig << (scanPos - groupStart);
      }
    }
    else scanPos = groupStart - 1;
    if (!count) continue;

    // ---- levels of the group: greater-than-1 flags (at most 8), one greater-than-2 flag, signs, Golomb-Rice remainders ----
    const bool hidden = signHiding && ((31 - __builtin_clz(sigMask)) - __builtin_ctz(sigMask) >= SBH_THRESHOLD);   // distance first .. last coefficient of the group
    const UInt ctxSet = getContextSetIndex(compID, (UInt)g, greater1Seen);
    ContextModel* ctxOne = m_cCUOneSCModel.get(0, 0) + NUM_ONE_FLAG_CTX_PER_SET * ctxSet;
    unsigned c1 = 1;
    int firstGreater1 = -1;
    bool escapes = count > C1FLAG_NUMBER;
    const int nFlags = count < C1FLAG_NUMBER ? count : C1FLAG_NUMBER;
    for (int i = 0; i < count; i++) level[i] = 1;
    for (int i = 0; i < nFlags; i++)
    {
      const unsigned b = e.bin(ctxOne[c1]);
      if (b)
      {
        c1 = 0;
        if (firstGreater1 < 0) firstGreater1 = i; else escapes = true;
      }
      else if (c1 > 0 && c1 < 3) c1++;
      level[i] = 1 + (int)b;
    }
    greater1Seen = (c1 == 0);
    if (firstGreater1 >= 0)
    {
      ContextModel* ctxAbs = m_cCUAbsSCModel.get(0, 0) + NUM_ABS_FLAG_CTX_PER_SET * ctxSet;
      const unsigned b = e.bin(ctxAbs[0]);
      level[firstGreater1] = 2 + (int)b;
      if (b) escapes = true;
    }
    const int nSigns = hidden ? count - 1 : count;
    uint32_t signs = nSigns ? e.bypassBins(nSigns) << (32 - nSigns) : 0;

    if (escapes)
    {
      unsigned rice = 0;
      int firstCoeff2 = 1;
      for (int i = 0; i < count; i++)
      {
        const int baseLevel = i < C1FLAG_NUMBER ? 2 + firstCoeff2 : 1;
        if (level[i] == baseLevel)
        {
          // coeff_abs_level_remaining: unary prefix, then a Rice / exp-Golomb suffix (TDecSbac.cpp xReadCoefRemainExGolomb, no limited prefix)
          unsigned prefix = 0;
          while (prefix < 32 && e.bypass()) prefix++;
          unsigned rem;
          if (prefix < COEF_REMAIN_BIN_REDUCTION) rem = (prefix << rice) + (rice ? e.bypassBins((int)rice) : 0);
          else
          {
            const int suffixBits = (int)(prefix - COEF_REMAIN_BIN_REDUCTION + rice);
            rem = (((1u << (prefix - COEF_REMAIN_BIN_REDUCTION)) + COEF_REMAIN_BIN_REDUCTION - 1) << rice) + (suffixBits ? e.bypassBins(suffixBits) : 0);
          }
          level[i] = (int)rem + baseLevel;
          if (level[i] > (3 << rice)) rice = rice < 4 ? rice + 1 : 4;
        }
        if (level[i] >= 2) firstCoeff2 = 0;
      }
    }

    int absSum = 0;
    for (int i = 0; i < count; i++)
    {
      absSum += level[i];
      int v = level[i];
      if (i == count - 1 && hidden) { if (absSum & 1) v = -v; }            // the sign of the last one parsed is the parity of the sum
      else { if (signs & 0x80000000u) v = -v; signs <<= 1; }
      coef[pos[i]] = v;
    }
  }

  hm->m_uiRange = e.range; hm->m_uiValue = e.value; hm->m_bitsNeeded = e.bitsNeeded;
  bs->m_fifo_idx = (UInt)(e.p - base);
}
