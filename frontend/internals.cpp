// internals.cpp — libHMDEC_get_internal_info served from a flat per-picture BLOCK INDEX.
//
// Reference behaviour: source/App/libHMDecoder/libHMDecoder.cpp:451-715 — every query walks the quadtree of every CTU
// and push_back()s one 16-byte libHMDec_BlockValue per block, 24 walks for the 24 info types of a picture, after the
// decoder has rewritten the picture's whole motion field to 16x16 granularity (TComPic::compressMotion).
//
// Here the first query for a picture makes ONE pass over its CTUs and leaves three flat tables behind — coding units,
// prediction units, transform-tree nodes, each with position, size and all the values any info type reports, plus the
// table ranges of every CTU — and each of the 24 types is then a linear sweep over one table into a vector sized up
// front.  The index lives in the decoder context and stays valid until the next libHMDec_push_nal_unit (pictures only
// change inside a push).  Nothing is added to the parse path: the pass reads HM's retained per-CTU arrays, like the
// reference, but once instead of 24 times, and it reads the motion field THROUGH the 16x16 decimation
// (TComMotionInfo.cpp:330-350: every run of N partitions reports the prediction mode, MV and reference index of its
// first one) instead of rewriting it, so this build never runs compressMotion at all (TMVP reads through the same
// decimation: hm_fast_memset.h).  The coefficient energies come from one incremental pass per CTU (the reference sums,
// for a TU of n levels, the first n levels of the CTU's buffer: a running sum sampled at the five possible n) and are
// only computed when an energy type is asked for.
//
// Observable quirks of the reference are reproduced on purpose and marked QUIRK; tests/test_internals_vs_reference.py
// compares all 24 lists with the reference wrapper itself on every golden stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <list>
#include <iostream>
#include "TLibCommon/CommonDef.h"
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComRom.h"
#include "libHMDecoder_api.h"
#include "hm_fast.h"
#include "hm_fast_memset.h"
#undef memset

namespace {

enum { MAX_TU_LEVELS = 8 };      // log2(CTU) - 2 + 1 distinct block sizes at most (64 -> 4: 5)

struct CuRec  { uint16_t x, y, size; uint8_t pred, part, skip, bypass, dirLuma, dirChroma, rootCbf; };
struct PuRec  { uint16_t x, y, w, h; uint8_t merge, dir; int8_t ref[2]; int16_t mv[2][2]; };
struct TuRec  { uint16_t x, y, w, h; uint8_t cbf, tskip, level; };     // cbf / tskip: bit c = component c; level = depth + transform depth
struct CtuRec
{
  uint32_t cuEnd, puEnd, tuEnd;        // table ranges end here (they start where the previous CTU's end)
  uint16_t x, y;
  bool     bypassOn;                   // pps.transquant_bypass_enable_flag of the CTU's slice
  bool     tskipOn;                    // pps.transform_skip_enabled_flag of the CTU's slice
  uint8_t  topLevel;                   // smallest level (= largest block) among the CTU's transform-tree nodes
  int      energy[2][MAX_TU_LEVELS];   // [Y, Cb][level]: filled by addEnergies()
};

} // namespace

struct HmInternalsCache
{
  TComPic* pic;
  bool     valid, energies;
  int      sliceIdx;
  int      ctuW, ctuH;
  std::vector<CuRec>  cu;
  std::vector<PuRec>  pu;
  std::vector<TuRec>  tu;
  std::vector<CtuRec> ctu;
  HmInternalsCache() : pic(NULL), valid(false), energies(false), sliceIdx(0), ctuW(0), ctuH(0) {}
};

namespace {

// ---------------------------------------------------------------- index construction ----------------------------------------------------------------

struct Builder
{
  HmInternalsCache& ix;
  TComDataCU* ctu;
  UInt picW, picH, segStart, maxDepth, runMask;
  int baseX, baseY;
  // the CTU's arrays (one load each instead of an accessor call per block)
  const UChar *depthOf, *trIdx, *interDir, *cbf[3], *tskip[3], *dirL, *dirC;
  const Char  *pred, *partSize;
  const Bool  *skip, *merge, *bypass;
  TComCUMvField* mvf[2];
  int nComp;

  Builder(HmInternalsCache& i) : ix(i) {}

  int xOf(UInt part) const { return baseX + (int)g_auiRasterToPelX[g_auiZscanToRaster[part]]; }
  int yOf(UInt part) const { return baseY + (int)g_auiRasterToPelY[g_auiZscanToRaster[part]]; }
  // prediction mode / motion as the reference sees them: after the 16x16 decimation of the motion field
  UInt run(UInt part) const { return part & runMask; }

  void transformNodes(UInt part, UInt depth, UInt trDepth)
  {
    if (trDepth < trIdx[part])
    {
      const UInt q = ctu->getTotalNumPart() >> ((depth + trDepth + 1) << 1);
      for (int i = 0; i < 4; i++) transformNodes(part + i * q, depth, trDepth + 1);
      // QUIRK: the reference does not return here — a split node is reported too, after its children (libHMDecoder.cpp:560-569)
    }
    TuRec t;
    t.x = (uint16_t)xOf(part); t.y = (uint16_t)yOf(part);
    t.w = (uint16_t)(g_uiMaxCUWidth >> (depth + trDepth)); t.h = (uint16_t)(g_uiMaxCUHeight >> (depth + trDepth));
    t.level = (uint8_t)(depth + trDepth);
    t.cbf = 0; t.tskip = 0;
    for (int c = 0; c < 3; c++)
    {
      if (cbf[c] && ((cbf[c][part] >> trDepth) & 1)) t.cbf |= 1 << c;
      if (tskip[c] && tskip[c][part]) t.tskip |= 1 << c;
    }
    ix.tu.push_back(t);
  }

  void predictionUnits(UInt part, UInt depth)
  {
    const PartSize ps = (PartSize)partSize[part];
    const int n = ps == SIZE_2Nx2N ? 1 : (ps == SIZE_NxN ? 4 : 2);
    const UInt step = (g_auiPUOffset[UInt(ps)] << ((maxDepth - depth) << 1)) >> 4;
    const int S = g_uiMaxCUWidth >> depth, H = S >> 1, Q = S >> 2;
    const int cx = xOf(part), cy = yOf(part);
    UInt sub = part;
    for (int i = 0; i < n; i++, sub += step)
    {
      int x = cx, y = cy, w = S, h = S;
      switch (ps)
      {
        case SIZE_2NxN:  h = H; y += i ? H : 0; break;
        case SIZE_Nx2N:  w = H; x += i ? H : 0; break;
        case SIZE_NxN:   w = h = H; x += (i & 1) ? H : 0; y += (i >> 1) ? H : 0; break;
        case SIZE_2NxnU: h = i ? Q + H : Q; y += i ? Q : 0; break;
        case SIZE_2NxnD: h = i ? Q : Q + H; y += i ? Q + H : 0; break;
        case SIZE_nLx2N: w = i ? Q + H : Q; x += i ? Q : 0; break;
        case SIZE_nRx2N: w = i ? Q : Q + H; x += i ? Q + H : 0; break;
        default: break;
      }
      PuRec p;
      p.x = (uint16_t)x; p.y = (uint16_t)y; p.w = (uint16_t)w; p.h = (uint16_t)h;
      p.merge = merge[sub] ? 1 : 0;
      p.dir = interDir[sub];
      const UInt r = run(sub);
      for (int l = 0; l < 2; l++)
      {
        p.ref[l] = (int8_t)mvf[l]->getRefIdx(r);
        p.mv[l][0] = (int16_t)mvf[l]->getMv(r).getHor(); p.mv[l][1] = (int16_t)mvf[l]->getMv(r).getVer();
      }
      ix.pu.push_back(p);
    }
  }

  void codingTree(UInt part, UInt depth)
  {
    const UInt lx = xOf(part), ty = yOf(part);
    const UInt rx = lx + (g_uiMaxCUWidth >> depth) - 1, by = ty + (g_uiMaxCUHeight >> depth) - 1;
    const UInt nParts = ctu->getPic()->getNumPartInCU() >> (depth << 1);
    const UInt scu = ctu->getSCUAddr() + part;
    const bool segmentStartsInside = scu + nParts > segStart && scu < segStart;
    if ((depth < depthOf[part] && depth < g_uiMaxCUDepth - g_uiAddCUDepth) || segmentStartsInside || rx >= picW || by >= picH)
    {
      const UInt q = ctu->getTotalNumPart() >> ((depth + 1) << 1);
      for (int i = 0; i < 4; i++)
        if ((UInt)xOf(part + i * q) < picW && (UInt)yOf(part + i * q) < picH) codingTree(part + i * q, depth + 1);
      return;
    }
    CuRec c;
    c.x = (uint16_t)lx; c.y = (uint16_t)ty; c.size = (uint16_t)(g_uiMaxCUWidth >> depth);
    c.pred = (uint8_t)pred[run(part)];
    c.part = (uint8_t)partSize[part];
    c.skip = skip[part] ? 1 : 0;
    c.bypass = bypass[part] ? 1 : 0;
    c.dirLuma = dirL[part]; c.dirChroma = dirC ? dirC[part] : 0;
    c.rootCbf = 0;
    for (int k = 0; k < nComp; k++) if (cbf[k][part] & 1) c.rootCbf = 1;
    ix.cu.push_back(c);
    if (c.pred == MODE_INTER) predictionUnits(part, depth);
    transformNodes(part, depth, 0);
  }

  void addCtu(TComDataCU* cu)
  {
    ctu = cu;
    baseX = cu->getCUPelX(); baseY = cu->getCUPelY();
    depthOf = cu->getDepth(); trIdx = cu->getTransformIdx(); interDir = cu->getInterDir();
    pred = cu->getPredictionMode(); partSize = cu->getPartitionSize();
    skip = cu->getSkipFlag(); merge = cu->getMergeFlag(); bypass = cu->getCUTransquantBypass();
    dirL = cu->getIntraDir(CHANNEL_TYPE_LUMA); dirC = cu->getIntraDir(CHANNEL_TYPE_CHROMA);
    nComp = cu->getPic()->getNumberValidComponents();
    for (int c = 0; c < 3; c++) { cbf[c] = cu->getCbf(ComponentID(c)); tskip[c] = cu->getTransformSkip(ComponentID(c)); }
    mvf[0] = cu->getCUMvField(REF_PIC_LIST_0); mvf[1] = cu->getCUMvField(REF_PIC_LIST_1);
    maxDepth = cu->getSlice()->getSPS()->getMaxCUDepth();
    codingTree(0, 0);
    CtuRec r;
    memset(&r, 0, sizeof r);
    r.cuEnd = (uint32_t)ix.cu.size(); r.puEnd = (uint32_t)ix.pu.size(); r.tuEnd = (uint32_t)ix.tu.size();
    r.x = (uint16_t)baseX; r.y = (uint16_t)baseY;
    r.bypassOn = cu->getSlice()->getPPS()->getTransquantBypassEnableFlag();
    r.tskipOn = cu->getSlice()->getPPS()->getUseTransformSkip();
    r.topLevel = MAX_TU_LEVELS;
    for (size_t i = ix.ctu.empty() ? 0 : ix.ctu.back().tuEnd; i < ix.tu.size(); i++) r.topLevel = std::min(r.topLevel, ix.tu[i].level);
    ix.ctu.push_back(r);
  }
};

bool buildIndex(HmInternalsCache& ix, TComPic* pic)
{
  TComPicSym* sym = pic->getPicSym();
  if (!sym) return false;
  ix.cu.clear(); ix.pu.clear(); ix.tu.clear(); ix.ctu.clear();
  ix.pic = pic; ix.energies = false;
  ix.sliceIdx = (int)pic->getCurrSliceIdx();
  ix.ctuW = (int)g_uiMaxCUWidth; ix.ctuH = (int)g_uiMaxCUHeight;
  Builder b(ix);
  // QUIRK: picture size and segment start come from the picture's CURRENT (= last decoded) slice for every CTU (libHMDecoder.cpp:609-617)
  TComSlice* slice = pic->getSlice(pic->getCurrSliceIdx());
  b.picW = slice->getSPS()->getPicWidthInLumaSamples(); b.picH = slice->getSPS()->getPicHeightInLumaSamples();
  b.segStart = slice->getSliceSegmentCurStartCUAddr();
  b.runMask = hm_fast_col_part(~0u, (int)pic->getMinCUWidth());      // partition index -> first partition of its 16x16 run
  const int n = sym->getNumberOfCUsInFrame();
  ix.ctu.reserve(n);
  for (int a = 0; a < n; a++) b.addCtu(sym->getCU(a));
  ix.valid = true;
  return true;
}

// QUIRK (libHMDecoder.cpp:581-598): the "energy" of a TU of n levels is the sum of squares of the FIRST n levels of the CTU's
// coefficient buffer, not of the TU's own; ENERGY_CB is tested twice there, so ENERGY_CR never computes anything.
// In this build HM's whole-CTU zero fill is skipped (hm_fast.cpp): only partitions whose own leaf TU is coded hold defined levels
// (cbf bit at the leaf's transform depth; 4:2:2 chroma one level deeper, the two square halves carry separate flags), every other
// partition counts as the zeros stock HM would have left there.
void addEnergies(HmInternalsCache& ix)
{
  TComPicSym* sym = ix.pic->getPicSym();
  int levels = 0;
  while (levels < MAX_TU_LEVELS && (ix.ctuW >> levels) >= 4 && (ix.ctuH >> levels) >= 4) levels++;
  for (size_t a = 0; a < ix.ctu.size(); a++)
  {
    TComDataCU* ctu = sym->getCU((UInt)a);
    for (int k = 0; k < 2; k++)
    {
      const ComponentID c = k ? COMPONENT_Cb : COMPONENT_Y;
      const TCoeff* co = ctu->getCoeff(c);
      const UChar* cbf = ctu->getCbf(c);
      const UChar* trIdx = ctu->getTransformIdx();
      const int unit = (int)(g_uiMaxCUWidth >> g_uiMaxCUDepth);                                                  // HM's partition: 4 samples, 8 / 16 / 32 with a larger smallest transform block
      const int perPart = (unit * unit) >> (ctu->getPic()->getComponentScaleX(c) + ctu->getPic()->getComponentScaleY(c));   // levels per partition
      const UInt deeper = (c != COMPONENT_Y && ctu->getPic()->getChromaFormat() == CHROMA_422) ? 1 : 0;
      int64_t e = 0;
      int done = 0;                                        // levels [0, done) are in e
      for (int lv = levels - 1; lv >= (int)ix.ctu[a].topLevel; lv--)   // smallest blocks first: n grows, the sum runs on; no block of the CTU is larger than topLevel's
      {
        const int w = ix.ctuW >> lv, h = ix.ctuH >> lv;
        const int n = k ? (w / 2) * (h / 2) : w * h;
        for (int p = done / perPart; p * perPart < n; p++)
        {
          if ((p & 7) == 0 && (p + 8) * perPart <= n && co && cbf)
          {
            uint64_t eight; memcpy(&eight, cbf + p, 8);    // eight uncoded partitions in a row (the common case) are skipped with one test
            if (!eight) { p += 7; continue; }
          }
          const int lo = std::max(done, p * perPart), hi = std::min(n, (p + 1) * perPart);
          if (co && cbf && ((cbf[p] >> (trIdx[p] + deeper)) & 1))
            for (int i = lo; i < hi; i++) e += (int64_t)(co[i] * co[i]);
        }
        done = std::max(done, n);
        ix.ctu[a].energy[k][lv] = e > MAX_INT ? MAX_INT : (int)e;
      }
    }
  }
  ix.energies = true;
}

// ---------------------------------------------------------------- the 24 sweeps ----------------------------------------------------------------

typedef std::vector<libHMDec_BlockValue> Out;

// QUIRK: value / value2 are left uninitialised by the reference when a type does not set them; they are zero here
inline libHMDec_BlockValue* put(libHMDec_BlockValue* o, int x, int y, int w, int h, int v, int v2 = 0)
{
  o->x = (unsigned short)x; o->y = (unsigned short)y; o->w = (unsigned short)w; o->h = (unsigned short)h; o->value = v; o->value2 = v2;
  return o + 1;
}

void sweepCus(const HmInternalsCache& ix, Out& out, libHMDec_info_type type)
{
  out.resize(ix.cu.size());
  libHMDec_BlockValue* o = out.empty() ? NULL : &out[0];
  libHMDec_BlockValue* const o0 = o;
  size_t i = 0;
  for (size_t a = 0; a < ix.ctu.size(); a++)
  {
    const size_t end = ix.ctu[a].cuEnd;
    if (type == LIBHMDEC_CU_TRQ_BYPASS && !ix.ctu[a].bypassOn) { i = end; continue; }
    for (; i < end; i++)
    {
      const CuRec& c = ix.cu[i];
      switch (type)
      {
        case LIBHMDEC_CU_PREDICTION_MODE:   o = put(o, c.x, c.y, c.size, c.size, c.pred); break;
        case LIBHMDEC_CU_TRQ_BYPASS:        o = put(o, c.x, c.y, c.size, c.size, c.bypass); break;
        case LIBHMDEC_CU_SKIP_FLAG:         o = put(o, c.x, c.y, c.size, c.size, c.skip); break;
        case LIBHMDEC_CU_PART_MODE:         o = put(o, c.x, c.y, c.size, c.size, c.part); break;
        case LIBHMDEC_CU_INTRA_MODE_LUMA:   if (c.pred == MODE_INTRA) o = put(o, c.x, c.y, c.size, c.size, c.dirLuma); break;
        case LIBHMDEC_CU_INTRA_MODE_CHROMA: if (c.pred == MODE_INTRA) o = put(o, c.x, c.y, c.size, c.size, c.dirChroma); break;
        case LIBHMDEC_CU_ROOT_CBF:          if (c.pred != MODE_INTER) o = put(o, c.x, c.y, c.size, c.size, c.rootCbf); break;   // QUIRK: inverted w.r.t. its documentation
        default: break;
      }
    }
  }
  out.resize(o - o0);
}

void sweepPus(const HmInternalsCache& ix, Out& out, libHMDec_info_type type)
{
  out.resize(ix.pu.size());
  libHMDec_BlockValue* o = out.empty() ? NULL : &out[0];
  for (size_t i = 0; i < ix.pu.size(); i++)
  {
    const PuRec& p = ix.pu[i];
    switch (type)
    {
      case LIBHMDEC_PU_MERGE_FLAG:        o = put(o, p.x, p.y, p.w, p.h, p.merge); break;
      case LIBHMDEC_PU_UNI_BI_PREDICTION: o = put(o, p.x, p.y, p.w, p.h, p.dir); break;
      case LIBHMDEC_PU_REFERENCE_POC_0:   o = put(o, p.x, p.y, p.w, p.h, p.ref[0]); break;     // QUIRK: a reference *index*, not a POC
      case LIBHMDEC_PU_MV_0:              o = put(o, p.x, p.y, p.w, p.h, p.mv[0][0], p.mv[0][1]); break;
      case LIBHMDEC_PU_REFERENCE_POC_1:   o = put(o, p.x, p.y, p.w, p.h, p.dir == 2 ? p.ref[1] : 0); break;   // QUIRK: only set for list-1-only PUs
      case LIBHMDEC_PU_MV_1:              o = p.dir == 2 ? put(o, p.x, p.y, p.w, p.h, p.mv[1][0], p.mv[1][1]) : put(o, p.x, p.y, p.w, p.h, 0); break;
      default: break;
    }
  }
}

void sweepTus(const HmInternalsCache& ix, Out& out, libHMDec_info_type type)
{
  out.resize(ix.tu.size());
  libHMDec_BlockValue* o = out.empty() ? NULL : &out[0];
  libHMDec_BlockValue* const o0 = o;
  const bool tskipType = type == LIBHMDEC_TU_COEFF_TR_SKIP_Y || type == LIBHMDEC_TU_COEFF_TR_SKIP_Cb || type == LIBHMDEC_TU_COEFF_TR_SKIP_Cr;
  size_t i = 0;
  for (size_t a = 0; a < ix.ctu.size(); a++)
  {
    const CtuRec& r = ix.ctu[a];
    const size_t end = r.tuEnd;
    if (tskipType && r.tskipOn) { i = end; continue; }       // QUIRK: inverted condition in the reference (libHMDecoder.cpp:689)
    for (; i < end; i++)
    {
      const TuRec& t = ix.tu[i];
      int v = 0;
      switch (type)
      {
        case LIBHMDEC_TU_CBF_Y:            v = t.cbf & 1; break;
        case LIBHMDEC_TU_CBF_CB:           v = (t.cbf >> 1) & 1; break;
        case LIBHMDEC_TU_CBF_CR:           v = (t.cbf >> 2) & 1; break;
        case LIBHMDEC_TU_COEFF_TR_SKIP_Y:  v = t.tskip & 1; break;
        case LIBHMDEC_TU_COEFF_TR_SKIP_Cb: v = (t.tskip >> 1) & 1; break;
        case LIBHMDEC_TU_COEFF_TR_SKIP_Cr: v = (t.tskip >> 2) & 1; break;
        case LIBHMDEC_TU_COEFF_ENERGY_Y:   v = r.energy[0][t.level]; break;
        case LIBHMDEC_TU_COEFF_ENERGY_CB:  v = r.energy[1][t.level]; break;
        default: break;                                        // ENERGY_CR: geometry only (see addEnergies)
      }
      o = put(o, t.x, t.y, t.w, t.h, v);
    }
  }
  out.resize(o - o0);
}

} // namespace

HmInternalsCache* hm_internals_cache_new() { return new HmInternalsCache; }
void hm_internals_cache_free(HmInternalsCache* c) { delete c; }
void hm_internals_cache_invalidate(HmInternalsCache* c) { if (c) c->valid = false; }

std::vector<libHMDec_BlockValue>* hm_collect_internals(HmInternalsCache* cache, std::vector<libHMDec_BlockValue>& out, TComPic* pic, libHMDec_info_type type)
{
  HmInternalsCache& ix = *cache;
  if (!ix.valid || ix.pic != pic) { ix.valid = false; if (!buildIndex(ix, pic)) return NULL; }
  if (type == LIBHMDEC_CTU_SLICE_INDEX)
  {
    out.resize(ix.ctu.size());
    libHMDec_BlockValue* o = out.empty() ? NULL : &out[0];
    // QUIRK: every CTU reports the index of the picture's last slice (libHMDecoder.cpp:697)
    for (size_t a = 0; a < ix.ctu.size(); a++) o = put(o, ix.ctu[a].x, ix.ctu[a].y, ix.ctuW, ix.ctuH, ix.sliceIdx);
  }
  else if (type >= LIBHMDEC_CU_PREDICTION_MODE && type <= LIBHMDEC_CU_ROOT_CBF) sweepCus(ix, out, type);
  // QUIRK: LIBHMDEC_PU_MERGE_INDEX is missing from the reference's dispatch lists (libHMDecoder.cpp:663) and yields nothing
  else if (type == LIBHMDEC_PU_MERGE_FLAG || (type >= LIBHMDEC_PU_UNI_BI_PREDICTION && type <= LIBHMDEC_PU_MV_1)) sweepPus(ix, out, type);
  else if (type >= LIBHMDEC_TU_CBF_Y && type <= LIBHMDEC_TU_COEFF_ENERGY_CR)
  {
    if ((type == LIBHMDEC_TU_COEFF_ENERGY_Y || type == LIBHMDEC_TU_COEFF_ENERGY_CB) && !ix.energies) addEnergies(ix);
    sweepTus(ix, out, type);
  }
  return &out;
}
