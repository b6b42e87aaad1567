// internals.cpp — libHMDEC_get_internal_info: per-block coding decisions of a decoded picture
// (reference behaviour: source/App/libHMDecoder/libHMDecoder.cpp:451-715).  Pure host data; the GPU
// path does not touch it.  Observable quirks of the reference are kept on purpose and marked QUIRK.
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

namespace {

typedef std::vector<libHMDec_BlockValue> Out;

struct Geometry { int x, y, w, h; };

bool isCuType(libHMDec_info_type t) { return t >= LIBHMDEC_CU_PREDICTION_MODE && t <= LIBHMDEC_CU_ROOT_CBF; }
bool isPuType(libHMDec_info_type t)
{
  // QUIRK: LIBHMDEC_PU_MERGE_INDEX is missing from the reference's PU dispatch list (libHMDecoder.cpp:663) and yields nothing
  return t == LIBHMDEC_PU_MERGE_FLAG || t == LIBHMDEC_PU_UNI_BI_PREDICTION || t == LIBHMDEC_PU_REFERENCE_POC_0 ||
         t == LIBHMDEC_PU_MV_0 || t == LIBHMDEC_PU_REFERENCE_POC_1 || t == LIBHMDEC_PU_MV_1;
}
bool isTuType(libHMDec_info_type t) { return t >= LIBHMDEC_TU_CBF_Y && t <= LIBHMDEC_TU_COEFF_ENERGY_CR; }

libHMDec_BlockValue block(const Geometry& g)
{
  // QUIRK: value/value2 are left uninitialised by the reference when a type does not set them; we zero them
  libHMDec_BlockValue b;
  b.x = (unsigned short)g.x; b.y = (unsigned short)g.y; b.w = (unsigned short)g.w; b.h = (unsigned short)g.h;
  b.value = 0; b.value2 = 0;
  return b;
}

void puEntries(Out& out, TComDataCU* ctu, UInt part, UInt depth, libHMDec_info_type type)
{
  const PartSize ps = ctu->getPartitionSize(part);
  const int n = ps == SIZE_2Nx2N ? 1 : (ps == SIZE_NxN ? 4 : 2);
  const UInt step = (g_auiPUOffset[UInt(ps)] << ((ctu->getSlice()->getSPS()->getMaxCUDepth() - depth) << 1)) >> 4;
  const int S = g_uiMaxCUWidth >> depth, H = S >> 1, Q = S >> 2;
  const int cx = ctu->getCUPelX() + g_auiRasterToPelX[g_auiZscanToRaster[part]];
  const int cy = ctu->getCUPelY() + g_auiRasterToPelY[g_auiZscanToRaster[part]];
  UInt sub = part;
  for (int i = 0; i < n; i++, sub += step)
  {
    Geometry g = {cx, cy, S, S};
    switch (ps)
    {
      case SIZE_2NxN:  g.h = H; g.y += i ? H : 0; break;
      case SIZE_Nx2N:  g.w = H; g.x += i ? H : 0; break;
      case SIZE_NxN:   g.w = g.h = H; g.x += (i & 1) ? H : 0; g.y += (i >> 1) ? H : 0; break;
      case SIZE_2NxnU: g.h = i ? Q + H : Q; g.y += i ? Q : 0; break;
      case SIZE_2NxnD: g.h = i ? Q : Q + H; g.y += i ? Q + H : 0; break;
      case SIZE_nLx2N: g.w = i ? Q + H : Q; g.x += i ? Q : 0; break;
      case SIZE_nRx2N: g.w = i ? Q : Q + H; g.x += i ? Q + H : 0; break;
      default: break;
    }
    libHMDec_BlockValue b = block(g);
    TComCUMvField* f0 = ctu->getCUMvField(REF_PIC_LIST_0);
    TComCUMvField* f1 = ctu->getCUMvField(REF_PIC_LIST_1);
    switch (type)
    {
      case LIBHMDEC_PU_MERGE_FLAG:        b.value = ctu->getMergeFlag(sub) ? 1 : 0; break;
      case LIBHMDEC_PU_UNI_BI_PREDICTION: b.value = (int)ctu->getInterDir(sub); break;
      case LIBHMDEC_PU_REFERENCE_POC_0:   b.value = f0->getRefIdx(sub); break;      // QUIRK: a reference *index*, not a POC
      case LIBHMDEC_PU_MV_0:              b.value = f0->getMv(sub).getHor(); b.value2 = f0->getMv(sub).getVer(); break;
      case LIBHMDEC_PU_REFERENCE_POC_1:   if (ctu->getInterDir(sub) == 2) b.value = f1->getRefIdx(sub); break;
      case LIBHMDEC_PU_MV_1:              if (ctu->getInterDir(sub) == 2) { b.value = f1->getMv(sub).getHor(); b.value2 = f1->getMv(sub).getVer(); } break;
      default: break;
    }
    out.push_back(b);
  }
}

void tuEntries(Out& out, TComDataCU* ctu, UInt part, UInt depth, UInt trDepth, libHMDec_info_type type)
{
  if (trDepth < ctu->getTransformIdx(part))
  {
    const UInt q = ctu->getTotalNumPart() >> ((depth + trDepth + 1) << 1);
    for (int i = 0; i < 4; i++) tuEntries(out, ctu, part + i * q, depth, trDepth + 1, type);
    // QUIRK: no return here in the reference — a split node reports an entry for itself after its children
  }
  Geometry g = { (int)(ctu->getCUPelX() + g_auiRasterToPelX[g_auiZscanToRaster[part]]),
                 (int)(ctu->getCUPelY() + g_auiRasterToPelY[g_auiZscanToRaster[part]]),
                 (int)(g_uiMaxCUWidth >> (depth + trDepth)), (int)(g_uiMaxCUHeight >> (depth + trDepth)) };
  libHMDec_BlockValue b = block(g);
  switch (type)
  {
    case LIBHMDEC_TU_CBF_Y:            b.value = ctu->getCbf(part, COMPONENT_Y,  trDepth) ? 1 : 0; break;
    case LIBHMDEC_TU_CBF_CB:           b.value = ctu->getCbf(part, COMPONENT_Cb, trDepth) ? 1 : 0; break;
    case LIBHMDEC_TU_CBF_CR:           b.value = ctu->getCbf(part, COMPONENT_Cr, trDepth) ? 1 : 0; break;
    case LIBHMDEC_TU_COEFF_TR_SKIP_Y:  b.value = ctu->getTransformSkip(part, COMPONENT_Y)  ? 1 : 0; break;
    case LIBHMDEC_TU_COEFF_TR_SKIP_Cb: b.value = ctu->getTransformSkip(part, COMPONENT_Cb) ? 1 : 0; break;
    case LIBHMDEC_TU_COEFF_TR_SKIP_Cr: b.value = ctu->getTransformSkip(part, COMPONENT_Cr) ? 1 : 0; break;
    case LIBHMDEC_TU_COEFF_ENERGY_Y:
    case LIBHMDEC_TU_COEFF_ENERGY_CB:
    {
      // QUIRK: the reference tests ENERGY_CB twice (libHMDecoder.cpp:581), so ENERGY_CR reports nothing; and it sums the
      // FIRST w*h (or w/2*h/2) levels of the CTU's coefficient buffer, not the TU's own.
      const ComponentID c = type == LIBHMDEC_TU_COEFF_ENERGY_Y ? COMPONENT_Y : COMPONENT_Cb;
      const int n = type == LIBHMDEC_TU_COEFF_ENERGY_Y ? g.w * g.h : (g.w / 2) * (g.h / 2);
      // In this build HM's whole-CTU zero fill is skipped (hm_fast.cpp): only blocks with coded levels hold defined
      // values, every other block counts as the zeros stock HM would have left there.
      const TCoeff* co = ctu->getCoeff(c);
      const int perPart = 16 >> (ctu->getPic()->getComponentScaleX(c) + ctu->getPic()->getComponentScaleY(c));   // levels per 4x4 partition
      int64_t e = 0;
      // a partition's levels are defined iff its own leaf TU is coded: cbf bit at the leaf's transform depth (4:2:2 chroma:
      // one level deeper, the two square halves of a TU carry separate flags)
      const UInt deeper = (c != COMPONENT_Y && ctu->getPic()->getChromaFormat() == CHROMA_422) ? 1 : 0;
      const UChar* cbf = ctu->getCbf(c);
      const UChar* trIdx = ctu->getTransformIdx();
      for (int p = 0; p * perPart < n; p++)                  // partition by partition: one flag test per 4x4 unit, not per level
      {
        if (((cbf[p] >> (trIdx[p] + deeper)) & 1) == 0) continue;
        const TCoeff* q = co + p * perPart;
        const int m = std::min(perPart, n - p * perPart);
        for (int i = 0; i < m; i++) e += (int64_t)(q[i] * q[i]);
      }
      b.value = e > MAX_INT ? MAX_INT : (int)e;
      break;
    }
    default: break;
  }
  out.push_back(b);
}

void cuWalk(Out& out, TComDataCU* ctu, UInt part, UInt depth, libHMDec_info_type type)
{
  TComPic* pic = ctu->getPic();
  TComSlice* slice = pic->getSlice(pic->getCurrSliceIdx());
  const UInt W = slice->getSPS()->getPicWidthInLumaSamples(), Hh = slice->getSPS()->getPicHeightInLumaSamples();
  const UInt lx = ctu->getCUPelX() + g_auiRasterToPelX[g_auiZscanToRaster[part]];
  const UInt ty = ctu->getCUPelY() + g_auiRasterToPelY[g_auiZscanToRaster[part]];
  const UInt rx = lx + (g_uiMaxCUWidth >> depth) - 1, by = ty + (g_uiMaxCUHeight >> depth) - 1;
  const UInt nParts = pic->getNumPartInCU() >> (depth << 1);
  const bool startInCU = ctu->getSCUAddr() + part + nParts > slice->getSliceSegmentCurStartCUAddr() && ctu->getSCUAddr() + part < slice->getSliceSegmentCurStartCUAddr();
  if ((depth < ctu->getDepth(part) && depth < g_uiMaxCUDepth - g_uiAddCUDepth) || startInCU || rx >= W || by >= Hh)
  {
    const UInt q = ctu->getTotalNumPart() >> ((depth + 1) << 1);
    UInt idx = part;
    for (int i = 0; i < 4; i++, idx += q)
    {
      const UInt qx = ctu->getCUPelX() + g_auiRasterToPelX[g_auiZscanToRaster[idx]];
      const UInt qy = ctu->getCUPelY() + g_auiRasterToPelY[g_auiZscanToRaster[idx]];
      if (qx < W && qy < Hh) cuWalk(out, ctu, idx, depth + 1, type);
    }
    return;
  }

  if (isCuType(type))
  {
    if (type == LIBHMDEC_CU_TRQ_BYPASS && !ctu->getSlice()->getPPS()->getTransquantBypassEnableFlag()) return;
    if ((type == LIBHMDEC_CU_INTRA_MODE_LUMA || type == LIBHMDEC_CU_INTRA_MODE_CHROMA) && !ctu->isIntra(part)) return;
    if (type == LIBHMDEC_CU_ROOT_CBF && ctu->isInter(part)) return;   // QUIRK: inverted w.r.t. its documentation
    Geometry g = {(int)lx, (int)ty, (int)(g_uiMaxCUWidth >> depth), (int)(g_uiMaxCUHeight >> depth)};
    libHMDec_BlockValue b = block(g);
    switch (type)
    {
      case LIBHMDEC_CU_PREDICTION_MODE:   b.value = (int)ctu->getPredictionMode(part); break;
      case LIBHMDEC_CU_TRQ_BYPASS:        b.value = ctu->getCUTransquantBypass(part) ? 1 : 0; break;
      case LIBHMDEC_CU_SKIP_FLAG:         b.value = ctu->isSkipped(part) ? 1 : 0; break;
      case LIBHMDEC_CU_PART_MODE:         b.value = (int)ctu->getPartitionSize(part); break;
      case LIBHMDEC_CU_INTRA_MODE_LUMA:   b.value = (int)ctu->getIntraDir(CHANNEL_TYPE_LUMA, part); break;
      case LIBHMDEC_CU_INTRA_MODE_CHROMA: b.value = (int)ctu->getIntraDir(CHANNEL_TYPE_CHROMA, part); break;
      case LIBHMDEC_CU_ROOT_CBF:          b.value = (int)ctu->getQtRootCbf(part); break;
      default: break;
    }
    out.push_back(b);
  }
  else if (isPuType(type)) { if (ctu->isInter(part)) puEntries(out, ctu, part, depth, type); }
  else if (isTuType(type) && type != LIBHMDEC_TU_COEFF_ENERGY_CR) tuEntries(out, ctu, part, depth, 0, type);
  else if (type == LIBHMDEC_TU_COEFF_ENERGY_CR) tuEntries(out, ctu, part, depth, 0, type);
}

} // namespace

std::vector<libHMDec_BlockValue>* hm_collect_internals(std::vector<libHMDec_BlockValue>& out, TComPic* pic, libHMDec_info_type type)
{
  hm_fast_ensure_motion_compressed(pic);     // non-reference pictures postpone compressMotion until somebody looks
  TComPicSym* sym = pic->getPicSym();
  if (!sym) return NULL;
  const int n = sym->getNumberOfCUsInFrame();
  for (int a = 0; a < n; a++)
  {
    TComDataCU* ctu = sym->getCU(a);
    const bool tskipType = type == LIBHMDEC_TU_COEFF_TR_SKIP_Y || type == LIBHMDEC_TU_COEFF_TR_SKIP_Cb || type == LIBHMDEC_TU_COEFF_TR_SKIP_Cr;
    if (tskipType && ctu->getSlice()->getPPS()->getUseTransformSkip()) continue;   // QUIRK: inverted condition in the reference (libHMDecoder.cpp:689)
    if (type == LIBHMDEC_CTU_SLICE_INDEX)
    {
      Geometry g = {(int)ctu->getCUPelX(), (int)ctu->getCUPelY(), (int)g_uiMaxCUWidth, (int)g_uiMaxCUHeight};
      libHMDec_BlockValue b = block(g);
      b.value = (int)ctu->getPic()->getCurrSliceIdx();
      out.push_back(b);
    }
    else cuWalk(out, ctu, 0, 0, type);
  }
  return &out;
}
