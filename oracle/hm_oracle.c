/* hm_oracle.c — CPU restatement (plain C, scalar, single thread) of HM-16.0's picture
 * reconstruction, driven by the same flat records the GPU engine consumes (include/hmr_records.h).
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load it; the product path (libhm_b200/, frontend/) never links or calls it.
 *
 * Parity pin: tests/test_oracle_golden.py runs orc_reconstruct_frame() over the records of every
 * stream in tests/golden/ and compares the per-stage picture MD5s with the ones produced by HM
 * itself (frontend/dump_sink.cpp, GOLD sections) — whose final stage additionally equals the
 * encoder-embedded SEI decoded-picture hash checked by the unmodified TAppDecoder.
 *
 * Each function cites the reference routine it restates (paths under /root/reference/source/Lib).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "hmr_records.h"

#define ORC_MAX_TU 32

typedef struct orc_pic {
  int16_t* plane[3];
  int      width[3], height[3], stride[3];
} orc_pic;

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int csx_of(int fmt, int comp) { return (comp && (fmt == HMR_CHROMA_420 || fmt == HMR_CHROMA_422)) ? 1 : 0; }
static inline int csy_of(int fmt, int comp) { return (comp && fmt == HMR_CHROMA_420) ? 1 : 0; }

/* ---------------------------------------------------------------------------------------------
 * Transform matrices.  HEVC core transform (TComRom.cpp:335-484): every entry of the 32-point matrix
 * is +-A[j], j = angle index in units of pi/64; N-point matrices are rows k*(32/N), columns 0..N-1. */
static const int16_t kCosTab[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                     61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
static const int16_t kDst4[4][4] = { {29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29} };
static int16_t g_T32[32][32];
static int g_tables_ready = 0;

static void build_tables(void)
{
  if (g_tables_ready) return;
  for (int k = 0; k < 32; k++)
    for (int n = 0; n < 32; n++)
    {
      if (k == 0) { g_T32[k][n] = 64; continue; }
      int m = ((2 * n + 1) * k) & 127;      /* angle, period 128 */
      if (m > 64) m = 128 - m;              /* cos(2pi - x) = cos x */
      g_T32[k][n] = (int16_t)(m > 32 ? -kCosTab[64 - m] : kCosTab[m]);
    }
  g_tables_ready = 1;
}

void orc_get_dct_matrix(int n, int16_t* out /* n*n */)
{
  build_tables();
  for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) out[k * n + j] = g_T32[k * (32 / n)][j];
}

/* one separable stage of xITrMxN (TComTrQuant.cpp:894-948; butterflies :468-828 are an exact factorisation of this
 * matrix product): dst[j*N + k] = clip((sum_n M[n][k] * src[n*N + j] + rnd) >> shift)   — output transposed */
static void inv_stage(const int32_t* src, int32_t* dst, int N, int use_dst, int shift, int lo, int hi)
{
  const int step = 32 / N;
  const int rnd = shift > 0 ? (1 << (shift - 1)) : 0;
  for (int j = 0; j < N; j++)
    for (int k = 0; k < N; k++)
    {
      int32_t acc = 0;
      for (int n = 0; n < N; n++)
      {
        const int m = use_dst ? kDst4[n][k] : g_T32[n * step][k];
        acc += m * src[n * N + j];
      }
      dst[j * N + k] = clip3(lo, hi, (acc + rnd) >> shift);
    }
}

/* Residual of one TU: TComTrQuant::invTransformNxN (TComTrQuant.cpp:1423-1548) = xDeQuant (:1203-1313, flat scaling)
 * + xIT / xITransformSkip (:1836-1866, :1920-1959) or bypass copy (:1475-1487) + invRdpcmNxN (:1737-1792). */
/* `scaling` = hmr_frame_desc.scaling or NULL: xDeQuant's scaling-list branch (TComTrQuant.cpp:1230-1276) */
void orc_tu_residual(const hmr_tu* t, const int16_t* level, int bit_depth, const uint8_t* scaling, int16_t* resi /* N*N, stride N */)
{
  static const int invq[6] = { 40, 45, 51, 57, 64, 72 };             /* g_invQuantScales, TComRom.cpp:326 */
  const int log2n = t->log2_size, N = 1 << log2n, NN = N * N;
  build_tables();
  if (!(t->flags & HMR_TU_CODED)) { memset(resi, 0, sizeof(int16_t) * NN); }
  else if (t->flags & HMR_TU_BYPASS)
  {
    for (int i = 0; i < NN; i++) resi[i] = level[(t->flags & HMR_TU_ROTATE) ? NN - 1 - i : i];
  }
  else
  {
    int32_t coef[ORC_MAX_TU * ORC_MAX_TU], tmp[ORC_MAX_TU * ORC_MAX_TU], blk[ORC_MAX_TU * ORC_MAX_TU];
    const int per = t->qp / 6, rem = t->qp % 6, scale = invq[rem];
    const int tr_shift = 15 - bit_depth - log2n;                       /* getTransformShift, TComChromaFormat.h:166 */
    /* getUseScalingList (TComTrQuant.h): enabled, and not a transform-skipped block other than 4x4 */
    const int use_sl = scaling != NULL && (!(t->flags & HMR_TU_TSKIP) || N == 4);
    const uint8_t* m = use_sl ? scaling + HMR_SCALING_OFFSET(log2n - 2) + ((t->flags & HMR_TU_INTRA) ? 0 : 3) * NN + t->comp * NN : NULL;
    const int rshift = 6 - (tr_shift + per) + (use_sl ? 4 : 0);        /* IQUANT_SHIFT = 6, LOG2_SCALING_LIST_NEUTRAL_VALUE = 4 */
    int in_bits = 32 + rshift - (use_sl ? 15 : 7); if (in_bits > 16) in_bits = 16;   /* targetInputBitDepth, :1243 / :1284 */
    const int in_min = -(1 << (in_bits - 1)), in_max = (1 << (in_bits - 1)) - 1;
    for (int i = 0; i < NN; i++)
    {
      const int q = clip3(in_min, in_max, level[i]);
      const int sc = use_sl ? scale * m[i] : scale;
      const int32_t v = rshift > 0 ? (q * sc + (1 << (rshift - 1))) >> rshift : (int32_t)((uint32_t)(q * sc) << (-rshift));
      coef[i] = clip3(-32768, 32767, v);
    }
    if (t->flags & HMR_TU_TSKIP)
    {
      const int rot = (t->flags & HMR_TU_ROTATE) != 0;
      if (tr_shift >= 0)
      {
        const int off = tr_shift == 0 ? 0 : (1 << (tr_shift - 1));
        for (int i = 0; i < NN; i++) resi[i] = (int16_t)((coef[rot ? NN - 1 - i : i] + off) >> tr_shift);
      }
      else for (int i = 0; i < NN; i++) resi[i] = (int16_t)(coef[rot ? NN - 1 - i : i] << (-tr_shift));
    }
    else
    {
      const int use_dst = (t->flags & HMR_TU_DST) != 0;
      inv_stage(coef, tmp, N, use_dst, 7, -32768, 32767);              /* shift_1st = 6 + 1 */
      inv_stage(tmp, blk, N, use_dst, 20 - bit_depth, -32768, 32767);  /* shift_2nd = (6 + 15 - 1) - bitDepth; Pel clip */
      for (int i = 0; i < NN; i++) resi[i] = (int16_t)blk[i];
    }
  }
  if (t->flags & HMR_TU_RDPCM_V)
    for (int y = 1; y < N; y++) for (int x = 0; x < N; x++) resi[y * N + x] = (int16_t)(resi[y * N + x] + resi[(y - 1) * N + x]);
  else if (t->flags & HMR_TU_RDPCM_H)
    for (int y = 0; y < N; y++) for (int x = 1; x < N; x++) resi[y * N + x] = (int16_t)(resi[y * N + x] + resi[y * N + x - 1]);
}

/* cross-component prediction, decoder direction (TComTrQuant.cpp:3294-3335, reverse = true) */
static void ccp_apply(int16_t* resi_c, const int16_t* resi_l, int n_samples, int alpha, int diff_bd)
{
  for (int i = 0; i < n_samples; i++)
  {
    const int l = diff_bd >= 0 ? (resi_l[i] >> diff_bd) : (resi_l[i] << (-diff_bd));
    resi_c[i] = (int16_t)(resi_c[i] + ((alpha * l) >> 3));
  }
}

/* ---------------------------------------------------------------------------------------------
 * Inter prediction: TComPrediction::xPredInterBlk (TComPrediction.cpp:660-698),
 * TComInterpolationFilter::filter (TComInterpolationFilter.cpp:166-251), filterCopy (:94-148), TComYuv::addAvg (TComYuv.cpp:336-391).
 * Reference pictures are read with coordinate clamping == HM's replicated border (TComPicYuv.cpp:173-217); clipMv keeps
 * every access inside that border. */
static const int8_t kLumaTaps[4][8] = { {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1} };
static const int8_t kChromaTaps[8][4] = { {0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4}, {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2} };

static inline int ref_at(const orc_pic* r, int c, int x, int y)
{
  x = clip3(0, r->width[c] - 1, x); y = clip3(0, r->height[c] - 1, y);
  return r->plane[c][(size_t)y * r->stride[c] + x];
}

/* Predict one block of one component from one list.  bi = 1: 14-bit intermediate (int16) output, bi = 0: final samples. */
static void mc_block(const orc_pic* ref, int c, int fmt, int bd, int x0, int y0, int w, int h, int mvx, int mvy, int bi, int16_t* out /* stride w */)
{
  const int sx = 2 + csx_of(fmt, c), sy = 2 + csy_of(fmt, c);
  const int ix = x0 + (mvx >> sx), iy = y0 + (mvy >> sy);
  const int fx = mvx & ((1 << sx) - 1), fy = mvy & ((1 << sy) - 1);
  const int ntaps = c ? 4 : 8, half = ntaps / 2 - 1;
  int8_t tx[8], ty[8];
  for (int i = 0; i < ntaps; i++)
  {
    tx[i] = c ? kChromaTaps[fx << (1 - csx_of(fmt, c))][i] : kLumaTaps[fx][i];
    ty[i] = c ? kChromaTaps[fy << (1 - csy_of(fmt, c))][i] : kLumaTaps[fy][i];
  }
  const int headroom = (14 - bd) > 2 ? (14 - bd) : 2;
  const int maxv = (1 << bd) - 1;
  if (fx == 0 && fy == 0)
  {
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
    {
      const int s = ref_at(ref, c, ix + x, iy + y);
      out[y * w + x] = bi ? (int16_t)((int16_t)(s << headroom) - (int16_t)8192) : (int16_t)s;
    }
    return;
  }
  if (fy == 0 || fx == 0)
  {
    /* single 1-D pass: isFirst = 1, isLast = !bi */
    const int8_t* taps = fy == 0 ? tx : ty;
    const int shift = bi ? 6 - headroom : 6;
    const int offset = bi ? -(8192 << shift) : (1 << (shift - 1));
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
    {
      int sum = 0;
      for (int i = 0; i < ntaps; i++)
        sum += taps[i] * (fy == 0 ? ref_at(ref, c, ix + x + i - half, iy + y) : ref_at(ref, c, ix + x, iy + y + i - half));
      int16_t val = (int16_t)((sum + offset) >> shift);
      if (!bi) { if (val < 0) val = 0; if (val > maxv) val = (int16_t)maxv; }
      out[y * w + x] = val;
    }
    return;
  }
  /* separable: horizontal first (isFirst, !isLast) over h + ntaps - 1 rows, then vertical (!isFirst, isLast = !bi) */
  {
    int16_t tmp[(64 + 7) * 64];
    const int rows = h + ntaps - 1;
    const int s1 = 6 - headroom, o1 = -(8192 << s1);
    for (int y = 0; y < rows; y++) for (int x = 0; x < w; x++)
    {
      int sum = 0;
      for (int i = 0; i < ntaps; i++) sum += tx[i] * ref_at(ref, c, ix + x + i - half, iy + y - half);
      tmp[y * w + x] = (int16_t)((sum + o1) >> s1);
    }
    const int s2 = bi ? 6 : 6 + headroom;
    const int o2 = bi ? 0 : (1 << (s2 - 1)) + (8192 << 6);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
    {
      int sum = 0;
      for (int i = 0; i < ntaps; i++) sum += ty[i] * tmp[(y + i) * w + x];
      int16_t val = (int16_t)((sum + o2) >> s2);
      if (!bi) { if (val < 0) val = 0; if (val > maxv) val = (int16_t)maxv; }
      out[y * w + x] = val;
    }
  }
}

/* One PU, all components, into dst: motionCompensation / xPredInterUni / xPredInterBi (TComPrediction.cpp:514-644) */
/* wp = hmr_frame_desc.wp or NULL; refidx = pu_refidx[pu].  With explicit weighted prediction both the uni and the bi case go
 * through the 14-bit intermediates (xPredInterUni(..., bi = true), TComPrediction.cpp:596-644) and then
 * TComWeightPrediction::addWeightUni / addWeightBi (TComWeightPrediction.cpp:44-53,75-196). */
void orc_predict_pu(const hmr_frame_hdr* h, const hmr_pu* p, const hmr_wp* wp, int refidx, const orc_pic* dpb, orc_pic* dst)
{
  int16_t a[64 * 64], b[64 * 64];
  const int fmt = h->chroma_format;
  for (int c = 0; c < (fmt == HMR_CHROMA_400 ? 1 : 3); c++)   /* getNumberValidComponents (TComPrediction.cpp:520) */
  {
    const int cx = csx_of(fmt, c), cy = csy_of(fmt, c);
    const int bd = c ? h->bit_depth_chroma : h->bit_depth_luma;
    const int x0 = p->x >> cx, y0 = p->y >> cy, w = p->w >> cx, hh = p->h >> cy;
    const int bi = (p->lists == (HMR_PU_L0 | HMR_PU_L1));
    const int inter = bi || wp != NULL;                      /* keep the 14-bit intermediate */
    if (p->lists & HMR_PU_L0) mc_block(&dpb[p->slots & 15], c, fmt, bd, x0, y0, w, hh, p->mv[0][0], p->mv[0][1], inter, a);
    if (p->lists & HMR_PU_L1) mc_block(&dpb[p->slots >> 4], c, fmt, bd, x0, y0, w, hh, p->mv[1][0], p->mv[1][1], inter, bi ? b : a);
    const int headroom = (14 - bd) > 2 ? (14 - bd) : 2;
    const int sh = headroom + 1, off = (1 << (sh - 1)) + 2 * 8192, maxv = (1 << bd) - 1;
    const hmr_wp* w0 = wp ? &wp[(0 * 16 + (refidx & 15)) * 3 + c] : NULL;
    const hmr_wp* w1 = wp ? &wp[(1 * 16 + (refidx >> 4)) * 3 + c] : NULL;
    for (int y = 0; y < hh; y++) for (int x = 0; x < w; x++)
    {
      int v;
      if (!wp) v = bi ? clip3(0, maxv, (a[y * w + x] + b[y * w + x] + off) >> sh) : a[y * w + x];
      else if (bi)
      {
        const int shift = w0->log2_denom + 1 + headroom, round = 1 << (shift - 1), offset = w0->offset + w1->offset;
        v = clip3(0, maxv, (w0->weight * (a[y * w + x] + 8192) + w1->weight * (b[y * w + x] + 8192) + round + (offset << (shift - 1))) >> shift);
      }
      else
      {
        const hmr_wp* q = (p->lists & HMR_PU_L0) ? w0 : w1;
        const int shift = q->log2_denom + headroom, round = shift > 0 ? 1 << (shift - 1) : 0;
        v = clip3(0, maxv, ((q->weight * (a[y * w + x] + 8192) + round) >> shift) + q->offset);
      }
      dst->plane[c][(size_t)(y0 + y) * dst->stride[c] + x0 + x] = (int16_t)v;
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * Intra prediction of one TU: initAdiPatternChType + fillReferenceSamples (TComPattern.cpp:107-520),
 * predIntraAng / xPredIntraAng / xPredIntraPlanar / DC (TComPrediction.cpp:182-491, 746-835). */
void orc_intra_predict(const hmr_frame_hdr* h, const hmr_intra* r, const orc_pic* pic, int16_t* pred /* N*N stride N */)
{
  const int c = r->comp, fmt = h->chroma_format;
  const int bd = c ? h->bit_depth_chroma : h->bit_depth_luma;
  const int N = 1 << r->log2_size, N2 = 2 * N, L = 4 * N + 1;
  const int uw = 4 >> csx_of(fmt, c), uh = 4 >> csy_of(fmt, c);
  const int16_t* pl = pic->plane[c];
  const int st = pic->stride[c];
  const int x0 = r->x, y0 = r->y;
  int line[4 * ORC_MAX_TU + 1], flt[4 * ORC_MAX_TU + 1];
  uint8_t av[4 * ORC_MAX_TU + 1];
  if (r->mode == HMR_INTRA_MODE_PCM) { memset(pred, 0, sizeof(int16_t) * N * N); return; }   /* xReconPCM: the samples arrive as the residual */

  /* line[0] = bottom-most below-left ... line[2N-1] = left y=0, line[2N] = corner, line[2N+1+x] = above x */
  int any = 0;
  for (int i = 0; i < L; i++)
  {
    int ok, v = 0;
    if (i < N2)
    {
      const int y = N2 - 1 - i;                         /* 0..2N-1 below the top edge */
      const int unit = (y % N) / uh;
      ok = y < N ? (r->avail_left >> unit) & 1 : (r->avail_below_left >> unit) & 1;
      if (ok) v = pl[(size_t)(y0 + y) * st + x0 - 1];
    }
    else if (i == N2) { ok = (r->flags & HMR_INTRA_AVAIL_CORNER) != 0; if (ok) v = pl[(size_t)(y0 - 1) * st + x0 - 1]; }
    else
    {
      const int x = i - N2 - 1;
      const int unit = (x % N) / uw;
      ok = x < N ? (r->avail_above >> unit) & 1 : (r->avail_above_right >> unit) & 1;
      if (ok) v = pl[(size_t)(y0 - 1) * st + x0 + x];
    }
    av[i] = (uint8_t)ok; line[i] = v; any |= ok;
  }
  if (!any) for (int i = 0; i < L; i++) line[i] = 1 << (bd - 1);
  else
  {
    if (!av[0]) { int k = 1; while (!av[k]) k++; line[0] = line[k]; }
    for (int i = 1; i < L; i++) if (!av[i]) line[i] = line[i - 1];
  }

  const int* ref = line;
  if (r->flags & HMR_INTRA_FILTER_REFS)
  {
    const int bl = line[0], tl = line[N2], tr = line[4 * N];
    int strong = (r->flags & HMR_INTRA_LUMA_RULES) && (h->flags & HMR_FRM_STRONG_INTRA_SMOOTHING) && N >= 32;
    if (strong)
    {
      const int thr = 1 << (bd - 5);
      if (!(iabs(bl + tl - 2 * line[N]) < thr && iabs(tl + tr - 2 * line[3 * N]) < thr)) strong = 0;
    }
    flt[0] = line[0]; flt[4 * N] = line[4 * N];
    if (strong)
    {
      const int sh = r->log2_size + 1;
      for (int i = 1; i < N2; i++) flt[i] = ((N2 - i) * bl + i * tl + N) >> sh;
      flt[N2] = tl;
      for (int i = 1; i < N2; i++) flt[N2 + i] = ((N2 - i) * tl + i * tr + N) >> sh;
    }
    else for (int i = 1; i < 4 * N; i++) flt[i] = (line[i - 1] + 2 * line[i] + line[i + 1] + 2) >> 2;
    ref = flt;
  }
#define LEFT(y) ref[N2 - 1 - (y)]   /* y = -1 is the corner */
#define TOP(x)  ref[N2 + 1 + (x)]   /* x = -1 is the corner */
  const int mode = r->mode, maxv = (1 << bd) - 1;
  const int luma_rules = (r->flags & HMR_INTRA_LUMA_RULES) != 0;
  if (mode == 0)
  {
    const int sh = r->log2_size + 1;
    for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
      pred[y * N + x] = (int16_t)(((N - 1 - x) * LEFT(y) + (x + 1) * TOP(N) + (N - 1 - y) * TOP(x) + (y + 1) * LEFT(N) + N) >> sh);
  }
  else if (mode == 1)
  {
    int sum = 0;
    for (int i = 0; i < N; i++) sum += TOP(i) + LEFT(i);
    const int dc = (sum + N) / (N2);
    for (int i = 0; i < N * N; i++) pred[i] = (int16_t)dc;
    if (luma_rules && N <= 16)
    {
      pred[0] = (int16_t)((TOP(0) + LEFT(0) + 2 * dc + 2) >> 2);
      for (int x = 1; x < N; x++) pred[x] = (int16_t)((TOP(x) + 3 * dc + 2) >> 2);
      for (int y = 1; y < N; y++) pred[y * N] = (int16_t)((LEFT(y) + 3 * dc + 2) >> 2);
    }
  }
  else
  {
    static const int ang_tab[9] = { 0, 2, 5, 9, 13, 17, 21, 26, 32 };
    static const int inv_tab[9] = { 0, 4096, 1638, 910, 630, 482, 390, 315, 256 };
    const int ver = mode >= 18;
    const int am = ver ? mode - 26 : -(mode - 10);
    const int aa = iabs(am), angle = am < 0 ? -ang_tab[aa] : ang_tab[aa], inv = inv_tab[aa];
    int buf[3 * ORC_MAX_TU + 2];
    int* rm = buf + ORC_MAX_TU;                 /* rm[-N .. 2N] : main reference, rm[0] = corner */
    /* main = above for vertical modes, left for horizontal ones; side the other */
    if (angle < 0)
    {
      for (int i = 0; i <= N; i++) rm[i] = ver ? TOP(i - 1) : LEFT(i - 1);
      int acc = 128;
      const int last = (N * angle) >> 5;
      for (int k = -1; k > last; k--) { acc += inv; rm[k] = ver ? LEFT((acc >> 8) - 1) : TOP((acc >> 8) - 1); }
    }
    else for (int i = 0; i <= N2; i++) rm[i] = ver ? TOP(i - 1) : LEFT(i - 1);
    const int edge = luma_rules && N <= 16 && !(r->flags & HMR_INTRA_NO_EDGE_FLT);
    for (int yy = 0; yy < N; yy++)        /* yy runs along the prediction direction's minor axis */
    {
      const int pos = (yy + 1) * angle, di = pos >> 5, df = pos & 31;
      for (int xx = 0; xx < N; xx++)
      {
        int v;
        if (angle == 0)
        {
          v = rm[xx + 1];
          if (edge && xx == 0) v = clip3(0, maxv, v + (((ver ? LEFT(yy) : TOP(yy)) - ref[N2]) >> 1));
        }
        else if (df) v = ((32 - df) * rm[xx + di + 1] + df * rm[xx + di + 2] + 16) >> 5;
        else v = rm[xx + di + 1];
        if (ver) pred[yy * N + xx] = (int16_t)v; else pred[xx * N + yy] = (int16_t)v;
      }
    }
  }
#undef LEFT
#undef TOP
}

/* ---------------------------------------------------------------------------------------------
 * Deblocking: xEdgeFilterLuma / xEdgeFilterChroma / xPelFilterLuma / xPelFilterChroma (TComLoopFilter.cpp:540-922) */
static const uint8_t kTc[54] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24 };
static const uint8_t kBeta[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64 };
static const uint8_t kChromaQp420[58] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51 };

static void deblock_luma_segment(int16_t* p, int step /* across the edge */, int line /* along the edge */, int bs, int qp, int beta_off, int tc_off, int bd, int nofilt_p, int nofilt_q)
{
  const int scale = 1 << (bd - 8);
  const int tc = kTc[clip3(0, 53, qp + 2 * (bs - 1) + (tc_off << 1))] * scale;
  const int beta = kBeta[clip3(0, 51, qp + (beta_off << 1))] * scale;
  const int side_thr = (beta + (beta >> 1)) >> 3, thr_cut = tc * 10;
#define S(k, i) p[(k) * step + (i) * line]            /* k = -4..3 across the edge, i = line 0..3 */
  const int dp0 = iabs(S(-3, 0) - 2 * S(-2, 0) + S(-1, 0)), dq0 = iabs(S(0, 0) - 2 * S(1, 0) + S(2, 0));
  const int dp3 = iabs(S(-3, 3) - 2 * S(-2, 3) + S(-1, 3)), dq3 = iabs(S(0, 3) - 2 * S(1, 3) + S(2, 3));
  const int d0 = dp0 + dq0, d3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = d0 + d3;
  if (d >= beta) return;
  const int fp = dp < side_thr, fq = dq < side_thr;
  const int sw0 = (iabs(S(-4, 0) - S(-1, 0)) + iabs(S(3, 0) - S(0, 0)) < (beta >> 3)) && (2 * d0 < (beta >> 2)) && (iabs(S(-1, 0) - S(0, 0)) < ((tc * 5 + 1) >> 1));
  const int sw3 = (iabs(S(-4, 3) - S(-1, 3)) + iabs(S(3, 3) - S(0, 3)) < (beta >> 3)) && (2 * d3 < (beta >> 2)) && (iabs(S(-1, 3) - S(0, 3)) < ((tc * 5 + 1) >> 1));
  const int strong = sw0 && sw3, maxv = (1 << bd) - 1;
  for (int i = 0; i < 4; i++)
  {
    const int m0 = S(-4, i), m1 = S(-3, i), m2 = S(-2, i), m3 = S(-1, i), m4 = S(0, i), m5 = S(1, i), m6 = S(2, i), m7 = S(3, i);
    int n1 = m1, n2 = m2, n3 = m3, n4 = m4, n5 = m5, n6 = m6;
    if (strong)
    {
      n3 = clip3(m3 - 2 * tc, m3 + 2 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
      n4 = clip3(m4 - 2 * tc, m4 + 2 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
      n2 = clip3(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
      n5 = clip3(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
      n1 = clip3(m1 - 2 * tc, m1 + 2 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
      n6 = clip3(m6 - 2 * tc, m6 + 2 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
    }
    else
    {
      int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
      if (iabs(delta) < thr_cut)
      {
        delta = clip3(-tc, tc, delta);
        n3 = clip3(0, maxv, m3 + delta);
        n4 = clip3(0, maxv, m4 - delta);
        const int tc2 = tc >> 1;
        if (fp) n2 = clip3(0, maxv, m2 + clip3(-tc2, tc2, ((((m1 + m3 + 1) >> 1) - m2 + delta) >> 1)));
        if (fq) n5 = clip3(0, maxv, m5 + clip3(-tc2, tc2, ((((m6 + m4 + 1) >> 1) - m5 - delta) >> 1)));
      }
    }
    if (!nofilt_p) { S(-1, i) = (int16_t)n3; S(-2, i) = (int16_t)n2; S(-3, i) = (int16_t)n1; }
    if (!nofilt_q) { S(0, i) = (int16_t)n4; S(1, i) = (int16_t)n5; S(2, i) = (int16_t)n6; }
  }
#undef S
}

static void deblock_chroma_line(int16_t* p, int step, int tc, int bd, int nofilt_p, int nofilt_q)
{
  const int m2 = p[-2 * step], m3 = p[-step], m4 = p[0], m5 = p[step], maxv = (1 << bd) - 1;
  const int delta = clip3(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
  if (!nofilt_p) p[-step] = (int16_t)clip3(0, maxv, m3 + delta);
  if (!nofilt_q) p[0] = (int16_t)clip3(0, maxv, m4 - delta);
}

/* dir 0: vertical edges (filter across x), dir 1: horizontal edges.  Whole picture, in place (TComLoopFilter.cpp:130-155). */
void orc_deblock_pass(const hmr_frame_desc* f, orc_pic* pic, int dir)
{
  const hmr_frame_hdr* h = f->hdr;
  if (!(h->flags & HMR_FRM_DEBLOCK) || !f->bs) return;
  const int W4 = (h->width + 3) >> 2, H4 = (h->height + 3) >> 2, W8 = (h->width + 7) >> 3;
  const int fmt = h->chroma_format, ctu_shift = h->log2_ctu, ctus_w = (h->width + (1 << ctu_shift) - 1) >> ctu_shift;
  const int cx = csx_of(fmt, 1), cy = csy_of(fmt, 1);
  for (int uy = 0; uy < H4; uy++)
    for (int ux = 0; ux < W4; ux++)
    {
      const int bs = (f->bs[(size_t)uy * W4 + ux] >> (dir ? 2 : 0)) & 3;
      if (!bs) continue;
      const int x = ux * 4, y = uy * 4;
      const int px = dir ? x : x - 1, py = dir ? y - 1 : y;       /* a sample on the P side */
      const int qp_q = f->qp[(size_t)(y >> 3) * W8 + (x >> 3)], qp_p = f->qp[(size_t)(py >> 3) * W8 + (px >> 3)];
      const int nf_q = f->cu_flags ? (f->cu_flags[(size_t)(y >> 3) * W8 + (x >> 3)] & HMR_CU_NOFILTER) : 0;
      const int nf_p = f->cu_flags ? (f->cu_flags[(size_t)(py >> 3) * W8 + (px >> 3)] & HMR_CU_NOFILTER) : 0;
      const hmr_ctu* cq = &f->ctu[(size_t)(y >> ctu_shift) * ctus_w + (x >> ctu_shift)];
      const int qp = (qp_p + qp_q + 1) >> 1;
      {
        int16_t* p = pic->plane[0] + (size_t)y * pic->stride[0] + x;
        deblock_luma_segment(p, dir ? pic->stride[0] : 1, dir ? 1 : pic->stride[0], bs, qp, cq->beta_offset_div2, cq->tc_offset_div2, h->bit_depth_luma, nf_p, nf_q);
      }
      /* chroma: BS 2 only, edges on a grid of 8 chroma samples (TComLoopFilter.cpp:684-692, 220-229) */
      if (bs > 1 && fmt != HMR_CHROMA_400)
      {
        const int grid = dir ? (8 << cy) : (8 << cx);
        if (((dir ? y : x) % grid) != 0) continue;
        const int nlines = dir ? (4 >> cx) : (4 >> cy);
        for (int c = 1; c < 3; c++)
        {
          int q = qp + (c == 1 ? h->pps_cb_qp_offset : h->pps_cr_qp_offset);
          if (q >= 58) { if (fmt == HMR_CHROMA_420) q -= 6; else if (q > 51) q = 51; }
          else if (q >= 0) q = fmt == HMR_CHROMA_420 ? kChromaQp420[q] : (q > 51 ? 51 : q);
          const int tc = kTc[clip3(0, 53, q + 2 * (bs - 1) + (cq->tc_offset_div2 << 1))] * (1 << (h->bit_depth_chroma - 8));
          int16_t* p = pic->plane[c] + (size_t)(y >> cy) * pic->stride[c] + (x >> cx);
          for (int i = 0; i < nlines; i++)
            deblock_chroma_line(p + (size_t)i * (dir ? 1 : pic->stride[c]), dir ? pic->stride[c] : 1, tc, h->bit_depth_chroma, nf_p, nf_q);
        }
      }
    }
}

/* ---------------------------------------------------------------------------------------------
 * SAO: offsetCTU / offsetBlock (TComSampleAdaptiveOffset.cpp:375-714), reading deblocked samples `src`, writing `dst`. */
static inline int sgn(int v) { return (v > 0) - (v < 0); }

void orc_sao(const hmr_frame_desc* f, const orc_pic* src, orc_pic* dst)
{
  const hmr_frame_hdr* h = f->hdr;
  const int fmt = h->chroma_format, ctu = 1 << h->log2_ctu, ctus_w = (h->width + ctu - 1) >> h->log2_ctu;
  for (int c = 0; c < 3; c++)
    for (int y = 0; y < src->height[c]; y++)
      memcpy(dst->plane[c] + (size_t)y * dst->stride[c], src->plane[c] + (size_t)y * src->stride[c], sizeof(int16_t) * src->width[c]);
  if (!(h->flags & HMR_FRM_SAO)) return;
  for (uint32_t a = 0; a < h->n_ctu; a++)
  {
    const hmr_ctu* cp = &f->ctu[a];
    const int av = cp->avail;
    const int L = av & HMR_AV_L, R = av & HMR_AV_R, A = av & HMR_AV_A, B = av & HMR_AV_B;
    const int AL = av & HMR_AV_AL, AR = av & HMR_AV_AR, BL = av & HMR_AV_BL, BR = av & HMR_AV_BR;
    for (int c = 0; c < (fmt == HMR_CHROMA_400 ? 1 : 3); c++)
    {
      const hmr_sao* s = &cp->sao[c];
      if (s->type == HMR_SAO_OFF) continue;
      const int cx = csx_of(fmt, c), cy = csy_of(fmt, c), bd = c ? h->bit_depth_chroma : h->bit_depth_luma, maxv = (1 << bd) - 1;
      const int bx = ((int)(a % ctus_w) << h->log2_ctu) >> cx, by = ((int)(a / ctus_w) << h->log2_ctu) >> cy;
      int bw = ctu >> cx, bh = ctu >> cy;
      if (bx + bw > src->width[c]) bw = src->width[c] - bx;
      if (by + bh > src->height[c]) bh = src->height[c] - by;
      const int st = src->stride[c];
      for (int y = 0; y < bh; y++) for (int x = 0; x < bw; x++)
      {
        const int16_t* p = src->plane[c] + (size_t)(by + y) * st + bx + x;
        const int v = *p;
        int out = v;
        /* TComSampleAdaptiveOffset::PCMLFDisableProcess / xPCMRestoration (:743-843): I_PCM (with pcm_loop_filter_disabled)
           and lossless CUs get their pre-filter samples back, i.e. SAO never applies to them */
        if (f->cu_flags && (f->cu_flags[(size_t)(((by + y) << cy) >> 3) * ((h->width + 7) >> 3) + (((bx + x) << cx) >> 3)] & HMR_CU_NOFILTER)) { }
        else if (s->type == HMR_SAO_BO)
        {
          const int k = ((v >> (bd - 5)) - s->band) & 31;
          if (k < 4) out = clip3(0, maxv, v + s->off[k]);
        }
        else
        {
          int ok, dx, dy;
          const int first_row = y == 0, last_row = y == bh - 1, first_col = x == 0, last_col = x == bw - 1;
          switch (s->type)
          {
            case HMR_SAO_EO_0:  dx = 1; dy = 0; ok = !(first_col && !L) && !(last_col && !R); break;
            case HMR_SAO_EO_90: dx = 0; dy = 1; ok = !(first_row && !A) && !(last_row && !B); break;
            case HMR_SAO_EO_135: /* neighbours (-1,-1) and (+1,+1) */
              dx = 1; dy = 1;
              if (first_row && bh > 1) ok = first_col ? (AL != 0) : (A && !(last_col && !R));
              else if (last_row)       ok = last_col ? (BR != 0) : (B && !(first_col && !L));
              else                     ok = !(first_col && !L) && !(last_col && !R);
              break;
            default: /* HMR_SAO_EO_45: neighbours (+1,-1) and (-1,+1) */
              dx = -1; dy = 1;
              if (first_row && bh > 1) ok = last_col ? (AR != 0) : (A && !(first_col && !L));
              else if (last_row)       ok = first_col ? (BL != 0) : (B && !(last_col && !R));
              else                     ok = !(first_col && !L) && !(last_col && !R);
              break;
          }
          if (ok)
          {
            const int a0 = p[-dy * st - dx], b0 = p[dy * st + dx];
            const int e = sgn(v - a0) + sgn(v - b0);
            if (e) out = clip3(0, maxv, v + s->off[e < 0 ? e + 2 : e + 1]);
          }
        }
        dst->plane[c][(size_t)(by + y) * dst->stride[c] + bx + x] = (int16_t)out;
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * Whole picture.  stage_mask bits: 1 inter prediction, 2 residual (+ add for inter TUs), 4 intra, 8 deblock V,
 * 16 deblock H, 32 SAO.  `work` receives the picture before SAO (TDecCu::decompressCU + loopFilterPic), dpb[out_slot]
 * the final picture.  `resid` is scratch for the compact residual buffer, n_coef int16 entries.  Returns 0. */
int orc_reconstruct_frame(const hmr_frame_desc* f, orc_pic* dpb, orc_pic* work, int16_t* resid, int stage_mask)
{
  const hmr_frame_hdr* h = f->hdr;
  const int fmt = h->chroma_format;
  if (h->magic != HMR_MAGIC || h->version != HMR_VERSION) return -1;

  if (stage_mask & 1)
    for (uint32_t i = 0; i < h->n_pu; i++)
      orc_predict_pu(h, &f->pu[i], (h->flags & HMR_FRM_WEIGHTED_PRED) ? f->wp : NULL, (h->flags & HMR_FRM_WEIGHTED_PRED) ? f->pu_refidx[i] : 0, dpb, work);

  if (stage_mask & 2)
  {
    int16_t blk[ORC_MAX_TU * ORC_MAX_TU];
    for (uint32_t i = 0; i < h->n_tu; i++)
    {
      const hmr_tu* t = &f->tu[i];
      const int c = t->comp, N = 1 << t->log2_size, bd = c ? h->bit_depth_chroma : h->bit_depth_luma;
      orc_tu_residual(t, f->coef + t->coef_off, bd, (h->flags & HMR_FRM_SCALING_LIST) ? f->scaling : NULL, blk);
      if (t->ccp_alpha && t->luma_off != HMR_NO_OFFSET)
        ccp_apply(blk, resid + t->luma_off, N * N, t->ccp_alpha, h->bit_depth_luma - h->bit_depth_chroma);
      const int keep = (t->flags & HMR_TU_INTRA) || (c == 0 && (h->flags & HMR_FRM_HAS_CCP));
      if (keep) memcpy(resid + t->coef_off, blk, sizeof(int16_t) * N * N);
      if (!(t->flags & HMR_TU_INTRA))
      {
        /* reco = ClipBD(pred + resi): TComYuv::addClip (TComYuv.cpp:264-299) */
        const int maxv = (1 << bd) - 1;
        for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
        {
          int16_t* d = work->plane[c] + (size_t)(t->y + y) * work->stride[c] + t->x + x;
          *d = (int16_t)clip3(0, maxv, *d + blk[y * N + x]);
        }
      }
    }
  }

  if (stage_mask & 4)
  {
    int16_t pred[ORC_MAX_TU * ORC_MAX_TU];
    for (uint32_t a = 0; a < h->n_ctu; a++)
      for (int c = 0; c < 3; c++)
      {
        const hmr_ctu_intra_range* rg = &f->intra_range[a];
        for (uint32_t k = 0; k < rg->count[c]; k++)
        {
          const hmr_intra* r = &f->intra[rg->first[c] + k];
          const int N = 1 << r->log2_size, bd = c ? h->bit_depth_chroma : h->bit_depth_luma, maxv = (1 << bd) - 1;
          orc_intra_predict(h, r, work, pred);
          const int16_t* rs = r->resid_off != HMR_NO_OFFSET ? resid + r->resid_off : NULL;
          for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
            work->plane[c][(size_t)(r->y + y) * work->stride[c] + r->x + x] = (int16_t)clip3(0, maxv, pred[y * N + x] + (rs ? rs[y * N + x] : 0));
        }
      }
  }
  (void)fmt;
  if (stage_mask & 8)  orc_deblock_pass(f, work, 0);
  if (stage_mask & 16) orc_deblock_pass(f, work, 1);
  if (stage_mask & 32) orc_sao(f, work, &dpb[h->out_slot]);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Picture hashes of the SEI decoded-picture-hash message other than MD5 (TComPicYuvMD5.cpp:87-175). */
uint32_t orc_checksum_plane(const int16_t* p, int w, int h, int stride, int bit_depth)
{
  uint32_t sum = 0;
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
  {
    const uint32_t mask = (uint32_t)((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8));
    const int v = p[(size_t)y * stride + x];
    sum += (uint32_t)((v & 0xff) ^ mask);
    if (bit_depth > 8) sum += (uint32_t)((v >> 8) ^ mask);
  }
  return sum;
}

uint32_t orc_crc_plane(const int16_t* p, int w, int h, int stride, int bit_depth)
{
  uint32_t crc = 0xffff;
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
  {
    const int v = p[(size_t)y * stride + x];
    for (int byte = 0; byte < (bit_depth > 8 ? 2 : 1); byte++)
    {
      const int b = byte ? (v >> 8) & 0xff : v & 0xff;
      for (int bit = 0; bit < 8; bit++)
      {
        const uint32_t msb = (crc >> 15) & 1, in = (uint32_t)(b >> (7 - bit)) & 1;
        crc = (((crc << 1) + in) & 0xffff) ^ (msb * 0x1021);
      }
    }
  }
  for (int bit = 0; bit < 16; bit++) { const uint32_t msb = (crc >> 15) & 1; crc = ((crc << 1) & 0xffff) ^ (msb * 0x1021); }
  return crc;
}
