// k_mc.cu — inter prediction: 8-tap luma / 4-tap chroma fractional-sample interpolation with bi-prediction average.
//
// Replaces TComPrediction::motionCompensation -> xPredInterUni/xPredInterBi -> xPredInterBlk
// (TComPrediction.cpp:514-698), TComInterpolationFilter::filter / filterCopy (TComInterpolationFilter.cpp:94-251)
// and TComYuv::addAvg (TComYuv.cpp:336-391).
//
// Work unit: one 16x16-luma tile of a PU, owned by ONE WARP — no block-level synchronisation anywhere.  A tiny pre-pass
// (mc_expand_kernel) turns the PU records into one 16-byte record per tile so that the main kernels need a single
// dependent load before they can fetch samples.  Two kinds of work item per tile: the luma tile, and the two co-located chroma
// tiles (half the shared memory per warp each: the kernels wait on memory, and warps in flight are what hides it) — one
// launch for both in 4:2:0 (mc_kernel_420: same shared memory per warp), mc_kernel<true> + mc_kernel<false> otherwise.  Cb and Cr of an 8-wide chroma tile (4:2:0, 4:2:2) have the same geometry,
// phases and taps: they are staged together (16 lanes each) and filtered together (half of the active lanes each).
//
// Per tile: every reference window (up to 2 lists; chroma: x 2 planes) is fetched up front with 16-byte cp.async copies
// that are all in flight together — only the vectors the window touches, rounds unrolled and row-predicated (rows start
// at the 16-byte boundary below the window; the sub-alignment `off` is handled when reading shared memory).  Windows
// that touch the picture border are gathered with clamped coordinates instead, which is exactly HM's replicated
// border (TComPicYuv::extendPicBorder, TComPicYuv.cpp:173-217) without ever materialising it.
//
// Arithmetic: every case of HM (copy / H only / V only / H+V, uni / bi) is ONE separable pipeline
// H -> 14-bit intermediate -> V with the identity tap set for a zero fraction; that is bit-identical to HM's special
// cases because the intermediate offset (8192 << s1) is a multiple of the first-stage divisor.  Both passes run on
// packed int16 pairs with dp2a (2 MACs per instruction, no unpacking): the H pass reads sample pairs along a row, and
// writes its 14-bit results packed as (row r, row r+1) pairs so that the V pass again sees pairs along its filter
// direction.  An output at an even position uses NT/2 dp2a with taps (t0,t1)(t2,t3)..; an output at an odd position
// uses NT/2+1 with the taps shifted by one half-word (0,t0)(t1,t2)..(t7,0).  All sums are exact int32, as in HM.
#include "common.cuh"
#include <cstdlib>

#ifndef MC_WARPS
#define MC_WARPS 4
#endif
#ifndef MC_MIN_BLOCKS
#define MC_MIN_BLOCKS 9               // 56 registers.  Measured: 10 (<= 51 regs, 4-24 B spilled) 47.0 -> 48.8 us, 12 (<= 42 regs) 53.2 us; value unchanged
#endif
#define MC_PITCH 40                 // int16 per staged window row: up to 32 loaded + 8 pad (80 B: 16-byte aligned, conflict-free row pairs)
#define MC_TMPW 16                  // words per row pair of the H-pass output

// HM's interpolation taps (TComInterpolationFilter.cpp:50-72), packed for dp2a: `e` = byte pairs (t0,t1)(t2,t3).. for an
// output at an even position, `o` = (0,t0)(t1,t2)..(t7,0) for an odd one.  Filled by launch_mc on first use.
struct McTapTable { int lumaE[4][4], lumaO[4][5], chromaE[8][2], chromaO[8][3]; };
__constant__ McTapTable c_taps;
static const int8_t h_lumaTaps[4][8] = { {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1} };
static const int8_t h_chromaTaps[8][4] = { {0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4}, {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2} };

template <int NT> static void pack_taps_host(const int8_t* t, int* e, int* o)
{
  for (int j = 0; j < NT / 2; j++) e[j] = (t[2 * j] & 0xff) | ((t[2 * j + 1] & 0xff) << 8);
  for (int j = 0; j <= NT / 2; j++) o[j] = (j > 0 ? (t[2 * j - 1] & 0xff) : 0) | (j < NT / 2 ? ((t[2 * j] & 0xff) << 8) : 0);
}
static void upload_taps()
{
  static bool done[64] = { false };
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && done[dev]) return;
  McTapTable T;
  for (int f = 0; f < 4; f++) pack_taps_host<8>(h_lumaTaps[f], T.lumaE[f], T.lumaO[f]);
  for (int f = 0; f < 8; f++) pack_taps_host<4>(h_chromaTaps[f], T.chromaE[f], T.chromaO[f]);
  cudaMemcpyToSymbol(c_taps, &T, sizeof(T));
  if (dev < 64) done[dev] = true;
}

__device__ __forceinline__ void mc_cp_async16(void* smem, const void* gmem)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void mc_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void mc_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// PU records -> one record per 16x16-luma tile (same struct: x, y, w, h describe the tile).  16 threads per PU (a PU has at
// most 4x4 tiles): the record load is a broadcast, the tile records of a PU are written by consecutive lanes.
__global__ void __launch_bounds__(256) mc_expand_kernel(const __grid_constant__ FrameParams P)
{
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t p = gid >> 4;
  const int k = gid & 15;
  if (p >= P.hdr.n_pu) return;
  const uint4 raw = __ldg((const uint4*)(P.pu + p));
  hmr_pu t = *(const hmr_pu*)&raw;
  const int tx = (t.w + 15) >> 4, ty = (t.h + 15) >> 4;
  if (k >= tx * ty) return;
  const int j = k / tx, i = k - j * tx;
  const int w = min(16, t.w - 16 * i), h = min(16, t.h - 16 * j);
  t.x = (uint16_t)(t.x + 16 * i); t.y = (uint16_t)(t.y + 16 * j);
  t.w = (uint8_t)w; t.h = (uint8_t)h;
  const uint32_t dstIdx = __ldg(P.pu_prefix + p) + k;
  *(uint4*)(P.mc_tiles + dstIdx) = *(const uint4*)&t;
  if (P.pu_refidx) P.mc_tile_refidx[dstIdx] = P.pu_refidx[p];
}

// Stage rows [iy, iy+rows) x columns [xa, xa + 8*nvec) of `ref` into s (pitch MC_PITCH).  Returns nothing; async on the fast path.
__device__ __forceinline__ void mc_stage(int16_t* s, const int16_t* __restrict__ ref, int rpitch, int Wc, int Hc,
                                         int ix, int iy, int xa, int rows, int cols, int nvec, int lane)
{
  const bool inside = iy >= 0 && iy + rows <= Hc && ix >= 0 && ix + cols <= Wc && xa + 32 <= rpitch;
  if (inside)
  {
    // up to 4 x 16 bytes per row (only the vectors the window touches): 8 rows per round, pointers only advance
    const int16_t* g = ref + (size_t)(iy + (lane >> 2)) * rpitch + xa + 8 * (lane & 3);
    int16_t* d = s + (lane >> 2) * MC_PITCH + 8 * (lane & 3);
    // rows <= 24: at most 3 rounds of 8 rows, unrolled with a row predicate (no loop-carried 64-bit pointer updates)
    if ((lane & 3) < nvec)
    {
#pragma unroll
      for (int k = 0; k < 3; k++)
        if ((lane >> 2) + 8 * k < rows) mc_cp_async16(d + 8 * k * MC_PITCH, g + (size_t)(8 * k) * rpitch);
    }
  }
  else
  {
    const int n = 8 * nvec;
    for (int i = lane; i < rows * n; i += 32)
    {
      const int r = i / n, c = i - r * n;
      const int yy = clip3i(0, Hc - 1, iy + r), xx = clip3i(0, Wc - 1, xa + c);
      s[r * MC_PITCH + c] = ref[(size_t)yy * rpitch + xx];
    }
  }
}

// The Cb and Cr windows of a tile have the same geometry: lanes 0-15 stage Cb, lanes 16-31 Cr (4 rows x 4 vectors per round).
__device__ __forceinline__ void mc_stage_pair(int16_t* sCb, int16_t* sCr, const int16_t* __restrict__ refCb, const int16_t* __restrict__ refCr,
                                              int rpitch, int Wc, int Hc, int ix, int iy, int xa, int rows, int cols, int nvec, int lane)
{
  const bool inside = iy >= 0 && iy + rows <= Hc && ix >= 0 && ix + cols <= Wc && xa + 32 <= rpitch;
  const int l16 = lane & 15;
  const int16_t* __restrict__ ref = (lane & 16) ? refCr : refCb;
  int16_t* s = (lane & 16) ? sCr : sCb;
  if (inside)
  {
    const int16_t* g = ref + (size_t)(iy + (l16 >> 2)) * rpitch + xa + 8 * (l16 & 3);
    int16_t* d = s + (l16 >> 2) * MC_PITCH + 8 * (l16 & 3);
    // rows <= 20 (4:2:2 tile: 16 + 4): at most 5 rounds of 4 rows, unrolled with a row predicate
    if ((l16 & 3) < nvec)
    {
#pragma unroll
      for (int k = 0; k < 5; k++)
        if ((l16 >> 2) + 4 * k < rows) mc_cp_async16(d + 4 * k * MC_PITCH, g + (size_t)(4 * k) * rpitch);
    }
  }
  else
  {
    const int n = 8 * nvec;
    for (int i = l16; i < rows * n; i += 16)
    {
      const int r = i / n, c = i - r * n;
      const int yy = clip3i(0, Hc - 1, iy + r), xx = clip3i(0, Wc - 1, xa + c);
      s[r * MC_PITCH + c] = ref[(size_t)yy * rpitch + xx];
    }
  }
}

template <int NT> struct McTaps { int e[NT / 2]; int o[NT / 2 + 1]; };

template <int NT>
__device__ __forceinline__ McTaps<NT> mc_load_taps(int frac)      // frac: quarter-sample (luma) / eighth-sample (chroma) phase
{
  McTaps<NT> r;
#pragma unroll
  for (int j = 0; j < NT / 2; j++) r.e[j] = NT == 8 ? c_taps.lumaE[frac][j] : c_taps.chromaE[frac][j];
#pragma unroll
  for (int j = 0; j <= NT / 2; j++) r.o[j] = NT == 8 ? c_taps.lumaO[frac][j] : c_taps.chromaO[frac][j];
  return r;
}

template <int NT> __device__ __forceinline__ int mc_even(const int* w, const McTaps<NT>& t)
{
  int s = 0;
#pragma unroll
  for (int j = 0; j < NT / 2; j++) s = __dp2a_lo(w[j], t.e[j], s);
  return s;
}
template <int NT> __device__ __forceinline__ int mc_odd(const int* w, const McTaps<NT>& t)
{
  int s = 0;
#pragma unroll
  for (int j = 0; j <= NT / 2; j++) s = __dp2a_lo(w[j], t.o[j], s);
  return s;
}

// H pass over one staged window: rows 2*rp, 2*rp+1 x output columns 2*cp, 2*cp+1 per item; results packed (row 2rp | row 2rp+1 << 16).
// PAIR: two windows with the same geometry and phase (the Cb and Cr windows of an 8-wide chroma tile) in one sweep; the
// second plane's results go to words 8..15 of each row pair.
template <int NT, bool PAIR>
__device__ __forceinline__ void mc_hpass(const int16_t* s, const int16_t* sB, uint32_t* tmp, int off, int rowPairs, int log2ColPairs,
                                         const McTaps<NT>& tx, int s1, int o1, int lane)
{
  const int nItems = rowPairs << (log2ColPairs + (PAIR ? 1 : 0));
  const bool oddStart = off & 1;
  for (int it = lane; it < nItems; it += 32)
  {
    const int cp = it & ((1 << log2ColPairs) - 1);
    const int second = PAIR ? (it >> log2ColPairs) & 1 : 0;
    const int rp = it >> (log2ColPairs + (PAIR ? 1 : 0));
    const int* r0 = (const int*)((second ? sB : s) + 2 * rp * MC_PITCH) + (off >> 1) + cp;
    const int* r1 = r0 + MC_PITCH / 2;
    int w0[NT / 2 + 1], w1[NT / 2 + 1];
#pragma unroll
    for (int j = 0; j <= NT / 2; j++) { w0[j] = r0[j]; w1[j] = r1[j]; }
    int a0, a1, b0, b1;             // a: row 2rp, b: row 2rp+1; 0/1: output columns 2cp, 2cp+1
    if (!oddStart) { a0 = mc_even<NT>(w0, tx); a1 = mc_odd<NT>(w0, tx); b0 = mc_even<NT>(w1, tx); b1 = mc_odd<NT>(w1, tx); }
    else           { a0 = mc_odd<NT>(w0, tx);  a1 = mc_even<NT>(w0 + 1, tx); b0 = mc_odd<NT>(w1, tx); b1 = mc_even<NT>(w1 + 1, tx); }
    a0 = (a0 + o1) >> s1; a1 = (a1 + o1) >> s1; b0 = (b0 + o1) >> s1; b1 = (b1 + o1) >> s1;     // truncated to Pel by the packing
    uint2 out;
    out.x = (uint32_t)(a0 & 0xffff) | ((uint32_t)b0 << 16);
    out.y = (uint32_t)(a1 & 0xffff) | ((uint32_t)b1 << 16);
    *(uint2*)(tmp + rp * MC_TMPW + 8 * second + 2 * cp) = out;
  }
}

// V pass for this lane's item: output rows 2*yg, 2*yg+1 x columns 4*x4 .. 4*x4+3.  v[0..3] = row 2yg, v[4..7] = row 2yg+1 (raw sums).
template <int NT>
__device__ __forceinline__ void mc_vpass(const uint32_t* tmp, int x4, int yg, const McTaps<NT>& ty, int v[8])
{
  int w[4][NT / 2 + 1];
#pragma unroll
  for (int j = 0; j <= NT / 2; j++)
  {
    const uint4 q = *(const uint4*)(tmp + (yg + j) * MC_TMPW + 4 * x4);
    w[0][j] = (int)q.x; w[1][j] = (int)q.y; w[2][j] = (int)q.z; w[3][j] = (int)q.w;
  }
#pragma unroll
  for (int c = 0; c < 4; c++) { v[c] = mc_even<NT>(w[c], ty); v[4 + c] = mc_odd<NT>(w[c], ty); }
}

// refIdx = refIdx of list 0 | list 1 << 4 (explicit weighted prediction only: TComWeightPrediction::addWeightUni / addWeightBi,
// TComWeightPrediction.cpp:44-53,75-196 — both cases start from the 14-bit intermediates, like bi-prediction)
// PAIR (8-wide chroma tiles, i.e. 4:2:0 and 4:2:2): planes `comp` and `comp + 1` together — the first half of the active lanes
// owns Cb, the second half Cr (same geometry, phases and offsets; windows srefB), which halves the instructions of a chroma tile.
template <int NT, bool PAIR>
__device__ __forceinline__ void mc_component(const FrameParams& P, const hmr_pu& t, int comp, int cx, int cy, int16_t* const sref[2],
                                             int16_t* const srefB[2], const int offs[2], uint32_t* tmp, int lane, int refIdx)
{
  const int tw = 16 >> cx, th = 16 >> cy;                    // full tile in this component
  const int x0 = t.x >> cx, y0 = t.y >> cy, w = t.w >> cx, h = t.h >> cy;
  const int bd = comp ? P.hdr.bit_depth_chroma : P.hdr.bit_depth_luma;
  const int headroom = max(2, 14 - bd), maxv = (1 << bd) - 1;
  const int s1 = 6 - headroom, o1 = -(8192 << s1);
  const bool bi = t.lists == (HMR_PU_L0 | HMR_PU_L1);
  const bool wpOn = P.wp != nullptr;
  const int second = PAIR ? (lane >= ((tw >> 2) * (th >> 1)) ? 1 : 0) : 0;
  const int laneInPlane = lane - second * ((tw >> 2) * (th >> 1));
  const int myComp = comp + second;
  int wpW[2] = {0, 0}, wpO[2] = {0, 0}, wpD = 0;
  if (wpOn)
  {
    for (int l = 0; l < 2; l++)
    {
      if (!(t.lists & (1 << l))) continue;
      const hmr_wp q = P.wp[(l * 16 + ((refIdx >> (4 * l)) & 15)) * 3 + myComp];
      wpW[l] = q.weight; wpO[l] = q.offset;
      if (l == 0 || !(t.lists & HMR_PU_L0)) wpD = q.log2_denom;          // bi uses the list-0 denominator (getWpScaling)
    }
  }
  const int rowPairs = (th + NT) >> 1;                       // th + NT - 1 rows, rounded up to even
  const int log2ColPairs = tw == 16 ? 3 : 2;
  const int nV1 = (tw >> 2) * (th >> 1);                     // V items of one plane: 4 columns x 2 rows each
  const int nV = PAIR ? 2 * nV1 : nV1;                       // <= 32
  const int x4 = laneInPlane & ((tw >> 2) - 1), yg = laneInPlane / (tw >> 2);
  int acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0;

  for (int list = 0; list < 2; list++)
  {
    if (!(t.lists & (1 << list))) continue;                  // warp-uniform
    const int mvx = t.mv[list][0], mvy = t.mv[list][1];
    const int fx = mvx & ((4 << cx) - 1), fy = mvy & ((4 << cy) - 1);
    if ((fx | fy) == 0)
    {
      // TComInterpolationFilter::filterCopy (TComInterpolationFilter.cpp:94-148): the staged window IS the block
      if (lane < nV)
      {
        const int16_t* src = (second ? srefB[list] : sref[list]) + 2 * yg * MC_PITCH + offs[list] + 4 * x4;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
          const int smp = src[(i >> 2) * MC_PITCH + (i & 3)];
          if (wpOn)    acc[i] += wpW[list] * (smp << headroom);                     // (P + IF_INTERNAL_OFFS), P = (s << headroom) - 8192
          else if (bi) acc[i] += (smp << headroom) - 8192;                          // fits Pel for bit depths <= 12 (engine limit): no wrap
          else         acc[i] = smp;
        }
      }
      continue;
    }
    const McTaps<NT> tx = mc_load_taps<NT>(NT == 8 ? fx : fx << (1 - cx));
    const McTaps<NT> ty = mc_load_taps<NT>(NT == 8 ? fy : fy << (1 - cy));
    mc_hpass<NT, PAIR>(sref[list], PAIR ? srefB[list] : nullptr, tmp, offs[list], rowPairs, log2ColPairs, tx, s1, o1, lane);
    __syncwarp();
    if (lane < nV)
    {
      int v[8];
      mc_vpass<NT>(tmp + 8 * second, x4, yg, ty, v);
      if (wpOn)
      {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] += wpW[list] * ((int)(int16_t)(v[i] >> 6) + 8192);
      }
      else if (bi)
      {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] += (int)(int16_t)(v[i] >> 6);       // xPredInterBlk, !isLast: Pel result
      }
      else
      {
        const int s2 = 6 + headroom, o2 = (1 << (s2 - 1)) + (8192 << 6);
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = clip3i(0, maxv, (int)(int16_t)((v[i] + o2) >> s2));
      }
    }
    __syncwarp();                                            // tmp is reused by the next list / component
  }
  if (lane >= nV) return;
  if (wpOn)
  {
    if (bi)
    {
      const int sh = wpD + 1 + headroom, add = (1 << (sh - 1)) + ((wpO[0] + wpO[1]) << (sh - 1));
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = clip3i(0, maxv, (acc[i] + add) >> sh);
    }
    else
    {
      const int sh = wpD + headroom, rnd = sh > 0 ? 1 << (sh - 1) : 0, o = (t.lists & HMR_PU_L0) ? wpO[0] : wpO[1];
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = clip3i(0, maxv, ((acc[i] + rnd) >> sh) + o);
    }
  }
  else if (bi)
  {
    const int sh = headroom + 1, off = (1 << (sh - 1)) + 2 * 8192;             // TComYuv::addAvg
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = clip3i(0, maxv, (acc[i] + off) >> sh);
  }
  const int dpitch = P.work.pitch[myComp];
  int16_t* dst = P.work.p[myComp] + (size_t)(y0 + 2 * yg) * dpitch + x0 + 4 * x4;
#pragma unroll
  for (int r = 0; r < 2; r++)
  {
    if (2 * yg + r >= h) break;
    int16_t* d = dst + (size_t)r * dpitch;
    const uint32_t lo = (uint32_t)(acc[4 * r] & 0xffff) | ((uint32_t)acc[4 * r + 1] << 16);
    const uint32_t hi = (uint32_t)(acc[4 * r + 2] & 0xffff) | ((uint32_t)acc[4 * r + 3] << 16);
    if (4 * x4 + 4 <= w && (((x0 + 4 * x4) & 3) == 0)) *(uint2*)d = make_uint2(lo, hi);
    else
    {
      if (4 * x4 + 2 <= w) *(uint32_t*)d = lo;
      if (4 * x4 + 4 <= w) *(uint32_t*)(d + 2) = hi;
    }
  }
}

// ---- integer-sample motion: TComInterpolationFilter::filterCopy (TComInterpolationFilter.cpp:94-148) + TComYuv::addAvg -----------
// When every list of a tile has a zero fraction in this component (HM's copy case) and no window leaves the picture, the tile never
// goes through shared memory and the filter pipeline: a lane owns 8 (luma) / 4 (chroma 4:2:0) consecutive samples of one row, fetches
// the two aligned 16-byte vectors that hold them straight from the DPB, realigns them with byte permutes and stores.  Bi-prediction of
// two copies is the rounded average: HM's (P0 + P1 + offset) >> shift on the 14-bit intermediates ((s << h) - 8192 each, offset =
// (1 << h) + 16384, shift = h + 1) equals (s0 + s1 + 1) >> 1 exactly, and is within the sample range by construction (no clip
// needed): one __vavgu2 per two samples.
// Word j (two samples) of the 8 samples that start `off` samples into the 16 samples q0|q1.
__device__ __forceinline__ void mc_realign8(const uint4 q0, const uint4 q1, const int off, uint32_t out[4])
{
  const uint32_t s[9] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, 0u };
  switch (off)
  {
#define MC_CASE(o) case o: _Pragma("unroll") for (int j = 0; j < 4; j++) out[j] = ((o) & 1) ? __funnelshift_r(s[((o) >> 1) + j], s[((o) >> 1) + j + 1], 16) : s[((o) >> 1) + j]; break;
    MC_CASE(0) MC_CASE(1) MC_CASE(2) MC_CASE(3) MC_CASE(4) MC_CASE(5) MC_CASE(6) default: MC_CASE(7)
#undef MC_CASE
  }
}
// explicit store instructions: blocks start at multiples of 4 (luma) / 2 (chroma) samples and may be that narrow (8x4 / 4x8 PUs), and
// the compiler was seen to fuse neighbouring narrow C++ stores into one wide (misaligned) store
__device__ __forceinline__ void mc_st32(int16_t* p, uint32_t a) { asm volatile("st.global.b32 [%0], %1;" :: "l"(p), "r"(a) : "memory"); }
__device__ __forceinline__ void mc_st64(int16_t* p, uint32_t a, uint32_t b) { asm volatile("st.global.v2.b32 [%0], {%1, %2};" :: "l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void mc_st128(int16_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// true: the tile was predicted here.  PLANES = 1: component `comp` of a tile that is tw x th there (16x16 luma; 16x16 chroma in 4:4:4);
// PLANES = 2: the 8x8 Cb and Cr tiles of 4:2:0 together (lanes 0-15 / 16-31).
template <int PLANES>
__device__ __forceinline__ bool mc_copy_tile(const FrameParams& P, const hmr_pu& t, const int comp, const int cx, const int cy, const int lane)
{
  const int tw = 16 >> cx;
  const int x0 = t.x >> cx, y0 = t.y >> cy, w = t.w >> cx, h = t.h >> cy;
  const int Wc = P.w[comp], Hc = P.h[comp];
  int ix[2] = {0, 0}, iy[2] = {0, 0};
#pragma unroll
  for (int l = 0; l < 2; l++)
  {
    if (!(t.lists & (1 << l))) continue;
    if ((t.mv[l][0] & ((4 << cx) - 1)) | (t.mv[l][1] & ((4 << cy) - 1))) return false;               // fractional: the filter path
    ix[l] = x0 + (t.mv[l][0] >> (2 + cx)); iy[l] = y0 + (t.mv[l][1] >> (2 + cy));
    const int pitch = P.dpb[l ? (t.slots >> 4) : (t.slots & 15)].pitch[comp];
    if (ix[l] < 0 || iy[l] < 0 || ix[l] + tw > Wc || iy[l] + h > Hc || (ix[l] & ~7) + 32 > pitch) return false;   // border: the clamped gather
  }
  const int plane = PLANES == 2 ? comp + (lane >> 4) : comp;
  const int li = PLANES == 2 ? (lane & 15) : lane;
  const int per = PLANES == 2 ? 4 : 8;                       // samples per lane
  const int row = li >> 1, c0 = (li & 1) * per;
  const bool live = row < h && c0 < w;
  uint32_t v[2][4];
#pragma unroll
  for (int l = 0; l < 2; l++)
  {
    if (!(t.lists & (1 << l))) continue;
    const PlaneSet& ref = P.dpb[l ? (t.slots >> 4) : (t.slots & 15)];
    const int sx = ix[l] + c0, xa = sx & ~7;
    const int16_t* g = ref.p[plane] + (size_t)(iy[l] + (live ? row : 0)) * ref.pitch[plane] + xa;
    const uint4 q0 = *(const uint4*)g, q1 = *(const uint4*)(g + 8);
    mc_realign8(q0, q1, sx & 7, v[l]);                        // (4:2:0 pair: the two halves of a row differ by 4 in the offset — two cases per warp)
  }
  if (!live) return true;
  uint32_t r[4];
  const bool bi = t.lists == (HMR_PU_L0 | HMR_PU_L1);
#pragma unroll
  for (int j = 0; j < 4; j++) r[j] = bi ? __vavgu2(v[0][j], v[1][j]) : ((t.lists & HMR_PU_L0) ? v[0][j] : v[1][j]);
  int16_t* d = P.work.p[plane] + (size_t)(y0 + row) * P.work.pitch[plane] + x0 + c0;
  const int n = min(per, w - c0);                             // luma: 4 or 8 (12-wide tiles: 4 in the second half); 4:2:0 chroma: 2 or 4
  const int al = x0 + c0;
  if (PLANES == 2)
  {
    if (n >= 4 && (al & 3) == 0) mc_st64(d, r[0], r[1]);
    else { mc_st32(d, r[0]); if (n >= 4) mc_st32(d + 2, r[1]); }
  }
  else if (n >= 8 && (al & 7) == 0) mc_st128(d, r[0], r[1], r[2], r[3]);
  else { mc_st64(d, r[0], r[1]); if (n >= 8) mc_st64(d + 4, r[2], r[3]); }
  return true;
}

// LUMA = true: the luma tile; false: the two co-located chroma tiles.  Two launches instead of one kernel doing all three
// components: each then needs half the shared memory per warp, which doubles the warps in flight per SM — and what this
// kernel waits for is memory latency (ncu: issue slots idle on cp.async completion), not arithmetic.
// One tile through the general path (fractional vectors, windows across the picture border, explicit weighted prediction).
template <bool LUMA>
__device__ __forceinline__ void mc_general(const FrameParams& P, const hmr_pu& t, const uint32_t tile, uint32_t* tmp, int16_t* win, const int chromaRows, const int lane)
{
  int16_t* sref[3][2];
  int offs[3][2];

  // ---- stage all windows of this launch's component(s), all copies in flight together ----
#pragma unroll
  for (int comp = LUMA ? 0 : 1; comp < (LUMA ? 1 : 2); comp++)     // chroma: Cb and Cr are staged together
  {
    const int cx = comp ? P.csx : 0, cy = comp ? P.csy : 0;
    const int nt = comp ? 4 : 8, half = nt / 2 - 1;
    const int tw = 16 >> cx, th = 16 >> cy;
    const int rows = (th + nt) & ~1, cols = tw + nt - 1;
#pragma unroll
    for (int list = 0; list < 2; list++)
    {
      int16_t* s = comp == 0 ? win + list * 24 * MC_PITCH : win + list * chromaRows * MC_PITCH;
      sref[comp][list] = s;
      offs[comp][list] = 0;
      if (!LUMA) { sref[2][list] = win + (2 + list) * chromaRows * MC_PITCH; offs[2][list] = 0; }
      if (!(t.lists & (1 << list))) continue;
      const int slot = list ? (t.slots >> 4) : (t.slots & 15);
      // integer-sample motion (HM's filterCopy case): only the block itself is needed, no filter support around it
      const bool whole = ((t.mv[list][0] & ((4 << cx) - 1)) | (t.mv[list][1] & ((4 << cy) - 1))) == 0;
      const int ix = (t.x >> cx) + (t.mv[list][0] >> (2 + cx)) - (whole ? 0 : half), iy = (t.y >> cy) + (t.mv[list][1] >> (2 + cy)) - (whole ? 0 : half);
      const int xa = ix & ~7, off = ix - xa;
      offs[comp][list] = off;
      const int wrows = whole ? th : rows, wcols = whole ? tw : cols;
      if (LUMA) mc_stage(s, P.dpb[slot].p[0], P.dpb[slot].pitch[0], P.w[0], P.h[0], ix, iy, xa, wrows, wcols, (off + wcols + 7) >> 3, lane);
      else
      {
        offs[2][list] = off;
        mc_stage_pair(s, sref[2][list], P.dpb[slot].p[1], P.dpb[slot].p[2], P.dpb[slot].pitch[1], P.w[1], P.h[1], ix, iy, xa, wrows, wcols, (off + wcols + 7) >> 3, lane);
      }
    }
  }
  mc_cp_async_commit();
  mc_cp_async_wait<0>();
  __syncwarp();
  const int refIdx = P.wp ? (int)P.mc_tile_refidx[tile] : 0;
  if (LUMA) mc_component<8, false>(P, t, 0, 0, 0, sref[0], sref[0], offs[0], tmp, lane, refIdx);
  else if (P.csx) mc_component<4, true>(P, t, 1, 1, P.csy, sref[1], sref[2], offs[1], tmp, lane, refIdx);   // 8-wide tiles: Cb and Cr together
  else
  {
    mc_component<4, false>(P, t, 1, 0, P.csy, sref[1], sref[1], offs[1], tmp, lane, refIdx);
    mc_component<4, false>(P, t, 2, 0, P.csy, sref[2], sref[2], offs[2], tmp, lane, refIdx);
  }
}

// One warp per tile, four tiles per CTA.  (Tried: a persistent grid whose warps walk the tiles with the next record prefetched — 44 -> 47 /
// 51 / 57 us per 2160p picture at 6 / 8 / 9 CTAs per SM; two tiles per warp with their loads in flight together — no gain, spills at 56
// registers.  The CTA scheduler hides the record -> samples round trips better than either.)
template <bool LUMA>
__device__ __forceinline__ void mc_tile_body(const FrameParams& P, const uint32_t block, const int warpBytes, const int chromaRows, uint8_t* s_mc)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tile = block * MC_WARPS + warp;
  if (tile >= P.hdr.n_mc_tiles) return;
  uint8_t* base = s_mc + (size_t)warp * warpBytes;
  uint32_t* tmp = (uint32_t*)base;                                        // 12 row pairs x 16 words
  int16_t* win = (int16_t*)(base + 12 * MC_TMPW * 4);                     // luma: 2 lists x 24 rows; chroma: 2 planes x 2 lists x chromaRows
  const uint4 raw = __ldg((const uint4*)(P.mc_tiles + tile));
  const hmr_pu t = *(const hmr_pu*)&raw;
  if (!P.wp)                                                               // (explicit weighted prediction: the general path)
  {
    if (LUMA) { if (mc_copy_tile<1>(P, t, 0, 0, 0, lane)) return; }
    else if (P.csx && P.csy) { if (mc_copy_tile<2>(P, t, 1, 1, 1, lane)) return; }
    else if (!P.csx) { if (mc_copy_tile<1>(P, t, 1, 0, P.csy, lane) && mc_copy_tile<1>(P, t, 2, 0, P.csy, lane)) return; }
  }
  mc_general<LUMA>(P, t, tile, tmp, win, chromaRows, lane);
}

template <bool LUMA>
__global__ void __launch_bounds__(MC_WARPS * 32, MC_MIN_BLOCKS) mc_kernel(const __grid_constant__ FrameParams P, const int warpBytes, const int chromaRows)
{
  extern __shared__ __align__(16) uint8_t s_mc[];
  mc_tile_body<LUMA>(P, blockIdx.x, warpBytes, chromaRows, s_mc);
}

// 4:2:0: a warp needs the same shared memory for a luma tile as for the two chroma tiles, so both go into ONE launch — CTAs
// with blockIdx.y == 0 predict luma, the others chroma.  One launch less per picture (a launch costs ~3 us of a saturated GPU's time,
// DESIGN.md "what bounds value"), and for one stream alone the two halves overlap instead of running back to back.
__global__ void __launch_bounds__(MC_WARPS * 32, MC_MIN_BLOCKS) mc_kernel_420(const __grid_constant__ FrameParams P, const int warpBytes, const int chromaRows)
{
  extern __shared__ __align__(16) uint8_t s_mc[];
  if (blockIdx.y == 0) mc_tile_body<true>(P, blockIdx.x, warpBytes, chromaRows, s_mc);
  else mc_tile_body<false>(P, blockIdx.x, warpBytes, chromaRows, s_mc);
}

int launch_mc(const FrameParams& P, cudaStream_t s)
{
  if (P.hdr.n_mc_tiles == 0) return 0;
  upload_taps();
  mc_expand_kernel<<<(P.hdr.n_pu * 16 + 255) / 256, 256, 0, s>>>(P);
  const int chromaRows = ((16 >> P.csy) + 4) & ~1;
  const int grid = (P.hdr.n_mc_tiles + MC_WARPS - 1) / MC_WARPS;
  const int lumaBytes = 12 * MC_TMPW * 4 + 2 * 24 * MC_PITCH * 2;
  const int chromaBytes = 12 * MC_TMPW * 4 + 4 * chromaRows * MC_PITCH * 2;
  static const bool split = getenv("HMR_MC_SPLIT") != NULL;               // A/B switch: luma and chroma as two launches always
  if (P.hdr.chroma_format == HMR_CHROMA_420 && chromaBytes <= lumaBytes && !split)
  {
    mc_kernel_420<<<dim3(grid, 2), MC_WARPS * 32, MC_WARPS * lumaBytes, s>>>(P, lumaBytes, chromaRows);
    return 2;
  }
  mc_kernel<true><<<grid, MC_WARPS * 32, MC_WARPS * lumaBytes, s>>>(P, lumaBytes, chromaRows);
  if (P.hdr.chroma_format == HMR_CHROMA_400) return 2;
  mc_kernel<false><<<grid, MC_WARPS * 32, MC_WARPS * chromaBytes, s>>>(P, chromaBytes, chromaRows);
  return 3;
}
