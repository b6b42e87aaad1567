// k_resid.cu — batched dequantisation + inverse transform over all TUs of a picture, fused with the
// reconstruction add for inter TUs.
//
// Replaces TComTrQuant::invTransformNxN (TComTrQuant.cpp:1423-1548): xDeQuant flat path (:1203-1313),
// xIT -> xITrMxN -> partialButterflyInverse4/8/16/32 + fastInverseDst (:437-491,539-573,638-685,765-828,894-948),
// xITransformSkip (:1920-1959), transquant-bypass copy (:1475-1487), invRdpcmNxN (:1737-1792),
// crossComponentPrediction (:3294-3335), and TComYuv::addClip (TComYuv.cpp:264-299) for inter CUs.
//
// ONE launch per picture: the host groups the TU records by size, and a CTA finds its size class from four block-count
// prefixes.  A group of N threads owns one NxN TU:
// stage 1: thread j transforms coefficient column j (all N levels of the column are loaded up front so that the loads
//          overlap; dequantising on load; coefficient rows that are zero across the whole warp are skipped);
// stage 2: thread y transforms row y of the transposed intermediate.  The two stages exchange the N x N
// intermediate through shared memory (pitch N+2 -> conflict-free both ways).  Integer MACs in registers; the
// HEVC matrices live in __constant__ memory and are always addressed uniformly across a warp.
#include "common.cuh"

#define RS_THREADS 128

// 32-bit entries: an IMAD then takes the matrix element straight from the constant bank (c[bank][imm]); with int16 entries
// every MAC paid a uniform load plus a sign-extending permute on top (3 instructions per MAC, 190 KB of code)
#define RS_PARTS(lg) ((lg) == 5 ? 4 : ((lg) == 4 ? 2 : 1))   // threads per column of a TU (see resid_block)
#ifndef RS_COMPACT_FROM
#define RS_COMPACT_FROM 4            // log2 transform size from which stage 1 runs as a loop (6 = never)
#endif
__constant__ int c_T32[32][32];
static const int16_t h_cosTab[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                      61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
__constant__ int c_dst4[4][4] = { {29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29} };
__constant__ int c_invq[6] = { 40, 45, 51, 57, 64, 72 };
static bool g_resid_tables_uploaded[64] = {false};

// HEVC core transform matrix (TComRom.cpp:335-484): entry (k, n) = +-cos-table[(2n+1)k folded to the first quadrant]
static void upload_tables(int device)
{
  if (device < 64 && g_resid_tables_uploaded[device]) return;
  static int T[32][32];
  for (int k = 0; k < 32; k++)
    for (int n = 0; n < 32; n++)
    {
      if (k == 0) { T[k][n] = 64; continue; }
      int m = ((2 * n + 1) * k) & 127;
      if (m > 64) m = 128 - m;
      T[k][n] = m > 32 ? -h_cosTab[64 - m] : h_cosTab[m];
    }
  cudaMemcpyToSymbol(c_T32, T, sizeof(T));
  if (device < 64) g_resid_tables_uploaded[device] = true;
}

template <int LOG2N>
__device__ __forceinline__ void resid_block(const FrameParams& P, int16_t* s_all, const uint32_t first, const uint32_t count, const uint32_t block, const int phase)
{
  // PARTS threads share a column (stage 1) / row (stage 2) of a large TU, KP of its N outputs each: a 32x32 TU is four warps' work, not
  // one warp's (a lone 32x32 TU in a B picture was a 20-30 us serial tail of the launch), and nobody holds more than 16 accumulators
  constexpr int N = 1 << LOG2N, NN = N * N, LD = N + 2, PARTS = RS_PARTS(LOG2N), KP = N / PARTS, PER_BLOCK = RS_THREADS / (N * PARTS), STEP = 32 / N;
  int16_t (*s_buf)[N * LD] = (int16_t (*)[N * LD])s_all;
  const int g = threadIdx.x / (N * PARTS);   // TU slot inside the CTA
  const int j = threadIdx.x % N;             // my column (stage 1) / row (stage 2) / column again (store)
  const int k0 = ((threadIdx.x / N) % PARTS) * KP;   // my outputs k0 .. k0 + KP - 1 (stage 1 / 2), my rows in the column-wise passes
  const uint32_t idx = block * PER_BLOCK + g;
  bool active = idx < count;
  hmr_tu t;
  if (active)
  {
    t = P.tu[first + idx];
    // phase 0: everything that does not need a luma residual; phase 1: chroma TUs with cross-component prediction
    if (phase >= 0 && ((t.ccp_alpha != 0) != (phase == 1))) active = false;
  }
  int16_t* sb = s_buf[g];
  const int bd = active ? (t.comp ? P.hdr.bit_depth_chroma : P.hdr.bit_depth_luma) : 8;
  const int16_t* __restrict__ lev = active ? P.coef + t.coef_off : nullptr;
  const bool coded = active && (t.flags & HMR_TU_CODED);
  const bool plainTransform = coded && !(t.flags & (HMR_TU_BYPASS | HMR_TU_TSKIP));

  int acc[KP];
  // ---------------- stage 1 (or the non-transform paths, written straight into residual layout [y][x]) ----------------
  if (plainTransform)
  {
    const int per = t.qp / 6, rem = t.qp - per * 6, scale = c_invq[rem];
    // scaling lists (xDeQuant, TComTrQuant.cpp:1230-1276): per-coefficient factor, 4 more bits of right shift
    const uint8_t* __restrict__ sl = P.scaling ? P.scaling + HMR_SCALING_OFFSET(LOG2N - 2) + (((t.flags & HMR_TU_INTRA) ? 0 : 3) + t.comp) * NN : nullptr;
    const int rshift = 6 - ((15 - bd - LOG2N) + per) + (sl ? 4 : 0);
    const int inBits = min(16, 32 + rshift - (sl ? 15 : 7));
    const int inMin = -(1 << (inBits - 1)), inMax = (1 << (inBits - 1)) - 1;
    const bool dst = (LOG2N == 2) && (t.flags & HMR_TU_DST);
#pragma unroll
    for (int k = 0; k < KP; k++) acc[k] = 0;
    if (LOG2N >= RS_COMPACT_FROM)
    {
      // 16- and 32-point transforms: the row loop is a LOOP over chunks of 4 rows (matrix rows through a register-indexed constant load),
      // not N x N multiply-adds of straight-line code — 16 KB for the 32-point stage alone, which every SM had to fetch cold: a B picture
      // with a single 32x32 TU spent 30 of its 38 us there, and the I picture's launch (IPC 1.7) was short of instructions, not of ALUs.
      // The levels of the next chunk are in flight while this one is multiplied.
      constexpr int CH = 4;
      int qa[CH], qb[CH];
#pragma unroll
      for (int i = 0; i < CH; i++) { qa[i] = __ldg(lev + i * N + j); qb[i] = 0; }
#pragma unroll 1
      for (int n0 = 0; n0 < N; n0 += CH)
      {
        if (n0 + CH < N)
        {
#pragma unroll
          for (int i = 0; i < CH; i++) qb[i] = __ldg(lev + (n0 + CH + i) * N + j);
        }
#pragma unroll
        for (int i = 0; i < CH; i++)
        {
          const int q = qa[i], n = n0 + i;
          if (q != 0)
          {
            const int qc = clip3i(inMin, inMax, q);
            const int sc = sl ? scale * (int)__ldg(sl + n * N + j) : scale;
            int c = rshift > 0 ? (qc * sc + (1 << (rshift - 1))) >> rshift : (int)((unsigned)(qc * sc) << (-rshift));
            c = clip3i(-32768, 32767, c);
            const int* __restrict__ row = c_T32[n * STEP] + k0;
#pragma unroll
            for (int k = 0; k < KP; k++) acc[k] += row[k] * c;
          }
        }
#pragma unroll
        for (int i = 0; i < CH; i++) qa[i] = qb[i];
      }
    }
    else
    {
    int lv[N];
#pragma unroll
    for (int n = 0; n < N; n++) lv[n] = __ldg(lev + n * N + j);
#pragma unroll
    for (int n = 0; n < N; n++)
    {
      const int q = lv[n];
      if (q == 0) continue;
      const int qc = clip3i(inMin, inMax, q);
      const int sc = sl ? scale * (int)__ldg(sl + n * N + j) : scale;
      int c = rshift > 0 ? (qc * sc + (1 << (rshift - 1))) >> rshift : (int)((unsigned)(qc * sc) << (-rshift));
      c = clip3i(-32768, 32767, c);
#pragma unroll
      for (int k = 0; k < N; k++)
      {
        const int m = (LOG2N == 2 && dst) ? c_dst4[n][k] : c_T32[n * STEP][k];
        acc[k] += m * c;
      }
    }
    }
    // tmp[j][k] = clip16((acc + 64) >> 7)
#pragma unroll
    for (int k = 0; k < KP; k++) sb[j * LD + k0 + k] = (int16_t)clip3i(-32768, 32767, (acc[k] + 64) >> 7);
  }
  __syncthreads();
  if (plainTransform)
  {
    const bool dst = (LOG2N == 2) && (t.flags & HMR_TU_DST);
#pragma unroll
    for (int k = 0; k < KP; k++) acc[k] = 0;
    for (int n = 0; n < N; n++)
    {
      const int c = sb[n * LD + j];          // tmp[n][y = j]
      if (c == 0) continue;
#pragma unroll
      for (int k = 0; k < KP; k++)
      {
        const int m = (LOG2N == 2 && dst) ? c_dst4[n][k] : c_T32[n * STEP][k0 + k];
        acc[k] += m * c;
      }
    }
  }
  __syncthreads();                           // everybody has read the intermediate; the tile now becomes resi[y][x]
  if (plainTransform)
  {
    const int shift2 = 20 - bd, rnd2 = 1 << (shift2 - 1);
#pragma unroll
    for (int k = 0; k < KP; k++) sb[j * LD + k0 + k] = (int16_t)clip3i(-32768, 32767, (acc[k] + rnd2) >> shift2);   // row y = j
  }
  else if (active)
  {
    // thread j fills column j of resi[y][x]
    if (!coded) { for (int y = k0; y < k0 + KP; y++) sb[y * LD + j] = 0; }
    else if (t.flags & HMR_TU_BYPASS)
    {
      for (int y = k0; y < k0 + KP; y++)
      {
        const int i = y * N + j;
        sb[y * LD + j] = lev[(t.flags & HMR_TU_ROTATE) ? NN - 1 - i : i];
      }
    }
    else   // transform skip: dequantise, then (c + rnd) >> tsShift
    {
      const int per = t.qp / 6, rem = t.qp - per * 6, scale = c_invq[rem];
      const int trShift = 15 - bd - LOG2N;
      // getUseScalingList: a transform-skipped block uses the lists only when it is 4x4
      const uint8_t* __restrict__ sl = (P.scaling && LOG2N == 2) ? P.scaling + HMR_SCALING_OFFSET(0) + (((t.flags & HMR_TU_INTRA) ? 0 : 3) + t.comp) * NN : nullptr;
      const int rshift = 6 - (trShift + per) + (sl ? 4 : 0);
      const int inBits = min(16, 32 + rshift - (sl ? 15 : 7));
      const int inMin = -(1 << (inBits - 1)), inMax = (1 << (inBits - 1)) - 1;
      for (int y = k0; y < k0 + KP; y++)
      {
        const int i = y * N + j;
        const int src = (t.flags & HMR_TU_ROTATE) ? NN - 1 - i : i;
        const int qc = clip3i(inMin, inMax, (int)lev[src]);
        const int sc = sl ? scale * (int)sl[src] : scale;
        int c = rshift > 0 ? (qc * sc + (1 << (rshift - 1))) >> rshift : (int)((unsigned)(qc * sc) << (-rshift));
        c = clip3i(-32768, 32767, c);
        const int r = trShift >= 0 ? (c + (trShift == 0 ? 0 : (1 << (trShift - 1)))) >> trShift : c << (-trShift);
        sb[y * LD + j] = (int16_t)r;
      }
    }
  }
  __syncthreads();
  // ---------------- RDPCM: running sums down columns / along rows, Pel (int16) wrap-around ----------------
  if (active && k0 == 0 && (t.flags & HMR_TU_RDPCM_V))        // (serial along the column / row: the first of the PARTS threads)
    for (int y = 1; y < N; y++) sb[y * LD + j] = (int16_t)(sb[y * LD + j] + sb[(y - 1) * LD + j]);
  if (active && k0 == 0 && (t.flags & HMR_TU_RDPCM_H))
    for (int x = 1; x < N; x++) sb[j * LD + x] = (int16_t)(sb[j * LD + x] + sb[j * LD + x - 1]);
  __syncthreads();
  if (!active) return;
  // ---------------- cross-component prediction, store / reconstruct: thread j owns column j ----------------
  const bool ccp = t.ccp_alpha != 0 && t.luma_off != HMR_NO_OFFSET;
  const int diffBd = P.hdr.bit_depth_luma - P.hdr.bit_depth_chroma;
  const bool keep = (t.flags & HMR_TU_INTRA) || (t.comp == 0 && (P.hdr.flags & HMR_FRM_HAS_CCP));
  const bool add = !(t.flags & HMR_TU_INTRA);
  const int maxv = (1 << bd) - 1;
  int16_t* plane = P.work.p[t.comp];
  const int pitch = P.work.pitch[t.comp];
  int16_t* d0 = plane + (size_t)t.y * pitch + t.x + j;
  // loads of a batch of rows first (prediction samples, luma residuals for CCP): independent loads in flight instead of
  // one load -> store round trip per row
  constexpr int BATCH = KP < 8 ? KP : 8;
#pragma unroll 1
  for (int y0 = k0; y0 < k0 + KP; y0 += BATCH)
  {
    int cur[BATCH], lum[BATCH];
#pragma unroll
    for (int i = 0; i < BATCH; i++)
    {
      cur[i] = add ? (int)d0[(size_t)(y0 + i) * pitch] : 0;
      lum[i] = ccp ? (int)P.resid[t.luma_off + (y0 + i) * N + j] : 0;
    }
#pragma unroll
    for (int i = 0; i < BATCH; i++)
    {
      const int y = y0 + i;
      int r = sb[y * LD + j];
      if (ccp)
      {
        const int ls = diffBd >= 0 ? (lum[i] >> diffBd) : (lum[i] << (-diffBd));
        r = (int16_t)(r + ((t.ccp_alpha * ls) >> 3));
      }
      if (keep) P.resid[t.coef_off + y * N + j] = (int16_t)r;
      if (add) d0[(size_t)y * pitch] = (int16_t)clip3i(0, maxv, cur[i] + r);
    }
  }
}

struct ResidGrid { uint32_t first[4], count[4], blockEnd[4]; };   // per size class 4x4 .. 32x32

__global__ void __launch_bounds__(RS_THREADS) resid_kernel(const __grid_constant__ FrameParams P, const ResidGrid G, const int phase)
{
  __shared__ __align__(16) int16_t s_all[(RS_THREADS / 32) * 32 * 34];     // largest user: 4 TUs of 32x32, pitch 34
  const uint32_t b = blockIdx.x;
  if (b < G.blockEnd[0])      resid_block<2>(P, s_all, G.first[0], G.count[0], b, phase);
  else if (b < G.blockEnd[1]) resid_block<3>(P, s_all, G.first[1], G.count[1], b - G.blockEnd[0], phase);
  else if (b < G.blockEnd[2]) resid_block<4>(P, s_all, G.first[2], G.count[2], b - G.blockEnd[1], phase);
  else                        resid_block<5>(P, s_all, G.first[3], G.count[3], b - G.blockEnd[2], phase);
}

int launch_resid(const FrameParams& P, cudaStream_t s)
{
  int dev = 0;
  cudaGetDevice(&dev);
  upload_tables(dev);
  ResidGrid G;
  uint32_t blocks = 0;
  for (int k = 0; k < 4; k++)
  {
    G.first[k] = P.hdr.tu_first[k];
    G.count[k] = P.hdr.tu_first[k + 1] - P.hdr.tu_first[k];
    const uint32_t perBlock = RS_THREADS / ((4u << k) * RS_PARTS(k + 2));
    blocks += (G.count[k] + perBlock - 1) / perBlock;
    G.blockEnd[k] = blocks;
  }
  if (!blocks) return 0;
  if (P.hdr.flags & HMR_FRM_HAS_CCP)
  {
    // phase 0: all luma (and non-CCP chroma) TUs of every size; phase 1: the chroma TUs that read luma residuals
    resid_kernel<<<blocks, RS_THREADS, 0, s>>>(P, G, 0);
    resid_kernel<<<blocks, RS_THREADS, 0, s>>>(P, G, 1);
    return 2;
  }
  resid_kernel<<<blocks, RS_THREADS, 0, s>>>(P, G, -1);
  return 1;
}
