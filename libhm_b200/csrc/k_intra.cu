// k_intra.cu — intra prediction + residual add in a CTU-row dependency wavefront.
//
// Replaces TDecCu::xReconIntraQT / xIntraRecQT / xIntraRecBlk (TDecCu.cpp:483-732),
// TComPrediction::initAdiPatternChType + fillReferenceSamples (TComPattern.cpp:107-520, reference-sample
// substitution and [1 2 1] / strong smoothing) and predIntraAng / xPredIntraAng / xPredIntraPlanar /
// predIntraGetPredValDC / xDCPredFiltering (TComPrediction.cpp:182-491, 746-835).
//
// Dependencies: an intra TU reads unfiltered reconstructed samples left / above / above-right / below-left of
// itself.  Inter samples are final before this kernel starts (k_mc + k_resid); intra samples are produced here in
// decode order.  One persistent CTA per (component, CTU row) job, launched cooperatively so that all CTAs are
// co-resident; row r may process CTU c once row r-1 has published c+2 finished CTUs (the above-right CTU) — the
// WPP dependency.  Progress counters carry an epoch so they never need clearing; a row publishes "all CTUs before
// my next CTU that has intra TUs", so rows/CTUs without intra blocks cost nothing.
//
// The per-TU chain is the critical path (an I picture at 2160p is ~1560 dependent luma TUs per CTU row; a dataflow simulation over
// real streams shows TU-level parallelism inside a CTU is < 1.2x, z-order makes every TU depend on its predecessor), and a warp
// walking it is latency-bound (one dependent instruction every ~4.5 cycles, measured with -DINTRA_PROFILE / tools/intra_profile.py),
// so the design keeps everything that does not need the PREVIOUS TU's samples off the chain:
//   * a pre-pass kernel (intra_prep_kernel, fully parallel over the picture) turns the records into reference-address
//     tables — for every TU, entry i = shared-memory position of reference sample i AFTER HM's substitution of
//     unavailable samples (pure function of the record, no sample data) — and into decoded 16-byte micro-ops;
//   * the CTA is warp-specialised and double-buffered: two stager warps do everything that touches global memory for CTU n+1
//     while the four CHAIN warps predict CTU n — the CTU's current samples plus the row above (x = -1 .. CTU+31, after waiting
//     for the row above) and the column to the left, the residuals of its TUs and its decoded TUs / address tables (TMA bulk
//     copies for the contiguous spans, 16-byte cp.async for the tile rows, all in flight together) — then write CTU n back and
//     publish the progress; hand-over through named barriers (bar.arrive / bar.sync), never a full __syncthreads;
//   * every predictor is evaluated as a TAP PROGRAM (tap_prep / tap_turn below): per sample up to four reference samples with
//     weights.  The four chain warps take the 4x4 / 8x8 TUs of a CTU round robin: a warp PREPARES its TU (tile addresses of the
//     taps through the address table, weights, destination, residual: no sample values needed) while the three TUs before it have
//     their turns, waits for the token, takes its TURN — loads, multiply-adds, + residual, clip, store — and passes the token on
//     (bar.arrive -> bar.sync of the next warp: ~95 cycles from "stored" to "first sample of the next TU loaded").  16x16 / 32x32
//     TUs are predicted by all four warps together, a quarter of the samples each;
//   * TUs whose reference line is smoothed first ([1 2 1] / strong filter) deposit the filtered line behind the tiles (line_phase:
//     three loads per entry, the four warps sharing the entries of a large TU) and their taps address that line instead;
//   * a tile is written back once, coalesced.  Cross-CTA reads go through L2 (ld.global.cg); the producer fences before
//     publishing;
//   * shared memory is sized by the largest CTU of the picture (records, address-table entries, residual span: measured
//     on the host records, intra_sizes_host): a resident CTA holds its buffers for the whole wavefront while issuing
//     almost nothing, and with several bitstreams on one GPU that footprint is what the other streams' kernels wait for;
//   * a CTA takes (component, row) jobs in ascending order, so pictures with more rows than co-resident CTAs still run.
// Measured (2160p Main10 I picture, 139 410 intra TUs): 1.50 ms against 2.17 ms for the single-chain-warp version of round 1; the
// rest is the wavefront itself: 34 rows, each ~26 us behind the one above (two CTUs + hand-over), 7 us per luma CTU.
#include <cstdlib>
#include "common.cuh"

#define IN_CHAIN_WARPS 4                    // warps that share the TUs of a CTU (round robin; large TUs by all four together)
#define IN_CHAIN (IN_CHAIN_WARPS * 32)
#define IN_STAGERS 64                      // threads that do everything touching global memory
#define IN_THREADS (IN_CHAIN + IN_STAGERS)
#define IN_MAXCT 64
#define IN_LD (8 + IN_MAXCT + 32)          // tile pitch: 8 columns of left margin (x = -1 lives at column 7), CTU, 32 above-right
#define IN_TILE ((IN_MAXCT + 1) * IN_LD)   // row 0 = y -1; element 0 (y = -1, x = -8) holds the "nothing available" constant
#define IN_MAXREC 256                      // intra records of one CTU and component (64x64 in 4x4 blocks)
#define IN_ADDR (16 * 16 * 17)             // address-table entries per CTU: 17 per 4x4 block, a TU owns the slots of its first block row
#define IN_MAXCOLS 512                     // CTU columns per picture row whose record ranges are cached (8192 / 16)
#define TIDX(y, x) (((y) + 1) * IN_LD + 8 + (x))

__constant__ int c_angTab[9] = { 0, 2, 5, 9, 13, 17, 21, 26, 32 };
__constant__ int c_invTab[9] = { 0, 4096, 1638, 910, 630, 482, 390, 315, 256 };

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- TMA bulk copies (cp.async.bulk, the 1-D mode of the tensor memory accelerator) for the contiguous spans a CTU needs:
// decoded micro-ops, reference-address tables, residuals.  One thread arms an mbarrier with the byte count and issues the
// copies; everybody who reads the data waits on the barrier's phase.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}\n"
               :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// first column >= from whose CTU has intra records of this component (warp-convergent), n if none
__device__ __forceinline__ int next_intra_ctu(const uint16_t* cnt, int from, int n, int lane)
{
  for (int b = from; b < n; b += 32)
  {
    const unsigned m = __ballot_sync(0xffffffffu, b + lane < n && cnt[b + lane] != 0);
    if (m) return b + __ffs(m) - 1;
  }
  return n;
}

struct IntraGeom { int ox, oy, CTW, CTH, uws, uhs, gw; };   // CTU origin / size in component samples, log2 unit size, 4x4 blocks per CTU row

__device__ __forceinline__ int intra_slot(const hmr_intra& r, const IntraGeom& g) { return ((((r.y - g.oy) >> 2) * g.gw) + ((r.x - g.ox) >> 2)) * 17; }

// Reference-sample positions of one TU with HM's substitution (TComPattern.cpp:309-520) resolved: entry i of the line
// ([0] bottom-most below-left ... [2N] corner ... [4N] last above-right) = tile index to read.  Unit-granular availability.
__device__ __forceinline__ void intra_addr_table(const hmr_intra& r, uint16_t* __restrict__ addr, const IntraGeom& g, int lane)
{
  const int N = 1 << r.log2_size, N2 = 2 * N, L = 4 * N + 1;
  const int x0 = r.x - g.ox, y0 = r.y - g.oy;
  const int nl = N >> g.uhs, na = N >> g.uws;
  unsigned long long M = 0;                               // bit u = unit u available, units in line order
  M |= (unsigned long long)(__brev((unsigned)r.avail_below_left) >> (32 - nl));
  M |= (unsigned long long)(__brev((unsigned)r.avail_left) >> (32 - nl)) << nl;
  if (r.flags & HMR_INTRA_AVAIL_CORNER) M |= 1ull << (2 * nl);
  M |= (unsigned long long)r.avail_above << (2 * nl + 1);
  M |= (unsigned long long)r.avail_above_right << (2 * nl + 1 + na);
  const int q = __ffsll((long long)M) - 1;                // first available unit
  const int qsrc = q < 2 * nl ? (q << g.uhs) : (q == 2 * nl ? N2 : N2 + 1 + ((q - 2 * nl - 1) << g.uws));
  for (int i = lane; i < L; i += 32)
  {
    int a = 0;                                            // nothing available: the constant slot
    if (M)
    {
      const int u = i < N2 ? (i >> g.uhs) : (i == N2 ? 2 * nl : 2 * nl + 1 + ((i - N2 - 1) >> g.uws));
      int src = i;
      if (!((M >> u) & 1))
      {
        const unsigned long long lower = M & ((1ull << u) - 1);
        if (lower)
        {
          const int p = 63 - __clzll((long long)lower);   // nearest available unit before: its LAST sample
          src = p < 2 * nl ? ((p + 1) << g.uhs) - 1 : (p == 2 * nl ? N2 : N2 + ((p - 2 * nl) << g.uws));
        }
        else src = qsrc;                                  // first available unit after: its FIRST sample
      }
      if (src < N2)       a = TIDX(min(y0 + N2 - 1 - src, g.CTH - 1), x0 - 1);
      else if (src == N2) a = TIDX(y0 - 1, x0 - 1);
      else                a = TIDX(y0 - 1, min(x0 + (src - N2 - 1), g.CTW + 31));
    }
    addr[i] = (uint16_t)a;
  }
}

// Everything the chain needs to know about a TU, decoded ahead of time by the stagers (16 bytes, one LDS.128).
// Kept as a plain uint4 (fields by shifts): a struct with narrow members ends up on the local-memory stack, and local memory
// sits behind the L1 that every __threadfence of the stager warps invalidates.
//   x: [15:0] org  = tile index of the TU's top-left sample     [31:16] slot = first entry of its reference-address table
//   y: res  = residual position relative to the CTU's first residual, or HMR_NO_OFFSET
//   z: [7:0] angle (signed intraPredAngle)  [15:8] log2 size  [23:16] class OP_*  [31:24] flags OPF_*
//   w: [15:0] inverse angle (negative angles)
typedef uint4 IntraOp;
__device__ __forceinline__ int op_org(const IntraOp& o)   { return o.x & 0xffff; }
__device__ __forceinline__ int op_slot(const IntraOp& o)  { return o.x >> 16; }
__device__ __forceinline__ int op_angle(const IntraOp& o) { return (int)(int8_t)(o.z & 0xff); }
__device__ __forceinline__ int op_lg(const IntraOp& o)    { return (o.z >> 8) & 0xff; }
__device__ __forceinline__ int op_cls(const IntraOp& o)   { return (o.z >> 16) & 0xff; }
__device__ __forceinline__ int op_flags(const IntraOp& o) { return o.z >> 24; }
__device__ __forceinline__ int op_inv(const IntraOp& o)   { return o.w & 0xffff; }
enum { OP_PLANAR = 0, OP_DC = 1, OP_ANG0 = 2, OP_ANGPOS = 3, OP_ANGNEG = 4, OP_PCM = 5 };
enum { OPF_FILTER = 1, OPF_STRONG = 2, OPF_EDGE = 4, OPF_VER = 8, OPF_DCEDGE = 16 };

__device__ __forceinline__ IntraOp intra_make_op(const hmr_intra& r, const IntraGeom& g, unsigned minoff, bool strongAllowed, int slot)
{
  const int N = 1 << r.log2_size;
  const bool luma = r.flags & HMR_INTRA_LUMA_RULES;
  int f = 0, cls, angle = 0, inv = 0;
  if (r.flags & HMR_INTRA_FILTER_REFS) f |= OPF_FILTER;
  if (N == 32 && luma && strongAllowed) f |= OPF_STRONG;
  if (luma && N <= 16) f |= OPF_DCEDGE;
  if (luma && N <= 16 && !(r.flags & HMR_INTRA_NO_EDGE_FLT)) f |= OPF_EDGE;
  const int mode = r.mode;
  if (mode == 0) cls = OP_PLANAR;
  else if (mode == 1) cls = OP_DC;
  else if (mode == HMR_INTRA_MODE_PCM) cls = OP_PCM;       // I_PCM: prediction 0, the samples arrive as the residual
  else
  {
    const bool ver = mode >= 18;
    const int am = ver ? mode - 26 : 10 - mode;
    const int aa = abs(am);
    angle = am < 0 ? -c_angTab[aa] : c_angTab[aa];
    if (ver) f |= OPF_VER;
    inv = c_invTab[aa];
    cls = angle == 0 ? OP_ANG0 : (angle > 0 ? OP_ANGPOS : OP_ANGNEG);
  }
  IntraOp op;
  op.x = (uint32_t)TIDX(r.y - g.oy, r.x - g.ox) | ((uint32_t)slot << 16);
  op.y = r.resid_off != HMR_NO_OFFSET ? r.resid_off - minoff : HMR_NO_OFFSET;
  op.z = (uint32_t)(angle & 0xff) | ((uint32_t)r.log2_size << 8) | ((uint32_t)cls << 16) | ((uint32_t)f << 24);
  op.w = (uint32_t)inv;
  return op;
}

#define IN_MAX_CTAS 36  // CTAs of one wavefront launch (see intra_kernel; sweep on 2160p: 36 = 102 for a single stream, +14 % with 8 streams; HMR_INTRA_CTAS overrides)
#define IN_NJMAX 5      // (4*32 + 1 + 31) / 32 reference samples per lane at most
#define IN_LINE 144     // int16 entries of the filtered reference line of one TU (4*32 + 1, plus the -1 / 4N+1 slots the 45-degree modes touch with weight 0)

// ---- a TU as a TAP PROGRAM --------------------------------------------------------------------------------------------------
// Every HM predictor is, per sample, a weighted sum of at most four reference samples (planar: LEFT(y), TOP(N), TOP(x), LEFT(N);
// angular: the two neighbours on the main reference, the side reference projected in for negative angles; DC: the mean plus one
// or two edge samples; pure horizontal / vertical: one sample plus the edge correction).  PREP (tap_prep: needs no sample values,
// runs while other TUs have their turn) resolves, per lane and sample, WHERE those four samples are and their weights, the
// destination and the residual.  "Where" is a tile index (0 = the constant slot, for taps of weight 0): either the reference sample itself, looked up in the TU's address table
// (HM's substitution of unavailable samples is already resolved there), or — TUs whose reference line is smoothed first
// (TComPrediction::filteringIntraReferenceSamples) — an entry of the filtered line the owner of the TU deposits behind the tiles
// (line_phase).  The TURN (tap_turn) is then loads, multiply-adds, + residual, clip, store: the only part of a TU that has to
// wait for its predecessor.  4x4 / 8x8 TUs are one warp's job (1 / 2 samples per lane); 16x16 / 32x32 TUs are shared by the four
// chain warps (2 / 8 samples per lane each).
enum { TV_GENERIC = 0, TV_DC = 1, TV_HVEDGE = 2 };
template <int S> struct TapProg
{
  int aA[S], aB[S], aC[S], aD[S];               // tile indices (the filtered line lives at lineBase + i)
  int wA[S], wB[S], wC[S], wD[S];
  int dst[S], res[S];
};
// samples per lane and round: 4x4 = 16 samples (half a warp), 8x8 = 64 (one warp, 2 per lane), 16x16 = 4 warps x 64, 32x32 = 4 rounds of 4 warps x 64
// (a program of 8 samples per lane would cost 80 registers; the CTA's register footprint is what other bitstreams' kernels wait for)
template <int LG> struct TapShape { static constexpr int S = LG == 2 ? 1 : 2; static constexpr int ROUNDS = LG == 5 ? 4 : 1; };

// Program of samples first + lane + 32 j (j < S) of the TU.  rnd / sh: rounding and shift of the weighted sum (TV_GENERIC).
// One loop per predictor class (the class is uniform over the TU): a sample costs a handful of integer operations plus the table
// look-ups of its taps.  Taps a class does not use keep weight 0 and point at entry 0 of the tile (the constant slot): always readable.
template <int LG>
__device__ __forceinline__ void tap_prep(const IntraOp op, const uint16_t* __restrict__ addrTab, const int16_t* __restrict__ resB, const int lineBase,
                                         const int first, const int lane, TapProg<TapShape<LG>::S>& g, int& variant, int& rnd, int& sh)
{
  constexpr int N = 1 << LG, N2 = 2 * N, S = TapShape<LG>::S;
  const uint16_t* t = addrTab + op_slot(op);
  const int cls = op_cls(op), flags = op_flags(op);
  const bool filtered = flags & OPF_FILTER;                  // taps read the smoothed line, not the tile
  const bool hasRes = op.y != HMR_NO_OFFSET;
  const int org = op_org(op);
  const int16_t* res = resB + (hasRes ? op.y : 0u);
  auto at = [&](int idx) -> int { return filtered ? lineBase + idx : (int)t[idx]; };
  variant = TV_GENERIC; rnd = 0; sh = 0;
#pragma unroll
  for (int j = 0; j < S; j++)
  {
    const int i = (first + lane + 32 * j) & (N * N - 1);     // 4x4: lanes 16-31 shadow lanes 0-15 (they never store)
    g.aA[j] = g.aB[j] = g.aC[j] = g.aD[j] = 0;
    g.wA[j] = g.wB[j] = g.wC[j] = g.wD[j] = 0;
    g.dst[j] = org + (i >> LG) * IN_LD + (i & (N - 1));
    g.res[j] = hasRes ? (int)res[i] : 0;
  }
  if (cls == OP_PLANAR)
  {
    rnd = N; sh = LG + 1;
    const int aTN = at(N2 + 1 + N), aLN = at(N2 - 1 - N);    // TOP(N), LEFT(N): the same for every sample
#pragma unroll
    for (int j = 0; j < S; j++)
    {
      const int i = (first + lane + 32 * j) & (N * N - 1), y = i >> LG, x = i & (N - 1);
      g.aA[j] = at(N2 - 1 - y); g.wA[j] = N - 1 - x;       // LEFT(y)
      g.aB[j] = aTN;            g.wB[j] = x + 1;
      g.aC[j] = at(N2 + 1 + x); g.wC[j] = N - 1 - y;       // TOP(x)
      g.aD[j] = aLN;            g.wD[j] = y + 1;
    }
  }
  else if (cls == OP_DC)
  {
    variant = TV_DC;                                         // tap C carries the mean; edge samples: (A + B + wC * dc + 2) >> 2  (TComPrediction.cpp:746-835)
    const bool edge = flags & OPF_DCEDGE;
#pragma unroll
    for (int j = 0; j < S; j++)
    {
      const int i = (first + lane + 32 * j) & (N * N - 1), y = i >> LG, x = i & (N - 1);
      g.wC[j] = 1;
      if (edge && (x == 0 || y == 0))
      {
        g.aA[j] = at(y == 0 ? N2 + 1 + x : N2 - 1 - y); g.wA[j] = 1; g.wC[j] = 3;
        if (x == 0 && y == 0) { g.aB[j] = at(N2 - 1); g.wB[j] = 1; g.wC[j] = 2; }
      }
    }
  }
  else if (cls == OP_ANG0)
  {
    const bool ver = flags & OPF_VER, edge = flags & OPF_EDGE;
    const int sgn = ver ? 1 : -1;
    if (edge) variant = TV_HVEDGE;                           // wB marks the edge samples, C = corner
    const int aCorner = edge ? at(N2) : 0;
#pragma unroll
    for (int j = 0; j < S; j++)
    {
      const int i = (first + lane + 32 * j) & (N * N - 1), y = i >> LG, x = i & (N - 1);
      const int yy = ver ? y : x, xx = ver ? x : y;
      g.aA[j] = at(N2 + sgn * (xx + 1)); g.wA[j] = 1;
      if (edge && xx == 0) { g.aB[j] = at(N2 - sgn * (yy + 1)); g.wB[j] = 1; g.aC[j] = aCorner; }
    }
  }
  else if (cls == OP_ANGPOS || cls == OP_ANGNEG)
  {
    rnd = 16; sh = 5;
    const bool ver = flags & OPF_VER;
    const int sgn = ver ? 1 : -1, angle = op_angle(op), inv = op_inv(op);
#pragma unroll
    for (int j = 0; j < S; j++)
    {
      const int i = (first + lane + 32 * j) & (N * N - 1), y = i >> LG, x = i & (N - 1);
      const int yy = ver ? y : x, xx = ver ? x : y;
      const int pos = (yy + 1) * angle, di = pos >> 5, df = pos & 31;
      const int k0 = xx + di + 1, k1 = k0 + 1;
      // negative angle: rm[k < 0] is the side edge projected onto the main edge: side sample ((128 - k*inv) >> 8) - 1  (TComPrediction.cpp:396-404)
      const int i0 = k0 >= 0 ? N2 + sgn * k0 : N2 - sgn * ((128 - k0 * inv) >> 8);
      const int i1 = k1 >= 0 ? N2 + sgn * k1 : N2 - sgn * ((128 - k1 * inv) >> 8);
      g.aA[j] = at(i0); g.wA[j] = 32 - df;
      g.aB[j] = at(min(4 * N, max(0, i1))); g.wB[j] = df;  // df == 0 at the end of the line: weight 0, any valid entry will do
    }
  }
  // OP_PCM: prediction 0, the samples arrive as the residual
}

// The turn proper.  Specialised per variant outside the sample loop: the common case is four loads and four multiply-adds per sample.
template <int LG>
__device__ __forceinline__ void tap_turn(const TapProg<TapShape<LG>::S>& g, const int variant, const int rnd, const int sh, const int dc,
                                         int16_t* __restrict__ tile, const int bd, const int lane)
{
  constexpr int N = 1 << LG, S = TapShape<LG>::S;
  const int maxv = (1 << bd) - 1;
  int A[S], B[S], C[S], D[S], p[S];
#pragma unroll
  for (int j = 0; j < S; j++) { A[j] = tile[g.aA[j]]; B[j] = tile[g.aB[j]]; C[j] = tile[g.aC[j]]; D[j] = tile[g.aD[j]]; }
  if (variant == TV_GENERIC)
  {
#pragma unroll
    for (int j = 0; j < S; j++) p[j] = (g.wA[j] * A[j] + g.wB[j] * B[j] + g.wC[j] * C[j] + g.wD[j] * D[j] + rnd) >> sh;
  }
  else if (variant == TV_DC)
  {
#pragma unroll
    for (int j = 0; j < S; j++) p[j] = g.wC[j] == 1 ? dc : (A[j] * g.wA[j] + B[j] * g.wB[j] + dc * g.wC[j] + 2) >> 2;
  }
  else
  {
#pragma unroll
    for (int j = 0; j < S; j++) p[j] = g.wB[j] ? clip3i(0, maxv, A[j] + ((B[j] - C[j]) >> 1)) : A[j];
  }
#pragma unroll
  for (int j = 0; j < S; j++)
    if (N * N >= 32 || lane < N * N) tile[g.dst[j]] = (int16_t)clip3i(0, maxv, (int)(int16_t)p[j] + g.res[j]);
}

// Mean of the 2N nearest reference samples (predIntraGetPredValDC, TComPrediction.cpp:182-207); sT / sL: tile index of this lane's TOP / LEFT sample.
template <int LG>
__device__ __forceinline__ int dc_value(const int16_t* __restrict__ tile, const int sT, const int sL, const int lane)
{
  constexpr int N = 1 << LG;
  const int part = (N >= 32 || lane < N) ? (int)tile[sT] + (int)tile[sL] : 0;
  return (__reduce_add_sync(0xffffffffu, part) + N) >> (LG + 1);
}

// Smoothing of the reference line of a filtered TU ([0] bottom-most below-left ... [2N] corner ... [4N] last above-right), in the turn:
// [1 2 1] / 4, or the bilinear "strong" variant of 32x32 luma blocks (TComPattern.cpp:219-306).  Entry i is one lane's job: three loads
// through the TU's address table (fetched ahead of the turn: line_addrs), one store to line[i]; the participating warps split the entries
// (4x4 / 8x8: one warp, 16x16 / 32x32: four).  PER = entries per lane.
template <int LG> struct LineShape { static constexpr int PER = LG == 3 ? 2 : (LG == 5 ? 2 : 1); static constexpr int WARPS = LG >= 4 ? IN_CHAIN_WARPS : 1; };
template <int LG>
__device__ __forceinline__ void line_addrs(const IntraOp op, const uint16_t* __restrict__ addrTab, const int part, const int lane, int a[2][3])
{
  constexpr int N = 1 << LG, L = 4 * N + 1, PER = LineShape<LG>::PER;
  const uint16_t* t = addrTab + op_slot(op);
#pragma unroll
  for (int j = 0; j < PER; j++)
  {
    const int i = min((part * PER + j) * 32 + lane, L - 1);  // lanes past the end repeat the last entry
    const bool end = i == 0 || i == L - 1;                   // the two end points are copied ((v + 2v + v + 2) >> 2 == v)
    a[j][0] = t[end ? i : i - 1]; a[j][1] = t[i]; a[j][2] = t[end ? i : i + 1];
  }
}
template <int LG>
__device__ __forceinline__ void line_phase(const IntraOp op, const uint16_t* __restrict__ addrTab, const int a[2][3], const int part, const int16_t* __restrict__ tile,
                                           int16_t* __restrict__ line, const int bd, const int lane)
{
  constexpr int N = 1 << LG, N2 = 2 * N, L = 4 * N + 1, PER = LineShape<LG>::PER;
  bool strong = false;
  int bl = 0, tl = 0, tr = 0;
  if (N == 32 && (op_flags(op) & OPF_STRONG))
  {
    const uint16_t* t = addrTab + op_slot(op);
    bl = tile[t[0]]; tl = tile[t[N2]]; tr = tile[t[4 * N]];
    const int mid0 = tile[t[N]], mid1 = tile[t[3 * N]];
    const int thr = 1 << (bd - 5);
    strong = abs(bl + tl - 2 * mid0) < thr && abs(tl + tr - 2 * mid1) < thr;
  }
#pragma unroll
  for (int j = 0; j < PER; j++)
  {
    const int i = min((part * PER + j) * 32 + lane, L - 1);
    const int up = tile[a[j][0]], v = tile[a[j][1]], dn = tile[a[j][2]];
    int r = (up + 2 * v + dn + 2) >> 2;
    if (strong && i > 0 && i < 4 * N)
      r = i < N2 ? ((N2 - i) * bl + i * tl + N) >> (LG + 1) : (i == N2 ? tl : ((N2 - (i - N2)) * tl + (i - N2) * tr + N) >> (LG + 1));
    line[i] = (int16_t)r;
  }
}

// ---- pre-pass: everything about the intra TUs that does not depend on sample data, for the whole picture at once ----
// One WARP per (component, CTU), four per CTA: reference-address tables (compact, TU after TU), decoded micro-ops, and the
// CTU's residual span.  The wavefront kernel below only copies these into shared memory.  (A warp, not a CTA, per CTU:
// most CTUs of an inter picture have no intra TU at all and a warp that finds nothing costs next to nothing.)
#define PREP_WARPS 4
__global__ void __launch_bounds__(PREP_WARPS * 32) intra_prep_kernel(const __grid_constant__ FrameParams P)
{
  __shared__ int s_offAll[PREP_WARPS][IN_MAXREC + 1];
  __shared__ uint4 s_recAll[PREP_WARPS][IN_MAXREC];                // the CTU's records (16 bytes each), read from global memory ONCE
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nctu = P.ctus_w * P.ctus_h;
  const int job = blockIdx.x * PREP_WARPS + warp;
  if (blockIdx.x == 0 && threadIdx.x == 0) P.intra_progress[3 * P.ctus_h] = 0ull;     // the job counter of the wavefront kernel that follows in the stream
  if (job >= 3 * nctu) return;
  const int comp = job / nctu, ctu = job - comp * nctu;
  if (comp > 0 && P.hdr.chroma_format == HMR_CHROMA_400) return;
  const uint32_t first = __ldg(&P.irange[ctu].first[comp]);
  const int count = (int)min(__ldg(&P.irange[ctu].count[comp]), (uint32_t)IN_MAXREC);
  if (count == 0) { if (lane == 0) P.intra_prep[job] = make_uint4(0, 0, 0, 0); return; }
  int* s_off = s_offAll[warp];
  uint4* s_rec = s_recAll[warp];
  // all record loads of the CTU in flight together (independent, coalesced); everything below reads shared memory — a warp that
  // fetched record k from global memory inside its TU loop paid one L2 round trip per TU (74 us for a 2160p I picture)
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.intra + first);
#pragma unroll 4
    for (int k = lane; k < count; k += 32) s_rec[k] = __ldg(src + k);
  }
  __syncwarp();
  const int csx = comp ? P.csx : 0, csy = comp ? P.csy : 0;
  IntraGeom g;
  g.CTW = (1 << P.hdr.log2_ctu) >> csx; g.CTH = (1 << P.hdr.log2_ctu) >> csy;
  g.uws = 2 - csx; g.uhs = 2 - csy; g.gw = g.CTW >> 2;
  g.ox = (ctu % P.ctus_w) * g.CTW; g.oy = (ctu / P.ctus_w) * g.CTH;
  // table lengths -> exclusive prefix (warp scan over chunks of 32), residual span (warp min / max)
  unsigned mn = 0xffffffffu, mx = 0;
  int running = 0;
  for (int k0 = 0; k0 < count; k0 += 32)
  {
    const int k = k0 + lane;
    int len = 0;
    if (k < count)
    {
      const hmr_intra r = *reinterpret_cast<const hmr_intra*>(&s_rec[k]);
      len = (4 << r.log2_size) + 1;
      if (r.resid_off != HMR_NO_OFFSET) { mn = min(mn, r.resid_off); mx = max(mx, r.resid_off + (1u << (2 * r.log2_size))); }
    }
    int incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (k < count) s_off[k] = running + incl - len;
    running += __shfl_sync(0xffffffffu, incl, 31);
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  __syncwarp();
  const bool strongAllowed = P.hdr.flags & HMR_FRM_STRONG_INTRA_SMOOTHING;
  uint16_t* tab = P.intra_tab + (size_t)job * IN_ADDR;
  // micro-ops: one lane per TU (16-byte stores, coalesced)
  for (int k = lane; k < count; k += 32)
  {
    const hmr_intra r = *reinterpret_cast<const hmr_intra*>(&s_rec[k]);
    P.intra_ops[first + k] = intra_make_op(r, g, mn, strongAllowed, s_off[k]);
  }
  // address tables: the warp walks the TUs, lanes = table entries
  for (int k = 0; k < count; k++)
  {
    const hmr_intra r = *reinterpret_cast<const hmr_intra*>(&s_rec[k]);
    intra_addr_table(r, tab + s_off[k], g, lane);
  }
  if (lane == 0) P.intra_prep[job] = make_uint4(mn, mx > mn ? mx - mn : 0u, (unsigned)running, 0u);
}

// -DINTRA_PROFILE: cycle counters of the chain warps (clock64 between the phases of chain_tu, per TU size), summed over the launch
// into g_intraProf and read back by hmr_debug_intra_profile (tools/intra_profile.py).  Slot = 8 * (lg - 2) + phase.
#ifdef INTRA_PROFILE
__device__ unsigned long long g_intraProf[64];
__device__ unsigned long long g_intraRows[3 * 128 * 4];     // per (component, row): globaltimer at job start, first CTU staged, last TU done, job end
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PROF_DECL long long pt_ = clock64(), pn_
#define PROF_MARK(phase) do { pn_ = clock64(); prof[8 * (LG - 2) + (phase)] += (unsigned long long)(pn_ - pt_); pt_ = pn_; } while (0)
#define PROF_ARG , unsigned long long* prof
#define PROF_PASS , prof
#else
#define PROF_DECL
#define PROF_MARK(phase)
#define PROF_ARG
#define PROF_PASS
#endif

// named barriers (id 0 is __syncthreads)
#define BAR_FULL 1     // +buffer: the fetcher warp arrives, chain warps wait  -> "CTU staged"
#define BAR_DONE 3     // +buffer: chain warp 0 arrives, the writer warp waits -> "CTU predicted"
#define BAR_CHAIN 6    // the chain warps among themselves: every TU of the CTU is in the tile
#define BAR_TOKEN 7    // + (k & 3): TU k is in the tile, the owner of TU k + 1 may take its turn
#define BAR_COOP 11    // large filtered TU: the four chain warps have each smoothed their part of the reference line
#define BAR_FREE 12    // +buffer: the writer has written the CTU back, the fetcher may stage the next but one into the buffer
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }

// One TU of the chain for this warp.  4x4 / 8x8 TUs belong to warp k & 3; 16x16 / 32x32 TUs are shared by all four chain warps, a
// quarter of the samples (and of the reference line) each.  Order inside: PREP (no sample values) -> token -> smoothing of the line
// (filtered TUs) -> TURN -> pass the token.  Token of TU k = named barrier BAR_TOKEN + (k & 3); who arrives and who waits depends on
// whether TU k and TU k + 1 are small (S) or large (L):
//     S -> S   owner(k) arrives, owner(k + 1) waits                                   (64 threads)
//     S -> L   owner(k) comes straight to the large TU; all four warps wait there     (128)
//     L -> S   the three warps that do not own k + 1 arrive, owner(k + 1) waits       (128)
//     L -> L   all four warps wait at the beginning of TU k + 1                       (128)
// Memory ordering: bar.sync orders like __syncthreads; in front of every bar.arrive there is a CTA-scope fence (see below).
template <int LG>
__device__ __forceinline__ void chain_tu(const IntraOp op, const int k, const int count, const bool prevLarge, const bool nextLarge, const int warp, const int lane,
                                         const uint16_t* __restrict__ addrTab, const int16_t* __restrict__ resB, int16_t* __restrict__ tile,
                                         const int lineBase, const int bd PROF_ARG)
{
  constexpr int N = 1 << LG, N2 = 2 * N;
  constexpr bool large = LG >= 4;
  PROF_DECL;
  const int flags = op_flags(op);
  const bool filtered = flags & OPF_FILTER;
  const int part = large ? warp : 0;
  constexpr int ROUNDS = TapShape<LG>::ROUNDS, PER_ROUND = N * N / ROUNDS;     // samples of one round, split over the participating warps
  TapProg<TapShape<LG>::S> g;
  int variant, rnd, sh, dc = 0;
  tap_prep<LG>(op, addrTab, resB, lineBase, part * (PER_ROUND / 4), lane, g, variant, rnd, sh);
  int la[2][3];
  int sT = 0, sL = 0;
  if (filtered) line_addrs<LG>(op, addrTab, part, lane, la);
  if (variant == TV_DC)
  {
    const uint16_t* t = addrTab + op_slot(op);
    const int l = lane & (N - 1);
    if (filtered) { sT = lineBase + N2 + 1 + l; sL = lineBase + N2 - 1 - l; }      // (HM never smooths for DC; kept consistent with the taps)
    else          { sT = t[N2 + 1 + l]; sL = t[N2 - 1 - l]; }
  }
  PROF_MARK(1);                                              // prep
  if (k > 0) bar_sync(BAR_TOKEN + ((k - 1) & 3), (large || prevLarge) ? IN_CHAIN : 64);          // TU k - 1 is in the tile
  PROF_MARK(2);                                              // token wait (nominal: the blocking is deferred to the first dependent access)
  if (filtered)
  {
    line_phase<LG>(op, addrTab, la, part, tile, tile + lineBase, bd, lane);
    if (large) bar_sync(BAR_COOP, IN_CHAIN); else __syncwarp();
  }
  if (variant == TV_DC) dc = dc_value<LG>(tile, sT, sL, lane);       // large TUs: every warp computes the mean itself
  PROF_MARK(3);                                              // line / DC phase (+ the real token wait)
  tap_turn<LG>(g, variant, rnd, sh, dc, tile, bd, lane);
#pragma unroll 1
  for (int r = 1; r < ROUNDS; r++)                           // 32x32: the remaining three quarters (nothing here reads what this TU writes)
  {
    tap_prep<LG>(op, addrTab, resB, lineBase, r * PER_ROUND + part * (PER_ROUND / 4), lane, g, variant, rnd, sh);
    tap_turn<LG>(g, variant, rnd, sh, dc, tile, bd, lane);
  }
  __syncwarp();
  PROF_MARK(5);                                              // turn
#ifdef INTRA_PROFILE
  prof[8 * (LG - 2)] += 1;
#endif
  if (k + 1 < count && !nextLarge)
  {
    // The stores of this turn must be PERFORMED before the token is passed: bar.arrive signals when it issues, and under load (other
    // CTAs' traffic in the SM's shared-memory pipeline) the next warp's loads have been seen to overtake stores still queued — found by
    // tools/multistream_check.py; a single bitstream never showed it.  The CTA-scope fence costs ~15 cycles here (measured).
    __threadfence_block();
    if (!large) bar_arrive(BAR_TOKEN + (k & 3), 64);
    else if (((k + 1) & 3) != warp) bar_arrive(BAR_TOKEN + (k & 3), IN_CHAIN);   // the next owner is one of us: it waits with count 128
  }
}

// Persistent CTAs take (CTU row, component) jobs from a queue, in the order of intra_job_index (common.cuh: rows ascending per component, luma ahead of chroma).
// A job only waits for the job of the same component one row up, which was handed out earlier to a CTA that is resident (cooperative
// launch): no deadlock however few CTAs there are.  A row lags the one above by two CTUs, so only ~a third of the rows of a 2160p
// picture are ever active at once: the launcher starts IN_MAX_CTAS CTAs, not one per job — what a CTA holds (registers, shared memory)
// while it waits for its turn is what other bitstreams' kernels on the same GPU wait for.
__global__ void __launch_bounds__(IN_THREADS, 3) intra_kernel(const __grid_constant__ FrameParams P, const int resSamples, const int maxRec, const int maxAddr)
{
  extern __shared__ __align__(16) uint8_t s_dyn[];
  constexpr int TILE_PAD = (IN_TILE + 7) & ~7;
  int16_t* s_tileB = (int16_t*)s_dyn;                                        // [2][TILE_PAD], then [IN_LINE]: the filtered reference line of the TU in flight
  int16_t* s_resB = s_tileB + 2 * TILE_PAD + IN_LINE;                        // [2][resSamples] residuals of a CTU, compact layout relative to minoff
  // capacities = the largest CTU of THIS picture (engine.cu measures the records): a resident CTA holds its shared memory for
  // the whole wavefront, and at saturation that footprint is what other streams' kernels wait for
  IntraOp* s_ops = (IntraOp*)(s_resB + 2 * resSamples);                      // [2][maxRec] decoded TUs of a CTU
  uint16_t* s_addr = (uint16_t*)(s_ops + 2 * maxRec);                        // [2][maxAddr]
  uint4* s_prep = (uint4*)(s_addr + 2 * maxAddr);                            // [ctus_w] per CTU of this row: x = first residual, y = residual span, z = table entries
  uint32_t* s_first = (uint32_t*)(s_prep + P.ctus_w);                        // [ctus_w]
  uint16_t* s_count = (uint16_t*)(s_first + P.ctus_w);                       // [ctus_w]
  __shared__ int16_t s_col[IN_MAXCT];     // right-most column of the CTU the chain just finished (left neighbours of the next one)
  __shared__ __align__(8) unsigned long long s_mbar[2];   // one per buffer: completion of the TMA bulk copies
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ctusW = P.ctus_w;
  if (tid == 0)
  {
    mbar_init(&s_mbar[0], 1); mbar_init(&s_mbar[1], 1);
    fence_proxy_async();                                       // the barriers exist before the async proxy touches them
  }
  __shared__ int s_job;
  int n = 0;                                                   // CTUs staged so far by this CTA (all jobs): buffer = n & 1, mbarrier phase = (n >> 1) & 1

  for (;;)
  {
  __syncthreads();                                             // the previous job's readers of the per-row arrays (and of s_job) are gone
  if (tid == 0)
  {
    // next job: the k-th (row, component) that has intra TUs at all (bit mask from the host, which walks the records anyway).  One
    // atomicAdd on a counter the pre-pass zeroed — a compare-and-swap loop on an epoch-tagged counter cost ~0.3 us per job under
    // the contention of 36 CTAs: 30 us for the 102 jobs of a 2160p picture, most of a B picture's launch.
    const int k = (int)atomicAdd((unsigned int*)(P.intra_progress + 3 * P.ctus_h), 1u);
    int jb = 3 * P.ctus_h;
    if (P.intra_jobs < 0) jb = min(k, jb);
    else if (k < P.intra_jobs)
    {
      int left = k;
      for (int w = 0; w < HMR_INTRA_JOB_WORDS; w++)
      {
        const int pc = __popc(P.intra_job_mask[w]);
        if (left < pc) { jb = 32 * w + __fns(P.intra_job_mask[w], 0, left + 1); break; }
        left -= pc;
      }
    }
    s_job = jb;
  }
  __syncthreads();
  const int job = s_job;
  if (job >= 3 * P.ctus_h) break;
  int row, comp;
  intra_job_decode(job, P.ctus_h, row, comp);
  if (comp > 0 && P.hdr.chroma_format == HMR_CHROMA_400) continue;
  unsigned long long* myProg = P.intra_progress + comp * P.ctus_h + row;
  // the row above only has to be waited for if it is a job itself (without intra TUs its samples were final before the launch)
  const int upJob = row > 0 ? intra_job_index(row - 1, comp, P.ctus_h) : 0;
  const bool upIsJob = row > 0 && (P.intra_jobs < 0 || ((P.intra_job_mask[upJob >> 5] >> (upJob & 31)) & 1));
  const volatile unsigned long long* upProg = upIsJob ? P.intra_progress + comp * P.ctus_h + row - 1 : nullptr;
  const unsigned long long base = P.epoch << 32;

  const int bd = comp ? P.hdr.bit_depth_chroma : P.hdr.bit_depth_luma;
  const int csx = comp ? P.csx : 0, csy = comp ? P.csy : 0;
  IntraGeom g;
  g.CTW = (1 << P.hdr.log2_ctu) >> csx; g.CTH = (1 << P.hdr.log2_ctu) >> csy;
  g.uws = 2 - csx; g.uhs = 2 - csy; g.gw = g.CTW >> 2;
  g.oy = row * g.CTH; g.ox = 0;
  const int CTW = g.CTW, CTH = g.CTH;
  const int W = P.w[comp], H = P.h[comp];
  int16_t* plane = P.work.p[comp];
  const int pitch = P.work.pitch[comp];
  const int oy = row * CTH, ch = min(CTH, H - oy);

  for (int c = tid; c < ctusW; c += IN_THREADS)
  {
    const hmr_ctu_intra_range* rg = P.irange + row * ctusW + c;
    s_first[c] = rg->first[comp];
    s_count[c] = (uint16_t)min(rg->count[comp], (uint32_t)maxRec);
    s_prep[c] = P.intra_prep[(size_t)comp * ctusW * P.ctus_h + (size_t)row * ctusW + c];
  }
  if (tid == 0) s_tileB[0] = s_tileB[TILE_PAD] = (int16_t)(1 << (bd - 1));   // the "nothing available" constant
  __syncthreads();

  const int c0 = next_intra_ctu(s_count, 0, ctusW, lane);
  if (tid == 0) *(volatile unsigned long long*)myProg = base + (unsigned long long)c0;   // nothing to do before CTU c0
  if (c0 >= ctusW) continue;

  if (warp < IN_CHAIN_WARPS)
  {
    // ============================ chain warps: TU after TU, shared memory only ============================
    int prev = -2;
#ifdef INTRA_PROFILE
    unsigned long long prof[40];
    for (int i = 0; i < 40; i++) prof[i] = 0;
    const long long tJob = clock64();
    if (tid == 0 && row < 128) g_intraRows[(comp * 128 + row) * 4 + 0] = gtime();
    bool firstCtu = true;
#endif
    for (int c = c0; c < ctusW; c = next_intra_ctu(s_count, c + 1, ctusW, lane), n++)
    {
      const int b = n & 1;
      int16_t* tile = s_tileB + b * TILE_PAD;
      const int lineBase = (2 - b) * TILE_PAD + 8;            // the line buffer behind the tiles, as an index of THIS tile (entries -1 .. 4N+1 valid)
      const IntraOp* ops = s_ops + b * maxRec;
      const uint16_t* addrTab = s_addr + b * maxAddr;
      const int16_t* resB = s_resB + b * resSamples;
      const int count = s_count[c];
      const int ox = c * CTW;
#ifdef INTRA_PROFILE
      const long long tf0 = clock64();
#endif
      bar_sync(BAR_FULL + b, IN_CHAIN + 32);                 // staged (fetcher warp): tile, decoded TUs, tables, residuals
      mbar_wait(&s_mbar[b], (n >> 1) & 1);                   // (already complete: makes the bulk-copied bytes visible to this warp)
#ifdef INTRA_PROFILE
      asm volatile("" :: "r"(ops[0].x) : "memory");
      prof[32] += (unsigned long long)(clock64() - tf0); prof[33] += 1;
      if (firstCtu && tid == 0 && row < 128) g_intraRows[(comp * 128 + row) * 4 + 1] = gtime();
      firstCtu = false;
#endif
      if (prev == c - 1)                                     // left neighbours = what the chain produced a moment ago
      {
        // EVERY chain warp writes the column (the same values): a large first TU is predicted by all four warps without a token, and a
        // warp must not read the column before it is there (warp 0 alone doing it was a race that showed with several bitstreams per GPU)
        for (int y = lane; y < ch; y += 32) tile[TIDX(y, -1)] = s_col[y];
        __syncwarp();
      }
      bool prevLarge = false;
      IntraOp opn = ops[0];
      for (int k = 0; k < count; k++)
      {
        const IntraOp op = opn;
        if (k + 1 < count) opn = ops[k + 1];
        const int lg = op_lg(op);                            // compare chain, most frequent first (a jump table costs an indirect branch per TU)
        const bool nextLarge = k + 1 < count && op_lg(opn) >= 4;
        if (lg <= 3)
        {
          if ((k & 3) == warp)
          {
            if (lg == 3) chain_tu<3>(op, k, count, prevLarge, nextLarge, warp, lane, addrTab, resB, tile, lineBase, bd PROF_PASS);
            else         chain_tu<2>(op, k, count, prevLarge, nextLarge, warp, lane, addrTab, resB, tile, lineBase, bd PROF_PASS);
          }
          prevLarge = false;
        }
        else
        {
          if (lg == 4) chain_tu<4>(op, k, count, prevLarge, nextLarge, warp, lane, addrTab, resB, tile, lineBase, bd PROF_PASS);
          else         chain_tu<5>(op, k, count, prevLarge, nextLarge, warp, lane, addrTab, resB, tile, lineBase, bd PROF_PASS);
          prevLarge = true;
        }
      }
      bar_sync(BAR_CHAIN, IN_CHAIN);                         // every TU of the CTU is in the tile
      if (warp == 0)
      {
        const int cwc = min(CTW, W - ox);
        for (int y = lane; y < ch; y += 32) s_col[y] = tile[TIDX(y, cwc - 1)];
        __syncwarp();
        __threadfence_block();
        bar_arrive(BAR_DONE + b, 64);                        // predicted: the writer warp writes it back and publishes
      }
      prev = c;
    }
#ifdef INTRA_PROFILE
    if (tid == 0 && row < 128) g_intraRows[(comp * 128 + row) * 4 + 2] = gtime();
    prof[34] += (unsigned long long)(clock64() - tJob);
    if (lane == 0) for (int i = 0; i < 40; i++) if (prof[i]) atomicAdd(&g_intraProf[i], prof[i]);
#endif
    continue;
  }

  // ============================ fetcher warp: stages CTU n + 1 while the chain predicts CTU n ============================
  if (warp == IN_CHAIN_WARPS)
  {
    int prev = -2;
    for (int c = c0; c < ctusW; n++)
    {
      const int b = n & 1;
      int16_t* tile = s_tileB + b * TILE_PAD;
      IntraOp* ops = s_ops + b * maxRec;
      uint16_t* addrTab = s_addr + b * maxAddr;
      int16_t* resB = s_resB + b * resSamples;
      const int count = s_count[c];
      const uint32_t first = s_first[c];
      const int cn = next_intra_ctu(s_count, c + 1, ctusW, lane);
      const int ox = c * CTW;
      const int cw = min(CTW, W - ox);
      if (n >= 2) bar_sync(BAR_FREE + b, 64);                // the writer has written back the buffer's previous tenant (CTU n - 2)

      // ---- stage CTU c into buffer b ----
      if ((cw & 7) == 0)
      {
        const int vecPerRow = cw >> 3;
        for (int i = lane; i < ch * vecPerRow; i += 32)
        {
          const int y = i / vecPerRow, v = i - y * vecPerRow;
          cp_async16(&tile[TIDX(y, 8 * v)], plane + (size_t)(oy + y) * pitch + ox + 8 * v);
        }
      }
      else
      {
        const int vecPerRow = cw >> 2;                       // widths are multiples of 4
        for (int i = lane; i < ch * vecPerRow; i += 32)
        {
          const int y = i / vecPerRow, v = i - y * vecPerRow;
          cp_async8(&tile[TIDX(y, 4 * v)], plane + (size_t)(oy + y) * pitch + ox + 4 * v);
        }
      }
      if (lane == 0)
      {
        // contiguous spans by TMA bulk copy: decoded TUs, reference-address tables, the CTU's residual span
        const uint4 prep = s_prep[c];
        const unsigned opsBytes = 16u * count;
        const unsigned tabBytes = 16u * (((unsigned)prep.z + 7) >> 3);
        const unsigned resBytes = 16u * (min(prep.y, (unsigned)resSamples) >> 3);
        fence_proxy_async();                                 // earlier generic-proxy reads of this buffer are ordered before the async writes
        mbar_expect_tx(&s_mbar[b], opsBytes + tabBytes + resBytes);
        bulk_copy_g2s(ops, P.intra_ops + first, opsBytes, &s_mbar[b]);
        if (tabBytes) bulk_copy_g2s(addrTab, P.intra_tab + ((size_t)comp * ctusW * P.ctus_h + (size_t)row * ctusW + c) * IN_ADDR, tabBytes, &s_mbar[b]);
        if (resBytes) bulk_copy_g2s(resB, P.resid + prep.x, resBytes, &s_mbar[b]);
      }
      if (ox > 0 && prev != c - 1)                           // left CTU has no intra blocks: its samples have been final since the kernel started
        for (int y = lane; y < ch; y += 32) tile[TIDX(y, -1)] = __ldcg(plane + (size_t)(oy + y) * pitch + ox - 1);
      // the row above (x = -1 .. CTW+31) needs the CTU above-right to be final
      if (row > 0)
      {
        if (upProg)
        {
          if (lane == 0)
          {
            const unsigned long long need = base + (unsigned long long)min(c + 2, ctusW);
            while (*upProg < need) __nanosleep(20);          // back off: the SM's issue slots belong to the chain warps
            __threadfence();
          }
          __syncwarp();
        }
        for (int x = lane - 1; x < CTW + 32; x += 32)
        {
          const int gx = ox + x;
          if (gx >= 0 && gx < W) tile[TIDX(-1, x)] = __ldcg(plane + (size_t)(oy - 1) * pitch + gx);
        }
      }
      cp_async_wait_all();
      mbar_wait(&s_mbar[b], (n >> 1) & 1);
      __threadfence_block();
      bar_arrive(BAR_FULL + b, IN_CHAIN + 32);
      prev = c;
      c = cn;
    }
    continue;
  }

  // ============================ writer warp: writes CTU n back and publishes the row's progress as soon as the chain is done with it ====
  // (its own warp: the write-back of CTU n must not wait behind the staging of CTU n + 1, which waits for the row above — that coupling
  // made every row trail the one above by almost four CTU-times instead of two)
  for (int c = c0; c < ctusW; n++)
  {
    const int b = n & 1;
    const int cn = next_intra_ctu(s_count, c + 1, ctusW, lane);
    bar_sync(BAR_DONE + b, 64);                              // the chain has predicted CTU c
    const int16_t* ptile = s_tileB + b * TILE_PAD;
    const int pox = c * CTW, pcw = min(CTW, W - pox);
    // The row below reads nothing of this CTU but its BOTTOM row (reference samples above / above-right): that row goes out first and the
    // progress is published right behind it — the fence then waits for one 128-byte store instead of the whole tile, which takes about a
    // microsecond out of every row-to-row hand-over (33 of them on the critical path of a 2160p I picture).  Everything else only has to be
    // in memory when the kernel ends (deblocking is stream-ordered behind it).
    {
      const int y = ch - 1;
      int16_t* grow = plane + (size_t)(oy + y) * pitch + pox;
      if ((pcw & 7) == 0) { if (lane < (pcw >> 3)) *((uint4*)grow + lane) = *(const uint4*)&ptile[TIDX(y, 8 * lane)]; }
      else                { if (lane < (pcw >> 2)) *((uint2*)grow + lane) = *(const uint2*)&ptile[TIDX(y, 4 * lane)]; }
    }
    __syncwarp();
    if (lane == 0)
    {
      __threadfence();
      *(volatile unsigned long long*)myProg = base + (unsigned long long)cn;    // every CTU before cn (the next intra CTU, or the end of the row) is final for the row below
    }
    if ((pcw & 7) == 0)
    {
      const int vecPerRow = pcw >> 3;
      for (int i = lane; i < (ch - 1) * vecPerRow; i += 32)
      {
        const int y = i / vecPerRow, v = i - y * vecPerRow;
        *((uint4*)(plane + (size_t)(oy + y) * pitch + pox) + v) = *(const uint4*)&ptile[TIDX(y, 8 * v)];
      }
    }
    else
    {
      const int vecPerRow = pcw >> 2;
      for (int i = lane; i < (ch - 1) * vecPerRow; i += 32)
      {
        const int y = i / vecPerRow, v = i - y * vecPerRow;
        *((uint2*)(plane + (size_t)(oy + y) * pitch + pox) + v) = *(const uint2*)&ptile[TIDX(y, 4 * v)];
      }
    }
    __syncwarp();
    __threadfence_block();
    bar_arrive(BAR_FREE + b, 64);                            // the fetcher may reuse the buffer (CTU n + 2)
    c = cn;
  }
  }   // jobs
}

static int g_coopLimit[16] = {0};          // co-resident intra CTAs per device at worst-case shared memory (intra_max_coresident_blocks)

static int intra_res_samples(const FrameParams& P)
{
  const int ct = 1 << P.hdr.log2_ctu;
  return ct * ct + 2 * ((ct >> P.csx) * (ct >> P.csy));
}
static size_t intra_dyn_smem(int resSamples, int maxRec, int maxAddr, int ctusW)
{
  return (size_t)IN_LINE * 2 + 2 * ((size_t)((IN_TILE + 7) & ~7) * 2 + (size_t)resSamples * 2 + (size_t)maxRec * 16 + (size_t)maxAddr * sizeof(uint16_t)) +
         (size_t)ctusW * (sizeof(uint4) + sizeof(uint32_t) + sizeof(uint16_t));
}

// Largest CTU of a picture, measured on the host records (the same quantities intra_prep_kernel derives per CTU on the device).
IntraSizes intra_sizes_host(const hmr_frame_hdr& h, const hmr_intra* rec, const hmr_ctu_intra_range* range)
{
  IntraSizes z;
  memset(&z, 0, sizeof(z));
  z.maxRec = 1; z.maxAddr = 8; z.resSpan = 8;
  IntraSizes worst;
  memset(&worst, 0, sizeof(worst));
  worst.jobs = -1;
  if (!rec || !range) { z.jobs = -1; return z; }
  const uint32_t ctusW = (h.width + (1u << h.log2_ctu) - 1) >> h.log2_ctu;
  const uint32_t rows = ctusW ? (h.n_ctu + ctusW - 1) / ctusW : 0;
  const bool maskFits = ctusW > 0 && 3 * rows + 1 <= 32 * HMR_INTRA_JOB_WORDS;
  for (uint32_t ctu = 0; ctu < h.n_ctu; ctu++)
    for (int c = 0; c < 3; c++)
    {
      const uint32_t first = range[ctu].first[c], count = range[ctu].count[c];
      if (!count) continue;
      if (maskFits) { const uint32_t jb = (uint32_t)intra_job_index((int)(ctu / ctusW), c, (int)rows); z.jobMask[jb >> 5] |= 1u << (jb & 31); }
      if (first > h.n_intra || count > h.n_intra - first) return worst;   // inconsistent ranges: worst-case capacities
      unsigned tab = 0, mn = 0xffffffffu, mx = 0;
      for (uint32_t k = 0; k < count; k++)
      {
        const hmr_intra& r = rec[first + k];
        if (r.log2_size < 2 || r.log2_size > 5) return worst;
        tab += (4u << r.log2_size) + 1;
        if (r.resid_off != HMR_NO_OFFSET) { mn = r.resid_off < mn ? r.resid_off : mn; const unsigned e = r.resid_off + (1u << (2 * r.log2_size)); mx = e > mx ? e : mx; }
      }
      if ((int)count > z.maxRec) z.maxRec = (int)count;
      if ((int)tab > z.maxAddr) z.maxAddr = (int)tab;
      if (mx > mn && (int)(mx - mn) > z.resSpan) z.resSpan = (int)(mx - mn);
    }
  if (maskFits) for (int w = 0; w < HMR_INTRA_JOB_WORDS; w++) z.jobs += __builtin_popcount(z.jobMask[w]);
  else z.jobs = -1;
  return z;
}

extern "C" int hmr_debug_intra_profile(unsigned long long out[64], int reset)
{
#ifdef INTRA_PROFILE
  if (cudaMemcpyFromSymbol(out, g_intraProf, sizeof(unsigned long long) * 64) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_intraProf, z, sizeof(z)); }
  return 0;
#else
  (void)out; (void)reset;
  return -3;      // built without -DINTRA_PROFILE
#endif
}
extern "C" int hmr_debug_intra_rows(unsigned long long out[3 * 128 * 4])
{
#ifdef INTRA_PROFILE
  return cudaMemcpyFromSymbol(out, g_intraRows, sizeof(unsigned long long) * 3 * 128 * 4) == cudaSuccess ? 0 : -1;
#else
  (void)out;
  return -3;      // built without -DINTRA_PROFILE
#endif
}

size_t intra_table_bytes(int nctu) { return (size_t)3 * nctu * IN_ADDR * sizeof(uint16_t); }

int intra_max_coresident_blocks(int device)
{
  int perSm = 0, sms = 0;
  const size_t worst = intra_dyn_smem(3 * IN_MAXCT * IN_MAXCT, IN_MAXREC, IN_ADDR, IN_MAXCOLS);
  cudaFuncSetAttribute(intra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)worst);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, intra_kernel, IN_THREADS, worst);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  g_coopLimit[device & 15] = perSm * sms;
  return perSm * sms;
}

cudaError_t launch_intra(const FrameParams& P, cudaStream_t s)
{
  if (P.hdr.n_intra == 0) return cudaSuccess;
  if (P.ctus_w > IN_MAXCOLS) return cudaErrorInvalidValue;
  // shared-memory capacities: this picture's largest CTU, rounded so that every buffer stays 16-byte aligned
  const int full = intra_res_samples(P);
  int resSamples = P.intra_res_span > 0 ? min(full, (P.intra_res_span + 7) & ~7) : full;
  int maxRec = P.intra_max_rec > 0 ? min(IN_MAXREC, P.intra_max_rec) : IN_MAXREC;
  int maxAddr = P.intra_max_addr > 0 ? min(IN_ADDR, (P.intra_max_addr + 7) & ~7) : IN_ADDR;
  intra_prep_kernel<<<(3 * P.ctus_w * P.ctus_h + PREP_WARPS - 1) / PREP_WARPS, PREP_WARPS * 32, 0, s>>>(P);
  static const int maxCtas = getenv("HMR_INTRA_CTAS") ? max(1, atoi(getenv("HMR_INTRA_CTAS"))) : IN_MAX_CTAS;   // (tuning knob)
  const int jobs = P.intra_jobs >= 0 ? P.intra_jobs : 3 * P.ctus_h;
  if (jobs == 0) return cudaSuccess;
  // An ordinary launch: a job only waits for a job that was handed out before it, i.e. to a CTA that is running — no CTA ever
  // waits for one that has not started, so the grid need not be co-resident (a cooperative launch would also have to wait until
  // the whole grid fits next to the other bitstreams' kernels).
  const int grid = min(jobs, maxCtas);
  intra_kernel<<<grid, IN_THREADS, intra_dyn_smem(resSamples, maxRec, maxAddr, P.ctus_w), s>>>(P, resSamples, maxRec, maxAddr);
  return cudaGetLastError();
}
