// k_intra.cu — intra prediction + residual add in a CTU-row dependency wavefront.
//
// Replaces TDecCu::xReconIntraQT / xIntraRecQT / xIntraRecBlk (TDecCu.cpp:483-732),
// TComPrediction::initAdiPatternChType + fillReferenceSamples (TComPattern.cpp:107-520, reference-sample
// substitution and [1 2 1] / strong smoothing) and predIntraAng / xPredIntraAng / xPredIntraPlanar /
// predIntraGetPredValDC / xDCPredFiltering (TComPrediction.cpp:182-491, 746-835).
//
// Dependencies: an intra TU reads unfiltered reconstructed samples left / above / above-right / below-left of
// itself.  Inter samples are final before this kernel starts (k_mc + k_resid); intra samples are produced here in
// decode order.  One persistent CTA per (component, CTU row), launched cooperatively so that all CTAs are
// co-resident; row r may process CTU c once row r-1 has published c+2 finished CTUs (the above-right CTU) — the
// WPP dependency.  Progress counters carry an epoch so they never need clearing; a row publishes "all CTUs before
// my next CTU that has intra TUs", so rows/CTUs without intra blocks cost nothing.
//
// The per-TU chain is the critical path (an I picture at 2160p is ~3300 dependent TU steps; a dataflow simulation over
// real streams shows TU-level parallelism inside a CTU is < 1.2x, z-order makes every TU depend on its predecessor),
// so the design minimises the latency of ONE warp walking the TUs:
//   * the CTA is warp-specialised and double-buffered: warps 1-3 ("stagers") do everything that touches global
//     memory for CTU n+1 while warp 0 ("chain") predicts CTU n — the CTU's current samples plus the row above
//     (x = -1 .. CTU+31, after waiting for the row above) and the column to the left, the residuals of its TUs
//     and its decoded TUs / address tables (TMA bulk copies for the contiguous spans, 16-byte cp.async for the tile
//     rows, all in flight together) — then write CTU n back and publish the progress;
//     hand-over through named barriers (bar.arrive / bar.sync), never a full __syncthreads;
//   * a pre-pass kernel (intra_prep_kernel, fully parallel over the picture) turns the records into reference-address
//     tables — for every TU, entry i = shared-memory position of reference sample i AFTER HM's substitution of
//     unavailable samples (pure function of the record, no sample data) — and into decoded 16-byte micro-ops; the
//     whole fillReferenceSamples logic and all mode decoding are off the critical path, the stagers only copy;
//   * warp 0 ("chain") runs size-templated, fully unrolled code per TU: gather the line through the table -> optional
//     smoothing -> prediction (main-reference projection folded into the index) -> + residual -> back into the tile,
//     with warp-level synchronisation only;
//   * a tile is written back once, coalesced.  Cross-CTA reads go through L2 (ld.global.cg); the producer fences before
//     publishing;
//   * shared memory is sized by the largest CTU of the picture (records, address-table entries, residual span: measured
//     on the host records, intra_sizes_host): a resident CTA holds its buffers for the whole wavefront while issuing
//     almost nothing, and with several bitstreams on one GPU that footprint is what the other streams' kernels wait for.
#include "common.cuh"

#define IN_THREADS 128
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

#define IN_NJMAX 5      // (4*32 + 1 + 31) / 32 reference samples per lane at most

// Reference-sample addresses of a TU for this lane (entry lane + 32 j), ahead of the TU's turn.
__device__ __forceinline__ void intra_fetch_addrs(const IntraOp& op, const uint16_t* __restrict__ addrTab, int lane, int a[IN_NJMAX])
{
  const int L = (4 << op_lg(op)) + 1;
  const uint16_t* t = addrTab + op_slot(op);
#pragma unroll
  for (int j = 0; j < IN_NJMAX; j++)
  {
    const int i = lane + 32 * j;
    a[j] = i < L ? (int)t[i] : 0;
  }
}

// One TU on one warp, everything in shared memory / registers.  `a` = this lane's reference addresses (prefetched);
// while the references are being gathered the addresses of the NEXT TU (`opn`) are fetched into `an`.
template <int LG>
__device__ __forceinline__ void intra_tu(const IntraOp op, const int a[IN_NJMAX], const bool hasNext, const IntraOp opn, int an[IN_NJMAX],
                                         const uint16_t* __restrict__ addrTab, int16_t* __restrict__ tile, const int16_t* __restrict__ resB,
                                         int* __restrict__ sref, const int bd, const int lane)
{
  constexpr int N = 1 << LG, N2 = 2 * N, L = 4 * N + 1, NJ = (L + 31) / 32, S = (N * N + 31) / 32;
  const int maxv = (1 << bd) - 1;
  // ---- gather the reference line: [0] bottom-most below-left ... [2N] corner ... [4N] last above-right ----
  int v[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) v[j] = tile[a[j]];             // a[] is 0 (the constant slot) past the end of the line
  if (hasNext) intra_fetch_addrs(opn, addrTab, lane, an);      // independent of the samples: overlaps the gather
  const int flags = op_flags(op);
  if (flags & OPF_FILTER)
  {
    // [1 2 1] / 4 smoothing (or the bilinear "strong" variant) of everything but the two end points, neighbours by shuffle
    bool strong = false;
    int bl = 0, tl = 0, tr = 0;
    if (N == 32 && (flags & OPF_STRONG))
    {
      bl = __shfl_sync(0xffffffffu, v[0], 0); tl = __shfl_sync(0xffffffffu, v[2 % NJ], 0); tr = __shfl_sync(0xffffffffu, v[4 % NJ], 0);
      const int mid0 = __shfl_sync(0xffffffffu, v[1 % NJ], 0), mid1 = __shfl_sync(0xffffffffu, v[3 % NJ], 0);     // line[N], line[3N]
      const int thr = 1 << (bd - 5);
      strong = abs(bl + tl - 2 * mid0) < thr && abs(tl + tr - 2 * mid1) < thr;
    }
    int f[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++)
    {
      const int i = lane + 32 * j;
      int up = __shfl_up_sync(0xffffffffu, v[j], 1), dn = __shfl_down_sync(0xffffffffu, v[j], 1);
      if (j > 0)      { const int w = __shfl_sync(0xffffffffu, v[j - 1], 31); if (lane == 0) up = w; }
      if (j + 1 < NJ) { const int w = __shfl_sync(0xffffffffu, v[j + 1], 0);  if (lane == 31) dn = w; }
      int r;
      if (i == 0 || i >= 4 * N) r = v[j];
      else if (strong) r = i < N2 ? ((N2 - i) * bl + i * tl + N) >> (LG + 1) : (i == N2 ? tl : ((N2 - (i - N2)) * tl + (i - N2) * tr + N) >> (LG + 1));
      else r = (up + 2 * v[j] + dn + 2) >> 2;
      f[j] = r;
    }
#pragma unroll
    for (int j = 0; j < NJ; j++) v[j] = f[j];
  }
#pragma unroll
  for (int j = 0; j < NJ; j++)
  {
    const int i = lane + 32 * j;
    if (i < L) sref[i] = v[j];
  }
  __syncwarp();
  const int* ref = sref;
#define LEFT(y) ref[N2 - 1 - (y)]
#define TOP(x)  ref[N2 + 1 + (x)]
#define EMIT(i, y, x, v) dst[(y) * IN_LD + (x)] = (int16_t)clip3i(0, maxv, (int)(int16_t)(v) + (hasRes ? (int)res[i] : 0))
  const bool hasRes = op.y != HMR_NO_OFFSET;
  const int16_t* res = resB + (hasRes ? op.y : 0u);
  int16_t* dst = tile + op_org(op);
  const int cls = op_cls(op);
  if (cls == OP_PLANAR)
  {
    const int tn = TOP(N), ln = LEFT(N);
#pragma unroll
    for (int j = 0; j < S; j++)
    {
      const int i = lane + 32 * j;
      if (N * N >= 32 || i < N * N)
      {
        const int y = i >> LG, x = i & (N - 1);
        const int p = ((N - 1 - x) * LEFT(y) + (x + 1) * tn + (N - 1 - y) * TOP(x) + (y + 1) * ln + N) >> (LG + 1);
        EMIT(i, y, x, p);
      }
    }
  }
  else if (cls == OP_PCM)
  {
#pragma unroll
    for (int j = 0; j < S; j++)
    {
      const int i = lane + 32 * j;
      if (N * N >= 32 || i < N * N) { const int y = i >> LG, x = i & (N - 1); EMIT(i, y, x, 0); }
    }
  }
  else if (cls == OP_DC)
  {
    const int part = lane < N ? TOP(lane) + LEFT(lane) : 0;
    const int dc = (__reduce_add_sync(0xffffffffu, part) + N) >> (LG + 1);
    const bool edge = flags & OPF_DCEDGE;
#pragma unroll
    for (int j = 0; j < S; j++)
    {
      const int i = lane + 32 * j;
      if (N * N >= 32 || i < N * N)
      {
        const int y = i >> LG, x = i & (N - 1);
        int p = dc;
        if (edge)
        {
          if (x == 0 && y == 0) p = (TOP(0) + LEFT(0) + 2 * dc + 2) >> 2;
          else if (y == 0) p = (TOP(x) + 3 * dc + 2) >> 2;
          else if (x == 0) p = (LEFT(y) + 3 * dc + 2) >> 2;
        }
        EMIT(i, y, x, p);
      }
    }
  }
  else
  {
    const bool ver = flags & OPF_VER;
    const int angle = op_angle(op);
    const int sgn = ver ? 1 : -1;                          // main reference rm[j >= 0] = ref[N2 + sgn*j]
    if (cls == OP_ANG0)
    {
      const bool edge = flags & OPF_EDGE;
      const int corner = ref[N2];
#pragma unroll
      for (int j = 0; j < S; j++)
      {
        const int i = lane + 32 * j;
        if (N * N >= 32 || i < N * N)
        {
          const int y = i >> LG, x = i & (N - 1);
          const int yy = ver ? y : x, xx = ver ? x : y;
          int p = ref[N2 + sgn * (xx + 1)];
          if (edge && xx == 0) p = clip3i(0, maxv, p + ((ref[N2 - sgn * (yy + 1)] - corner) >> 1));
          EMIT(i, y, x, p);
        }
      }
    }
    else if (cls == OP_ANGPOS)
    {
#pragma unroll
      for (int j = 0; j < S; j++)
      {
        const int i = lane + 32 * j;
        if (N * N >= 32 || i < N * N)
        {
          const int y = i >> LG, x = i & (N - 1);
          const int yy = ver ? y : x, xx = ver ? x : y;
          const int pos = (yy + 1) * angle, di = pos >> 5, df = pos & 31;
          const int k = xx + di + 1;
          const int pa = ref[N2 + sgn * k], pb = ref[N2 + sgn * (k + 1)];   // df == 0: pb has weight 0 (the index stays inside the padded line)
          const int p = ((32 - df) * pa + df * pb + 16) >> 5;
          EMIT(i, y, x, p);
        }
      }
    }
    else
    {
      const int inv = op_inv(op);
      // negative angle: rm[k < 0] is the side edge projected onto the main edge: side sample ((128 - k*inv) >> 8) - 1  (TComPrediction.cpp:396-404)
#pragma unroll
      for (int j = 0; j < S; j++)
      {
        const int i = lane + 32 * j;
        if (N * N >= 32 || i < N * N)
        {
          const int y = i >> LG, x = i & (N - 1);
          const int yy = ver ? y : x, xx = ver ? x : y;
          const int pos = (yy + 1) * angle, di = pos >> 5, df = pos & 31;
          const int k0 = xx + di + 1, k1 = k0 + 1;
          const int i0 = k0 >= 0 ? N2 + sgn * k0 : N2 - sgn * ((128 - k0 * inv) >> 8);
          const int i1 = k1 >= 0 ? N2 + sgn * k1 : N2 - sgn * ((128 - k1 * inv) >> 8);
          const int pa = ref[i0], pb = ref[i1];
          const int p = ((32 - df) * pa + df * pb + 16) >> 5;
          EMIT(i, y, x, p);
        }
      }
    }
  }
#undef LEFT
#undef TOP
#undef EMIT
}

// ---- pre-pass: everything about the intra TUs that does not depend on sample data, for the whole picture at once ----
// One WARP per (component, CTU), four per CTA: reference-address tables (compact, TU after TU), decoded micro-ops, and the
// CTU's residual span.  The wavefront kernel below only copies these into shared memory.  (A warp, not a CTA, per CTU:
// most CTUs of an inter picture have no intra TU at all and a warp that finds nothing costs next to nothing.)
#define PREP_WARPS 4
__global__ void __launch_bounds__(PREP_WARPS * 32) intra_prep_kernel(const __grid_constant__ FrameParams P)
{
  __shared__ int s_offAll[PREP_WARPS][IN_MAXREC + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nctu = P.ctus_w * P.ctus_h;
  const int job = blockIdx.x * PREP_WARPS + warp;
  if (job >= 3 * nctu) return;
  const int comp = job / nctu, ctu = job - comp * nctu;
  if (comp > 0 && P.hdr.chroma_format == HMR_CHROMA_400) return;
  const uint32_t first = __ldg(&P.irange[ctu].first[comp]);
  const int count = (int)min(__ldg(&P.irange[ctu].count[comp]), (uint32_t)IN_MAXREC);
  if (count == 0) { if (lane == 0) P.intra_prep[job] = make_uint4(0, 0, 0, 0); return; }
  int* s_off = s_offAll[warp];
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
      const hmr_intra r = P.intra[first + k];
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
  for (int k = 0; k < count; k++)
  {
    const hmr_intra r = P.intra[first + k];
    const int off = s_off[k];
    intra_addr_table(r, tab + off, g, lane);
    if (lane == 0) P.intra_ops[first + k] = intra_make_op(r, g, mn, strongAllowed, off);
  }
  if (lane == 0) P.intra_prep[job] = make_uint4(mn, mx > mn ? mx - mn : 0u, (unsigned)running, 0u);
}

// named barriers (id 0 is __syncthreads)
#define BAR_FULL 1     // +buffer: stagers arrive, chain waits  -> "CTU staged"
#define BAR_DONE 3     // +buffer: chain arrives, stagers wait  -> "CTU predicted"
#define BAR_STAGE 5    // the three stager warps among themselves
#define IN_STAGERS (IN_THREADS - 32)
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(IN_THREADS) intra_kernel(const __grid_constant__ FrameParams P, const int resSamples, const int maxRec, const int maxAddr)
{
  extern __shared__ __align__(16) uint8_t s_dyn[];
  constexpr int TILE_PAD = (IN_TILE + 7) & ~7;
  int16_t* s_tileB = (int16_t*)s_dyn;                                        // [2][TILE_PAD]
  int16_t* s_resB = s_tileB + 2 * TILE_PAD;                                  // [2][resSamples] residuals of a CTU, compact layout relative to minoff
  // capacities = the largest CTU of THIS picture (engine.cu measures the records): a resident CTA holds its shared memory for
  // the whole wavefront, and at saturation that footprint is what other streams' kernels wait for
  IntraOp* s_ops = (IntraOp*)(s_resB + 2 * resSamples);                      // [2][maxRec] decoded TUs of a CTU
  uint16_t* s_addr = (uint16_t*)(s_ops + 2 * maxRec);                        // [2][maxAddr]
  uint4* s_prep = (uint4*)(s_addr + 2 * maxAddr);                            // [ctus_w] per CTU of this row: x = first residual, y = residual span, z = table entries
  uint32_t* s_first = (uint32_t*)(s_prep + P.ctus_w);                        // [ctus_w]
  uint16_t* s_count = (uint16_t*)(s_first + P.ctus_w);                       // [ctus_w]
  __shared__ int s_refBuf[2][4 * 32 + 8]; // reference line of a TU: [0] bottom-most below-left ... [2N] corner ... [4N] last above-right
                                          // (two copies, alternating per TU: the next TU may write while a slow lane still reads)
  __shared__ int16_t s_col[IN_MAXCT];     // right-most column of the CTU the chain just finished (left neighbours of the next one)
  __shared__ __align__(8) unsigned long long s_mbar[2];   // one per buffer: completion of the TMA bulk copies
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int comp = blockIdx.x / P.ctus_h, row = blockIdx.x % P.ctus_h;
  if (comp > 0 && P.hdr.chroma_format == HMR_CHROMA_400) return;
  unsigned long long* myProg = P.intra_progress + comp * P.ctus_h + row;
  const volatile unsigned long long* upProg = row > 0 ? P.intra_progress + comp * P.ctus_h + row - 1 : nullptr;
  const unsigned long long base = P.epoch << 32;

  const int bd = comp ? P.hdr.bit_depth_chroma : P.hdr.bit_depth_luma;
  const int csx = comp ? P.csx : 0, csy = comp ? P.csy : 0;
  IntraGeom g;
  g.CTW = (1 << P.hdr.log2_ctu) >> csx; g.CTH = (1 << P.hdr.log2_ctu) >> csy;
  g.uws = 2 - csx; g.uhs = 2 - csy; g.gw = g.CTW >> 2;
  g.oy = row * g.CTH; g.ox = 0;
  const int CTW = g.CTW, CTH = g.CTH;
  const int W = P.w[comp], H = P.h[comp];
  const int ctusW = P.ctus_w;
  int16_t* plane = P.work.p[comp];
  const int pitch = P.work.pitch[comp];
  const int oy = row * CTH, ch = min(CTH, H - oy);

  for (int c = tid; c < ctusW; c += IN_THREADS)
  {
    const hmr_ctu_intra_range rg = P.irange[row * ctusW + c];
    s_first[c] = rg.first[comp];
    s_count[c] = (uint16_t)min(rg.count[comp], (uint32_t)maxRec);
    s_prep[c] = P.intra_prep[(size_t)comp * ctusW * P.ctus_h + (size_t)row * ctusW + c];
  }
  if (tid == 0)
  {
    s_tileB[0] = s_tileB[TILE_PAD] = (int16_t)(1 << (bd - 1));
    mbar_init(&s_mbar[0], 1); mbar_init(&s_mbar[1], 1);
    fence_proxy_async();                                       // the barriers exist before the async proxy touches them
  }
  __syncthreads();

  const int c0 = next_intra_ctu(s_count, 0, ctusW, lane);
  if (tid == 0) *(volatile unsigned long long*)myProg = base + (unsigned long long)c0;   // nothing to do before CTU c0
  if (c0 >= ctusW) return;

  if (warp == 0)
  {
    // ============================ chain warp: TU after TU, shared memory only ============================
    int prev = -2, n = 0;
    for (int c = c0; c < ctusW; c = next_intra_ctu(s_count, c + 1, ctusW, lane), n++)
    {
      const int b = n & 1;
      int16_t* tile = s_tileB + b * TILE_PAD;
      const IntraOp* ops = s_ops + b * maxRec;
      const uint16_t* addrTab = s_addr + b * maxAddr;
      const int16_t* resB = s_resB + b * resSamples;
      const int count = s_count[c];
      const int ox = c * CTW;
      bar_sync(BAR_FULL + b, IN_THREADS);                    // staged: tile, decoded TUs, tables, residuals
      mbar_wait(&s_mbar[b], (n >> 1) & 1);                   // (already complete: makes the bulk-copied bytes visible to this warp)
      if (prev == c - 1)                                     // left neighbours = what this warp produced a moment ago
        for (int y = lane; y < ch; y += 32) tile[TIDX(y, -1)] = s_col[y];
      IntraOp op = ops[0];
      int a[IN_NJMAX], an[IN_NJMAX];
      intra_fetch_addrs(op, addrTab, lane, a);
      __syncwarp();
      for (int k = 0; k < count; k++)
      {
        const bool hasNext = k + 1 < count;
        const IntraOp opn = ops[hasNext ? k + 1 : k];
        int* sref = &s_refBuf[0][2] + (k & 1) * (4 * 32 + 8);   // indices -1 and 4N+1 are touched (with weight 0) by the 45-degree modes
        const int lg = op_lg(op);                            // compare chain, most frequent first (a jump table costs an indirect branch per TU)
        if (lg == 3)      intra_tu<3>(op, a, hasNext, opn, an, addrTab, tile, resB, sref, bd, lane);
        else if (lg == 2) intra_tu<2>(op, a, hasNext, opn, an, addrTab, tile, resB, sref, bd, lane);
        else if (lg == 4) intra_tu<4>(op, a, hasNext, opn, an, addrTab, tile, resB, sref, bd, lane);
        else              intra_tu<5>(op, a, hasNext, opn, an, addrTab, tile, resB, sref, bd, lane);
        __syncwarp();        // this TU's samples are in the tile before the next TU gathers its reference line
        op = opn;
#pragma unroll
        for (int j = 0; j < IN_NJMAX; j++) a[j] = an[j];
      }
      const int cwc = min(CTW, W - ox);
      for (int y = lane; y < ch; y += 32) s_col[y] = tile[TIDX(y, cwc - 1)];
      prev = c;
      __syncwarp();
      __threadfence_block();
      bar_arrive(BAR_DONE + b, IN_THREADS);                  // predicted: the stagers write it back and publish
    }
    return;
  }

  // ============================ stager warps: everything that touches global memory ============================
  const int st = tid - 32;
  int prev = -2, prevB = 0, n = 0;
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

    // ---- stage CTU c into buffer b (free: its previous tenant was written back in the last iteration) ----
    if ((cw & 7) == 0)
    {
      const int vecPerRow = cw >> 3;
      for (int i = st; i < ch * vecPerRow; i += IN_STAGERS)
      {
        const int y = i / vecPerRow, v = i - y * vecPerRow;
        cp_async16(&tile[TIDX(y, 8 * v)], plane + (size_t)(oy + y) * pitch + ox + 8 * v);
      }
    }
    else
    {
      const int vecPerRow = cw >> 2;                         // widths are multiples of 4
      for (int i = st; i < ch * vecPerRow; i += IN_STAGERS)
      {
        const int y = i / vecPerRow, v = i - y * vecPerRow;
        cp_async8(&tile[TIDX(y, 4 * v)], plane + (size_t)(oy + y) * pitch + ox + 4 * v);
      }
    }
    if (st == 0)
    {
      // contiguous spans by TMA bulk copy: decoded TUs, reference-address tables, the CTU's residual span
      const uint4 prep = s_prep[c];
      const unsigned opsBytes = 16u * count;
      const unsigned tabBytes = 16u * (((unsigned)prep.z + 7) >> 3);
      const unsigned resBytes = 16u * (min(prep.y, (unsigned)resSamples) >> 3);
      fence_proxy_async();                                   // earlier generic-proxy reads of this buffer are ordered before the async writes
      mbar_expect_tx(&s_mbar[b], opsBytes + tabBytes + resBytes);
      bulk_copy_g2s(ops, P.intra_ops + first, opsBytes, &s_mbar[b]);
      if (tabBytes) bulk_copy_g2s(addrTab, P.intra_tab + ((size_t)comp * ctusW * P.ctus_h + (size_t)row * ctusW + c) * IN_ADDR, tabBytes, &s_mbar[b]);
      if (resBytes) bulk_copy_g2s(resB, P.resid + prep.x, resBytes, &s_mbar[b]);
    }
    if (ox > 0 && prev != c - 1)                             // left CTU has no intra blocks: its samples have been final since the kernel started
      for (int y = st; y < ch; y += IN_STAGERS) tile[TIDX(y, -1)] = __ldcg(plane + (size_t)(oy + y) * pitch + ox - 1);
    // the row above (x = -1 .. CTW+31) needs the CTU above-right to be final
    if (upProg)
    {
      if (st == 0)
      {
        const unsigned long long need = base + (unsigned long long)min(c + 2, ctusW);
        while (*upProg < need) { }
        __threadfence();
      }
      bar_sync(BAR_STAGE, IN_STAGERS);
      for (int x = st - 1; x < CTW + 32; x += IN_STAGERS)
      {
        const int gx = ox + x;
        if (gx >= 0 && gx < W) tile[TIDX(-1, x)] = __ldcg(plane + (size_t)(oy - 1) * pitch + gx);
      }
    }
    cp_async_wait_all();
    mbar_wait(&s_mbar[b], (n >> 1) & 1);
    __threadfence_block();
    bar_arrive(BAR_FULL + b, IN_THREADS);

    // ---- write back + publish the CTU the chain is finishing meanwhile ----
    if (prev >= 0)
    {
      bar_sync(BAR_DONE + prevB, IN_THREADS);
      const int16_t* ptile = s_tileB + prevB * TILE_PAD;
      const int pox = prev * CTW, pcw = min(CTW, W - pox);
      if ((pcw & 7) == 0)
      {
        const int vecPerRow = pcw >> 3;
        for (int i = st; i < ch * vecPerRow; i += IN_STAGERS)
        {
          const int y = i / vecPerRow, v = i - y * vecPerRow;
          *((uint4*)(plane + (size_t)(oy + y) * pitch + pox) + v) = *(const uint4*)&ptile[TIDX(y, 8 * v)];
        }
      }
      else
      {
        const int vecPerRow = pcw >> 2;
        for (int i = st; i < ch * vecPerRow; i += IN_STAGERS)
        {
          const int y = i / vecPerRow, v = i - y * vecPerRow;
          *((uint2*)(plane + (size_t)(oy + y) * pitch + pox) + v) = *(const uint2*)&ptile[TIDX(y, 4 * v)];
        }
      }
      bar_sync(BAR_STAGE, IN_STAGERS);
      if (st == 0)
      {
        __threadfence();
        *(volatile unsigned long long*)myProg = base + (unsigned long long)c;    // every CTU before c (the next intra CTU) is final
      }
    }
    prev = c; prevB = b;
    c = cn;
  }
  // ---- the last CTU of the row ----
  {
    bar_sync(BAR_DONE + prevB, IN_THREADS);
    const int16_t* ptile = s_tileB + prevB * TILE_PAD;
    const int pox = prev * CTW, pcw = min(CTW, W - pox);
    const int unit = (pcw & 7) == 0 ? 8 : 4, vecPerRow = pcw / unit;
    for (int i = st; i < ch * vecPerRow; i += IN_STAGERS)
    {
      const int y = i / vecPerRow, v = i - y * vecPerRow;
      if (unit == 8) *((uint4*)(plane + (size_t)(oy + y) * pitch + pox) + v) = *(const uint4*)&ptile[TIDX(y, 8 * v)];
      else           *((uint2*)(plane + (size_t)(oy + y) * pitch + pox) + v) = *(const uint2*)&ptile[TIDX(y, 4 * v)];
    }
    bar_sync(BAR_STAGE, IN_STAGERS);
    if (st == 0)
    {
      __threadfence();
      *(volatile unsigned long long*)myProg = base + (unsigned long long)ctusW;
    }
  }
}

static int intra_res_samples(const FrameParams& P)
{
  const int ct = 1 << P.hdr.log2_ctu;
  return ct * ct + 2 * ((ct >> P.csx) * (ct >> P.csy));
}
static size_t intra_dyn_smem(int resSamples, int maxRec, int maxAddr, int ctusW)
{
  return 2 * ((size_t)((IN_TILE + 7) & ~7) * 2 + (size_t)resSamples * 2 + (size_t)maxRec * 16 + (size_t)maxAddr * sizeof(uint16_t)) +
         (size_t)ctusW * (sizeof(uint4) + sizeof(uint32_t) + sizeof(uint16_t));
}

// Largest CTU of a picture, measured on the host records (the same quantities intra_prep_kernel derives per CTU on the device).
IntraSizes intra_sizes_host(const hmr_frame_hdr& h, const hmr_intra* rec, const hmr_ctu_intra_range* range)
{
  IntraSizes z = { 1, 8, 8 };
  if (!rec || !range) return z;
  for (uint32_t ctu = 0; ctu < h.n_ctu; ctu++)
    for (int c = 0; c < 3; c++)
    {
      const uint32_t first = range[ctu].first[c], count = range[ctu].count[c];
      if (!count) continue;
      if (first > h.n_intra || count > h.n_intra - first) { const IntraSizes worst = { 0, 0, 0 }; return worst; }   // inconsistent ranges: worst-case capacities
      unsigned tab = 0, mn = 0xffffffffu, mx = 0;
      for (uint32_t k = 0; k < count; k++)
      {
        const hmr_intra& r = rec[first + k];
        if (r.log2_size < 2 || r.log2_size > 5) { const IntraSizes worst = { 0, 0, 0 }; return worst; }
        tab += (4u << r.log2_size) + 1;
        if (r.resid_off != HMR_NO_OFFSET) { mn = r.resid_off < mn ? r.resid_off : mn; const unsigned e = r.resid_off + (1u << (2 * r.log2_size)); mx = e > mx ? e : mx; }
      }
      if ((int)count > z.maxRec) z.maxRec = (int)count;
      if ((int)tab > z.maxAddr) z.maxAddr = (int)tab;
      if (mx > mn && (int)(mx - mn) > z.resSpan) z.resSpan = (int)(mx - mn);
    }
  return z;
}

size_t intra_table_bytes(int nctu) { return (size_t)3 * nctu * IN_ADDR * sizeof(uint16_t); }

int intra_max_coresident_blocks(int device)
{
  int perSm = 0, sms = 0;
  const size_t worst = intra_dyn_smem(3 * IN_MAXCT * IN_MAXCT, IN_MAXREC, IN_ADDR, IN_MAXCOLS);
  cudaFuncSetAttribute(intra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)worst);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, intra_kernel, IN_THREADS, worst);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
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
  void* args[] = { (void*)&P, (void*)&resSamples, (void*)&maxRec, (void*)&maxAddr };
  return cudaLaunchCooperativeKernel((const void*)intra_kernel, dim3(3 * P.ctus_h), dim3(IN_THREADS), args, intra_dyn_smem(resSamples, maxRec, maxAddr, P.ctus_w), s);
}
