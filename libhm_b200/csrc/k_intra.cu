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
// The per-TU chain is the critical path (an I picture at 2160p is ~3300 dependent TU steps), so it is latency
// engineered: when a CTU has intra TUs, all 4 warps stage in shared memory (a) the CTU's current samples plus the
// row above (x = -1 .. CTU+31) and the column to the left, (b) the CTU's intra records, (c) the residuals of those
// TUs, with 8/16-byte loads that are all in flight together; then ONE warp runs the TUs back to back out of shared
// memory (reference line -> smoothing -> prediction -> + residual -> back into the tile) with warp-level
// synchronisation only, and the tile is written back once, coalesced, by all warps.  Cross-CTA reads go through L2
// (ld.global.cg); the producer fences before publishing.
#include "common.cuh"

#define IN_THREADS 128
#define IN_MAXCT 64
#define IN_LD (8 + IN_MAXCT + 32)          // tile pitch: 8 columns of left margin (x = -1 lives at column 7), CTU, 32 above-right
#define IN_MAXREC 256                      // intra records of one CTU and component (64x64 in 4x4 blocks)
#define IN_MAXRES (IN_MAXCT * IN_MAXCT * 3) // compact residual span of one CTU, all components (4:4:4 worst case)
#define IN_MAXCOLS 512                     // CTU columns per picture row whose record ranges are cached (8192 / 16)

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

__global__ void __launch_bounds__(IN_THREADS) intra_kernel(const __grid_constant__ FrameParams P)
{
  __shared__ __align__(16) int16_t s_tile[(IN_MAXCT + 1) * IN_LD];   // row 0 = y -1
  __shared__ __align__(16) int16_t s_res[IN_MAXRES];                 // residuals of this CTU, compact layout relative to s_minoff
  __shared__ __align__(16) hmr_intra s_rec[IN_MAXREC];
  __shared__ int s_line[4 * 32 + 1];      // unfiltered reference line: [0] bottom-most below-left ... [2N] corner ... [4N] last above-right
  __shared__ int s_flt[4 * 32 + 1];       // smoothed
  __shared__ int s_rm[3 * 32 + 2];        // angular main reference, index -N..2N stored at +32
  __shared__ uint32_t s_first[IN_MAXCOLS];
  __shared__ uint16_t s_count[IN_MAXCOLS];
  __shared__ unsigned s_minoff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int comp = blockIdx.x / P.ctus_h, row = blockIdx.x % P.ctus_h;
  if (comp > 0 && P.hdr.chroma_format == HMR_CHROMA_400) return;
  unsigned long long* myProg = P.intra_progress + comp * P.ctus_h + row;
  const volatile unsigned long long* upProg = row > 0 ? P.intra_progress + comp * P.ctus_h + row - 1 : nullptr;
  const unsigned long long base = P.epoch << 32;

  const int bd = comp ? P.hdr.bit_depth_chroma : P.hdr.bit_depth_luma;
  const int maxv = (1 << bd) - 1;
  const int csx = comp ? P.csx : 0, csy = comp ? P.csy : 0;
  const int uw = 4 >> csx, uh = 4 >> csy;
  const int CTW = (1 << P.hdr.log2_ctu) >> csx, CTH = (1 << P.hdr.log2_ctu) >> csy;
  const int W = P.w[comp], H = P.h[comp];
  const int ctusW = P.ctus_w;
  int16_t* plane = P.work.p[comp];
  const int pitch = P.work.pitch[comp];
#define T(y, x) s_tile[((y) + 1) * IN_LD + 8 + (x)]

  for (int c = tid; c < ctusW; c += IN_THREADS)
  {
    const hmr_ctu_intra_range rg = P.irange[row * ctusW + c];
    s_first[c] = rg.first[comp];
    s_count[c] = (uint16_t)min(rg.count[comp], (uint32_t)IN_MAXREC);
  }
  if (tid == 0) s_minoff = 0xffffffffu;
  __syncthreads();

  int c = next_intra_ctu(s_count, 0, ctusW, lane);
  if (tid == 0) *(volatile unsigned long long*)myProg = base + (unsigned long long)c;   // nothing to do before CTU c

  while (c < ctusW)
  {
    const uint32_t first = s_first[c];
    const int count = s_count[c];
    if (upProg)
    {
      if (tid == 0)
      {
        const unsigned long long need = base + (unsigned long long)min(c + 2, ctusW);
        while (*upProg < need) { }
        __threadfence();
      }
      __syncthreads();
    }
    const int ox = c * CTW, oy = row * CTH;                 // CTU origin in this component
    const int cw = min(CTW, W - ox), ch = min(CTH, H - oy); // part inside the picture

    // ---- stage 1: tile interior (async), records, row above, column to the left: all loads in flight together ----
    if ((cw & 7) == 0)
    {
      const int vecPerRow = cw >> 3;
      for (int i = tid; i < ch * vecPerRow; i += IN_THREADS)
      {
        const int y = i / vecPerRow, v = i - y * vecPerRow;
        cp_async16(&T(y, 8 * v), plane + (size_t)(oy + y) * pitch + ox + 8 * v);
      }
    }
    else
    {
      const int vecPerRow = cw >> 2;                         // widths are multiples of 4
      for (int i = tid; i < ch * vecPerRow; i += IN_THREADS)
      {
        const int y = i / vecPerRow, v = i - y * vecPerRow;
        cp_async8(&T(y, 4 * v), plane + (size_t)(oy + y) * pitch + ox + 4 * v);
      }
    }
    {
      uint4 rec0 = make_uint4(0, 0, 0, 0), rec1 = rec0;
      int top = 0, left = 0;
      const int gx = ox + tid - 1;
      const bool hasTop = oy > 0 && tid < CTW + 33 && gx >= 0 && gx < W;
      const bool hasLeft = ox > 0 && tid < ch;
      if (tid < count) rec0 = __ldcg((const uint4*)(P.intra + first) + tid);
      if (tid + IN_THREADS < count) rec1 = __ldcg((const uint4*)(P.intra + first) + tid + IN_THREADS);
      if (hasTop) top = __ldcg(plane + (size_t)(oy - 1) * pitch + gx);
      if (hasLeft) left = __ldcg(plane + (size_t)(oy + tid) * pitch + ox - 1);
      if (tid < count) { ((uint4*)s_rec)[tid] = rec0; if (rec0.w != HMR_NO_OFFSET) atomicMin(&s_minoff, rec0.w); }
      if (tid + IN_THREADS < count) { ((uint4*)s_rec)[tid + IN_THREADS] = rec1; if (rec1.w != HMR_NO_OFFSET) atomicMin(&s_minoff, rec1.w); }
      if (hasTop) T(-1, tid - 1) = (int16_t)top;
      if (hasLeft) T(tid, -1) = (int16_t)left;
    }
    __syncthreads();
    // ---- stage 2: residuals of this CTU's TUs, compact layout (one warp per TU, 16-byte async copies) ----
    const unsigned minoff = s_minoff;
    for (int k = warp; k < count; k += IN_THREADS / 32)
    {
      const uint32_t off = s_rec[k].resid_off;
      if (off == HMR_NO_OFFSET) continue;
      const int units = 1 << (2 * s_rec[k].log2_size - 3);   // N*N int16 in 16-byte units
      const uint32_t rel = off - minoff;
      if (rel + 8u * units > (uint32_t)IN_MAXRES) continue;  // cannot happen for a well-formed frame (one CTU's levels are contiguous)
      for (int u = lane; u < units; u += 32) cp_async16(s_res + rel + 8 * u, P.resid + off + 8 * u);
    }
    cp_async_wait_all();
    __syncthreads();
    if (tid == 0) s_minoff = 0xffffffffu;                    // everybody holds `minoff`; next use is two barriers away

    // ---- the dependent chain: one warp, one TU after the other, shared memory only ----
    if (warp == 0)
    {
      for (int k = 0; k < count; k++)
      {
        const hmr_intra r = s_rec[k];
        const int lg = r.log2_size, N = 1 << lg, N2 = 2 * N, L = 4 * N + 1;
        const int x0 = r.x - ox, y0 = r.y - oy;
        // reference samples with substitution (TComPattern.cpp:309-520), unit-granular availability
        const int nl = N / uh, na = N / uw;
        unsigned long long M = 0;                             // bit u = unit u available, units in line order
        M |= (unsigned long long)(__brev((unsigned)r.avail_below_left) >> (32 - nl));
        M |= (unsigned long long)(__brev((unsigned)r.avail_left) >> (32 - nl)) << nl;
        if (r.flags & HMR_INTRA_AVAIL_CORNER) M |= 1ull << (2 * nl);
        M |= (unsigned long long)r.avail_above << (2 * nl + 1);
        M |= (unsigned long long)r.avail_above_right << (2 * nl + 1 + na);
        for (int i = lane; i < L; i += 32)
        {
          int v;
          if (M == 0) v = 1 << (bd - 1);
          else
          {
            const int u = i < N2 ? i / uh : (i == N2 ? 2 * nl : 2 * nl + 1 + (i - N2 - 1) / uw);
            int src = i;
            if (!((M >> u) & 1))
            {
              const unsigned long long lower = M & ((1ull << u) - 1);
              if (lower)
              {
                const int p = 63 - __clzll((long long)lower);                 // nearest available unit before: its LAST sample
                src = p < 2 * nl ? (p + 1) * uh - 1 : (p == 2 * nl ? N2 : N2 + (p - 2 * nl) * uw);
              }
              else
              {
                const int q = __ffsll((long long)M) - 1;                      // first available unit after: its FIRST sample
                src = q < 2 * nl ? q * uh : (q == 2 * nl ? N2 : N2 + 1 + (q - 2 * nl - 1) * uw);
              }
            }
            if (src < N2)       v = T(min(y0 + N2 - 1 - src, CTH - 1), x0 - 1);
            else if (src == N2) v = T(y0 - 1, x0 - 1);
            else                v = T(y0 - 1, min(x0 + (src - N2 - 1), CTW + 31));
          }
          s_line[i] = v;
        }
        __syncwarp();
        const int* ref = s_line;
        if (r.flags & HMR_INTRA_FILTER_REFS)
        {
          const int bl = s_line[0], tl = s_line[N2], tr = s_line[4 * N];
          bool strong = (r.flags & HMR_INTRA_LUMA_RULES) && (P.hdr.flags & HMR_FRM_STRONG_INTRA_SMOOTHING) && N >= 32;
          if (strong)
          {
            const int thr = 1 << (bd - 5);
            strong = abs(bl + tl - 2 * s_line[N]) < thr && abs(tl + tr - 2 * s_line[3 * N]) < thr;
          }
          for (int i = lane; i < L; i += 32)
          {
            int v;
            if (i == 0 || i == 4 * N) v = s_line[i];
            else if (strong)
            {
              const int sh = lg + 1;
              v = i < N2 ? ((N2 - i) * bl + i * tl + N) >> sh : (i == N2 ? tl : ((N2 - (i - N2)) * tl + (i - N2) * tr + N) >> sh);
            }
            else v = (s_line[i - 1] + 2 * s_line[i] + s_line[i + 1] + 2) >> 2;
            s_flt[i] = v;
          }
          ref = s_flt;
          __syncwarp();
        }
#define LEFT(y) ref[N2 - 1 - (y)]
#define TOP(x)  ref[N2 + 1 + (x)]
        const int mode = r.mode;
        const bool lumaRules = r.flags & HMR_INTRA_LUMA_RULES;
        const bool hasRes = r.resid_off != HMR_NO_OFFSET;
        const int16_t* res = s_res + (hasRes ? r.resid_off - minoff : 0u);
        int dc = 0, angle = 0;
        bool ver = true;
        if (mode == 1)
        {
          const int part = lane < N ? TOP(lane) + LEFT(lane) : 0;
          dc = (__reduce_add_sync(0xffffffffu, part) + N) >> (lg + 1);
        }
        else if (mode >= 2)
        {
          ver = mode >= 18;
          const int am = ver ? mode - 26 : -(mode - 10);
          const int aa = abs(am);
          angle = am < 0 ? -c_angTab[aa] : c_angTab[aa];
          const int inv = c_invTab[aa];
          const int last = (N * angle) >> 5;
          // main reference rm[-N..2N] (stored at +32): rm[0] = corner, rm[i>0] = main edge, rm[i<0] = projected side edge
          for (int i = lane - 32; i <= N2; i += 32)
          {
            if (i >= 0) { if (angle < 0 && i > N) continue; s_rm[32 + i] = ver ? TOP(i - 1) : LEFT(i - 1); }
            else if (angle < 0 && i > last)
            {
              const int sidx = ((128 + (-i) * inv) >> 8) - 1;
              s_rm[32 + i] = ver ? LEFT(sidx) : TOP(sidx);
            }
          }
          __syncwarp();
        }
        const bool edge = lumaRules && N <= 16 && !(r.flags & HMR_INTRA_NO_EDGE_FLT);
        for (int i = lane; i < N * N; i += 32)
        {
          const int y = i >> lg, x = i & (N - 1);
          int v;
          if (mode == 0)
            v = ((N - 1 - x) * LEFT(y) + (x + 1) * TOP(N) + (N - 1 - y) * TOP(x) + (y + 1) * LEFT(N) + N) >> (lg + 1);
          else if (mode == 1)
          {
            v = dc;
            if (lumaRules && N <= 16)
            {
              if (x == 0 && y == 0) v = (TOP(0) + LEFT(0) + 2 * dc + 2) >> 2;
              else if (y == 0) v = (TOP(x) + 3 * dc + 2) >> 2;
              else if (x == 0) v = (LEFT(y) + 3 * dc + 2) >> 2;
            }
          }
          else
          {
            const int yy = ver ? y : x, xx = ver ? x : y;       // coordinates in the (possibly transposed) prediction frame
            const int pos = (yy + 1) * angle, di = pos >> 5, df = pos & 31;
            const int* rm = s_rm + 32;
            if (angle == 0)
            {
              v = rm[xx + 1];
              if (edge && xx == 0) v = clip3i(0, maxv, v + (((ver ? LEFT(yy) : TOP(yy)) - ref[N2]) >> 1));
            }
            else if (df) v = ((32 - df) * rm[xx + di + 1] + df * rm[xx + di + 2] + 16) >> 5;
            else v = rm[xx + di + 1];
          }
          v = (int16_t)v;
          const int rr = hasRes ? res[i] : 0;
          T(y0 + y, x0 + x) = (int16_t)clip3i(0, maxv, v + rr);
        }
#undef LEFT
#undef TOP
        __syncwarp();        // this TU's samples are in the tile before the next TU builds its reference line
      }
    }
    __syncthreads();
    // ---- write the CTU back (inter samples are rewritten with the values they had) ----
    if ((cw & 7) == 0)
    {
      const int vecPerRow = cw >> 3;
      for (int i = tid; i < ch * vecPerRow; i += IN_THREADS)
      {
        const int y = i / vecPerRow, v = i - y * vecPerRow;
        *((uint4*)(plane + (size_t)(oy + y) * pitch + ox) + v) = *(const uint4*)&T(y, 8 * v);
      }
    }
    else
    {
      const int vecPerRow = cw >> 2;
      for (int i = tid; i < ch * vecPerRow; i += IN_THREADS)
      {
        const int y = i / vecPerRow, v = i - y * vecPerRow;
        *((uint2*)(plane + (size_t)(oy + y) * pitch + ox) + v) = *(const uint2*)&T(y, 4 * v);
      }
    }
    c = next_intra_ctu(s_count, c + 1, ctusW, lane);
    __syncthreads();
    if (tid == 0)
    {
      __threadfence();
      *(volatile unsigned long long*)myProg = base + (unsigned long long)c;      // every CTU before the next intra CTU is final
    }
  }
#undef T
}

int intra_max_coresident_blocks(int device)
{
  int perSm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, intra_kernel, IN_THREADS, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return perSm * sms;
}

cudaError_t launch_intra(const FrameParams& P, cudaStream_t s)
{
  if (P.hdr.n_intra == 0) return cudaSuccess;
  if (P.ctus_w > IN_MAXCOLS) return cudaErrorInvalidValue;
  void* args[] = { (void*)&P };
  return cudaLaunchCooperativeKernel((const void*)intra_kernel, dim3(3 * P.ctus_h), dim3(IN_THREADS), args, 0, s);
}
