// k_intra.cu — intra prediction + residual add in a CTU-row dependency wavefront.
//
// Replaces TDecCu::xReconIntraQT / xIntraRecQT / xIntraRecBlk (TDecCu.cpp:483-732),
// TComPrediction::initAdiPatternChType + fillReferenceSamples (TComPattern.cpp:107-520, reference-sample
// substitution and [1 2 1] / strong smoothing) and predIntraAng / xPredIntraAng / xPredIntraPlanar /
// predIntraGetPredValDC / xDCPredFiltering (TComPrediction.cpp:182-491, 746-835).
//
// Dependencies: an intra TU reads unfiltered reconstructed samples left / above / above-right / below-left of
// itself.  Inter samples are final before this kernel starts (k_mc + k_resid); intra samples are produced
// here in decode order.  One persistent CTA per (component, CTU row), launched cooperatively so that all
// CTAs are co-resident; row r may process CTU c once row r-1 has published c+2 finished CTUs (the
// above-right CTU), exactly the WPP dependency.  Progress counters carry an epoch so they never need clearing.
// Cross-CTA sample reads go through L2 (ld.global.cg); the producer fences before publishing.
#include "common.cuh"
#include <cooperative_groups.h>

#define IN_THREADS 128

__device__ __forceinline__ int ldcg16(const int16_t* p) { return __ldcg(p); }

__global__ void __launch_bounds__(IN_THREADS) intra_kernel(const __grid_constant__ FrameParams P)
{
  __shared__ int s_line[4 * 32 + 1];      // unfiltered reference line: [0] bottom-most below-left ... [2N] corner ... [4N] last above-right
  __shared__ int s_flt[4 * 32 + 1];       // filtered
  __shared__ int s_rm[3 * 32 + 2];        // angular main reference, index -N..2N stored at +32
  const int tid = threadIdx.x;
  const int comp = blockIdx.x / P.ctus_h, row = blockIdx.x % P.ctus_h;
  if (comp > 0 && P.hdr.chroma_format == HMR_CHROMA_400) return;
  unsigned long long* myProg = P.intra_progress + comp * P.ctus_h + row;
  const volatile unsigned long long* upProg = row > 0 ? P.intra_progress + comp * P.ctus_h + row - 1 : nullptr;
  const unsigned long long base = P.epoch << 32;

  const int bd = comp ? P.hdr.bit_depth_chroma : P.hdr.bit_depth_luma;
  const int maxv = (1 << bd) - 1;
  const int uw = 4 >> (comp ? P.csx : 0), uh = 4 >> (comp ? P.csy : 0);
  int16_t* plane = P.work.p[comp];
  const int pitch = P.work.pitch[comp];

  for (int c = 0; c < P.ctus_w; c++)
  {
    const hmr_ctu_intra_range rg = P.irange[row * P.ctus_w + c];
    const uint32_t first = rg.first[comp], count = rg.count[comp];
    if (count && upProg)
    {
      if (tid == 0)
      {
        const unsigned long long need = base + (unsigned long long)min(c + 2, P.ctus_w);
        while (*upProg < need) { }
        __threadfence();
      }
      __syncthreads();
    }
    for (uint32_t k = 0; k < count; k++)
    {
      const hmr_intra r = P.intra[first + k];
      const int N = 1 << r.log2_size, N2 = 2 * N, L = 4 * N + 1;
      const int x0 = r.x, y0 = r.y;
      // ---- reference samples with substitution (TComPattern.cpp:309-520), unit-granular availability ----
      const int nl = N / uh, na = N / uw;                   // units on the left / above edge
      unsigned long long M = 0;                             // bit u = unit u available, units in line order
      M |= (unsigned long long)(__brev((unsigned)r.avail_below_left) >> (32 - nl));
      M |= (unsigned long long)(__brev((unsigned)r.avail_left) >> (32 - nl)) << nl;
      if (r.flags & HMR_INTRA_AVAIL_CORNER) M |= 1ull << (2 * nl);
      M |= (unsigned long long)r.avail_above << (2 * nl + 1);
      M |= (unsigned long long)r.avail_above_right << (2 * nl + 1 + na);
      for (int i = tid; i < L; i += IN_THREADS)
      {
        int v;
        if (M == 0) v = 1 << (bd - 1);
        else
        {
          int u = i < N2 ? i / uh : (i == N2 ? 2 * nl : 2 * nl + 1 + (i - N2 - 1) / uw);
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
          const int16_t* a = src < N2 ? plane + (size_t)(y0 + N2 - 1 - src) * pitch + x0 - 1
                           : (src == N2 ? plane + (size_t)(y0 - 1) * pitch + x0 - 1
                                        : plane + (size_t)(y0 - 1) * pitch + x0 + (src - N2 - 1));
          v = ldcg16(a);
        }
        s_line[i] = v;
      }
      __syncthreads();
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
        for (int i = tid; i < L; i += IN_THREADS)
        {
          int v;
          if (i == 0 || i == 4 * N) v = s_line[i];
          else if (strong)
          {
            const int sh = r.log2_size + 1;
            v = i < N2 ? ((N2 - i) * bl + i * tl + N) >> sh : (i == N2 ? tl : ((N2 - (i - N2)) * tl + (i - N2) * tr + N) >> sh);
          }
          else v = (s_line[i - 1] + 2 * s_line[i] + s_line[i + 1] + 2) >> 2;
          s_flt[i] = v;
        }
        ref = s_flt;
        __syncthreads();
      }
#define LEFT(y) ref[N2 - 1 - (y)]
#define TOP(x)  ref[N2 + 1 + (x)]
      const int mode = r.mode;
      const bool lumaRules = r.flags & HMR_INTRA_LUMA_RULES;
      const int16_t* rs = r.resid_off != HMR_NO_OFFSET ? P.resid + r.resid_off : nullptr;
      int dc = 0, angle = 0;
      bool ver = true;
      if (mode == 1)
      {
        int sum = 0;
        for (int i = 0; i < N; i++) sum += TOP(i) + LEFT(i);     // every thread redundantly (N <= 32, broadcast reads)
        dc = (sum + N) / N2;
      }
      else if (mode >= 2)
      {
        const int angTab[9] = { 0, 2, 5, 9, 13, 17, 21, 26, 32 };
        const int invTab[9] = { 0, 4096, 1638, 910, 630, 482, 390, 315, 256 };
        ver = mode >= 18;
        const int am = ver ? mode - 26 : -(mode - 10);
        const int aa = abs(am);
        angle = am < 0 ? -angTab[aa] : angTab[aa];
        const int inv = invTab[aa];
        const int last = (N * angle) >> 5;
        // main reference rm[-N..2N] (stored at +32): rm[0] = corner, rm[i>0] = main edge, rm[i<0] = projected side edge
        for (int i = tid - 32; i <= N2; i += IN_THREADS)
        {
          if (i >= 0) { if (angle < 0 && i > N) continue; s_rm[32 + i] = ver ? TOP(i - 1) : LEFT(i - 1); }
          else if (angle < 0 && i > last)
          {
            const int sidx = ((128 + (-i) * inv) >> 8) - 1;
            s_rm[32 + i] = ver ? LEFT(sidx) : TOP(sidx);
          }
        }
        __syncthreads();
      }
      const bool edge = lumaRules && N <= 16 && !(r.flags & HMR_INTRA_NO_EDGE_FLT);
      for (int i = tid; i < N * N; i += IN_THREADS)
      {
        const int y = i >> r.log2_size, x = i & (N - 1);
        int v;
        if (mode == 0)
          v = ((N - 1 - x) * LEFT(y) + (x + 1) * TOP(N) + (N - 1 - y) * TOP(x) + (y + 1) * LEFT(N) + N) >> (r.log2_size + 1);
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
        const int rr = rs ? rs[i] : 0;
        plane[(size_t)(y0 + y) * pitch + x0 + x] = (int16_t)clip3i(0, maxv, v + rr);
      }
#undef LEFT
#undef TOP
      __syncthreads();     // this TU's samples are visible to the CTA before the next TU reads them
    }
    if (count || c == P.ctus_w - 1 || ((c & 3) == 3))
    {
      // publish progress (always for CTUs that wrote samples; every 4th and the last CTU otherwise)
      __syncthreads();
      if (tid == 0)
      {
        __threadfence();
        *(volatile unsigned long long*)myProg = base + (unsigned long long)(c + 1);
      }
    }
  }
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
  void* args[] = { (void*)&P };
  return cudaLaunchCooperativeKernel((const void*)intra_kernel, dim3(3 * P.ctus_h), dim3(IN_THREADS), args, 0, s);
}
