// k_deblock.cu — in-loop deblocking filter: all vertical edges of the picture (DIR 0), then all horizontal
// edges (DIR 1) on the result, exactly the two-pass order of TComLoopFilter::loopFilterPic (TComLoopFilter.cpp:130-155).
//
// Replaces xEdgeFilterLuma / xEdgeFilterChroma / xPelFilterLuma / xPelFilterChroma / xUseStrongFiltering /
// xCalcDP / xCalcDQ and the tc/beta tables (TComLoopFilter.cpp:59-67, 540-922).  Boundary strengths arrive
// precomputed from the host (north_star), one byte per 4x4 luma unit.
//
// One thread per 4-sample edge segment on the 8x8 luma grid.  Within a pass every segment touches a disjoint set of
// samples (edges are 8 apart, a filter reads 4 and writes <= 3 samples per side), so the pass is embarrassingly
// parallel and works in place.  The thread that owns a luma segment also filters the co-located chroma lines when
// the edge lies on the 8-sample chroma grid.
#include "common.cuh"

__constant__ uint8_t c_tc[54] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24 };
__constant__ uint8_t c_beta[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64 };
__constant__ uint8_t c_chromaQp420[58] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51 };

template <int DIR>
__global__ void __launch_bounds__(256) deblock_kernel(const __grid_constant__ FrameParams P)
{
  // DIR 0: thread = (edge column ex -> x = 8 ex, unit row uy);   DIR 1: thread = (unit column ux, edge row ey -> y = 8 ey)
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y * blockDim.y + threadIdx.y;
  int ux, uy;
  if (DIR == 0) { ux = 2 * a; uy = b; } else { ux = a; uy = 2 * b; }
  if (ux >= P.w4 || uy >= P.h4) return;
  const int bs = (P.bs[(size_t)uy * P.w4 + ux] >> (DIR ? 2 : 0)) & 3;
  if (!bs) return;
  const int x = ux * 4, y = uy * 4;
  const int px = DIR ? x : x - 1, py = DIR ? y - 1 : y;
  const int qpQ = P.qp[(size_t)(y >> 3) * P.w8 + (x >> 3)], qpP = P.qp[(size_t)(py >> 3) * P.w8 + (px >> 3)];
  bool nfP = false, nfQ = false;
  if (P.cu_flags)
  {
    nfQ = P.cu_flags[(size_t)(y >> 3) * P.w8 + (x >> 3)] & HMR_CU_NOFILTER;
    nfP = P.cu_flags[(size_t)(py >> 3) * P.w8 + (px >> 3)] & HMR_CU_NOFILTER;
  }
  const hmr_ctu* cq = P.ctu + (size_t)(y >> P.hdr.log2_ctu) * P.ctus_w + (x >> P.hdr.log2_ctu);
  const int betaOff = cq->beta_offset_div2, tcOff = cq->tc_offset_div2;
  const int qp = (qpP + qpQ + 1) >> 1;

  // ------------------------------------------------ luma ------------------------------------------------
  {
    const int bd = P.hdr.bit_depth_luma, scale = 1 << (bd - 8), maxv = (1 << bd) - 1;
    const int tc = c_tc[clip3i(0, 53, qp + 2 * (bs - 1) + (tcOff << 1))] * scale;
    const int beta = c_beta[clip3i(0, 51, qp + (betaOff << 1))] * scale;
    const int pitch = P.work.pitch[0];
    int16_t* p = P.work.p[0] + (size_t)y * pitch + x;
    const int step = DIR ? pitch : 1, line = DIR ? 1 : pitch;
    int m[4][8];                                       // m[i][k+4]: line i, position k = -4..3 across the edge
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int k = 0; k < 8; k++) m[i][k] = p[(k - 4) * step + i * line];
    const int dp0 = abs(m[0][1] - 2 * m[0][2] + m[0][3]), dq0 = abs(m[0][4] - 2 * m[0][5] + m[0][6]);
    const int dp3 = abs(m[3][1] - 2 * m[3][2] + m[3][3]), dq3 = abs(m[3][4] - 2 * m[3][5] + m[3][6]);
    const int d0 = dp0 + dq0, d3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = d0 + d3;
    if (d < beta)
    {
      const int sideThr = (beta + (beta >> 1)) >> 3, thrCut = tc * 10;
      const bool fp = dp < sideThr, fq = dq < sideThr;
      const bool sw0 = (abs(m[0][0] - m[0][3]) + abs(m[0][7] - m[0][4]) < (beta >> 3)) && (2 * d0 < (beta >> 2)) && (abs(m[0][3] - m[0][4]) < ((tc * 5 + 1) >> 1));
      const bool sw3 = (abs(m[3][0] - m[3][3]) + abs(m[3][7] - m[3][4]) < (beta >> 3)) && (2 * d3 < (beta >> 2)) && (abs(m[3][3] - m[3][4]) < ((tc * 5 + 1) >> 1));
      const bool strong = sw0 && sw3;
#pragma unroll
      for (int i = 0; i < 4; i++)
      {
        const int m0 = m[i][0], m1 = m[i][1], m2 = m[i][2], m3 = m[i][3], m4 = m[i][4], m5 = m[i][5], m6 = m[i][6], m7 = m[i][7];
        int n1 = m1, n2 = m2, n3 = m3, n4 = m4, n5 = m5, n6 = m6;
        if (strong)
        {
          n3 = clip3i(m3 - 2 * tc, m3 + 2 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
          n4 = clip3i(m4 - 2 * tc, m4 + 2 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
          n2 = clip3i(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
          n5 = clip3i(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
          n1 = clip3i(m1 - 2 * tc, m1 + 2 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
          n6 = clip3i(m6 - 2 * tc, m6 + 2 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
        }
        else
        {
          int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
          if (abs(delta) < thrCut)
          {
            delta = clip3i(-tc, tc, delta);
            n3 = clip3i(0, maxv, m3 + delta);
            n4 = clip3i(0, maxv, m4 - delta);
            const int tc2 = tc >> 1;
            if (fp) n2 = clip3i(0, maxv, m2 + clip3i(-tc2, tc2, ((((m1 + m3 + 1) >> 1) - m2 + delta) >> 1)));
            if (fq) n5 = clip3i(0, maxv, m5 + clip3i(-tc2, tc2, ((((m6 + m4 + 1) >> 1) - m5 - delta) >> 1)));
          }
        }
        int16_t* q = p + i * line;
        if (!nfP) { if (n3 != m3) q[-step] = (int16_t)n3; if (n2 != m2) q[-2 * step] = (int16_t)n2; if (n1 != m1) q[-3 * step] = (int16_t)n1; }
        if (!nfQ) { if (n4 != m4) q[0] = (int16_t)n4; if (n5 != m5) q[step] = (int16_t)n5; if (n6 != m6) q[2 * step] = (int16_t)n6; }
      }
    }
  }

  // ------------------------------------------------ chroma: BS 2 only, 8-sample chroma grid ------------------------------------------------
  if (bs > 1 && P.hdr.chroma_format != HMR_CHROMA_400)
  {
    const int grid = DIR ? (8 << P.csy) : (8 << P.csx);
    if (((DIR ? y : x) % grid) != 0) return;
    const int nlines = DIR ? (4 >> P.csx) : (4 >> P.csy);
    const int bd = P.hdr.bit_depth_chroma, maxv = (1 << bd) - 1;
    for (int c = 1; c < 3; c++)
    {
      int q = qp + (c == 1 ? P.hdr.pps_cb_qp_offset : P.hdr.pps_cr_qp_offset);
      if (q >= 58) { if (P.hdr.chroma_format == HMR_CHROMA_420) q -= 6; else if (q > 51) q = 51; }
      else if (q >= 0) q = P.hdr.chroma_format == HMR_CHROMA_420 ? c_chromaQp420[q] : min(q, 51);
      const int tc = c_tc[clip3i(0, 53, q + 2 * (bs - 1) + (tcOff << 1))] * (1 << (bd - 8));
      const int pitch = P.work.pitch[c];
      int16_t* p = P.work.p[c] + (size_t)(y >> P.csy) * pitch + (x >> P.csx);
      const int step = DIR ? pitch : 1, line = DIR ? 1 : pitch;
      for (int i = 0; i < nlines; i++)
      {
        int16_t* s = p + i * line;
        const int m2 = s[-2 * step], m3 = s[-step], m4 = s[0], m5 = s[step];
        const int delta = clip3i(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
        if (!nfP) s[-step] = (int16_t)clip3i(0, maxv, m3 + delta);
        if (!nfQ) s[0] = (int16_t)clip3i(0, maxv, m4 - delta);
      }
    }
  }
}

void launch_deblock(const FrameParams& P, int dir, cudaStream_t s)
{
  if (!(P.hdr.flags & HMR_FRM_DEBLOCK) || !P.bs) return;
  dim3 block(32, 8);
  if (dir == 0)
  {
    dim3 grid(((P.w4 + 1) / 2 + 31) / 32, (P.h4 + 7) / 8);
    deblock_kernel<0><<<grid, block, 0, s>>>(P);
  }
  else
  {
    dim3 grid((P.w4 + 31) / 32, ((P.h4 + 1) / 2 + 7) / 8);
    deblock_kernel<1><<<grid, block, 0, s>>>(P);
  }
}
