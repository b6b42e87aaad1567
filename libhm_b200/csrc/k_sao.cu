// k_sao.cu — sample adaptive offset, reading the deblocked working picture and writing the DPB slot.
//
// Replaces TComSampleAdaptiveOffset::SAOProcess / offsetCTU / offsetBlock (TComSampleAdaptiveOffset.cpp:375-734).
// HM first copies the whole deblocked picture (so that a CTU never sees SAO'd neighbours) and then rewrites the CTUs
// that have SAO on; here source and destination are different surfaces, so the copy IS the kernel: every sample is
// read once from `work` and written once to `out`, with or without an offset.  Merge candidates are already
// resolved on the host (reconstructBlkSAOParams, :348-372).  The per-CTU availability bits reproduce HM's
// skipping of the first/last row/column and of the four corners for the diagonal classes (:501-503,549-551,585-587,626-628).
#include "common.cuh"

__device__ __forceinline__ int sgn3(int v) { return (v > 0) - (v < 0); }

__global__ void __launch_bounds__(256) sao_kernel(const __grid_constant__ FrameParams P)
{
  const int comp = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= P.w[comp] || y >= P.h[comp]) return;
  const int cx = comp ? P.csx : 0, cy = comp ? P.csy : 0;
  const int st = P.work.pitch[comp];
  const int16_t* __restrict__ p = P.work.p[comp] + (size_t)y * st + x;
  const int v = *p;
  int out = v;
  if (P.hdr.flags & HMR_FRM_SAO)
  {
    const int lc = P.hdr.log2_ctu;
    const int cxs = lc - cx, cys = lc - cy;                   // CTU size in this component (log2)
    const int ctuX = x >> cxs, ctuY = y >> cys;
    const hmr_ctu* cp = P.ctu + (size_t)ctuY * P.ctus_w + ctuX;
    const hmr_sao s = cp->sao[comp];
    if (s.type != HMR_SAO_OFF)
    {
      const int bd = comp ? P.hdr.bit_depth_chroma : P.hdr.bit_depth_luma, maxv = (1 << bd) - 1;
      if (s.type == HMR_SAO_BO)
      {
        const int k = ((v >> (bd - 5)) - s.band) & 31;
        if (k < 4) out = clip3i(0, maxv, v + s.off[k]);
      }
      else
      {
        const int bx = ctuX << cxs, by = ctuY << cys;
        const int bw = min(1 << cxs, P.w[comp] - bx), bh = min(1 << cys, P.h[comp] - by);
        const bool firstRow = y == by, lastRow = y == by + bh - 1, firstCol = x == bx, lastCol = x == bx + bw - 1;
        const int av = cp->avail;
        const bool L = av & HMR_AV_L, R = av & HMR_AV_R, A = av & HMR_AV_A, B = av & HMR_AV_B;
        const bool colOk = !(firstCol && !L) && !(lastCol && !R);
        bool ok; int dx, dy;
        switch (s.type)
        {
          case HMR_SAO_EO_0:  dx = 1; dy = 0; ok = colOk; break;
          case HMR_SAO_EO_90: dx = 0; dy = 1; ok = !(firstRow && !A) && !(lastRow && !B); break;
          case HMR_SAO_EO_135:
            dx = 1; dy = 1;
            if (firstRow && bh > 1) ok = firstCol ? (av & HMR_AV_AL) != 0 : (A && !(lastCol && !R));
            else if (lastRow)       ok = lastCol ? (av & HMR_AV_BR) != 0 : (B && !(firstCol && !L));
            else                    ok = colOk;
            break;
          default:
            dx = -1; dy = 1;
            if (firstRow && bh > 1) ok = lastCol ? (av & HMR_AV_AR) != 0 : (A && !(firstCol && !L));
            else if (lastRow)       ok = firstCol ? (av & HMR_AV_BL) != 0 : (B && !(lastCol && !R));
            else                    ok = colOk;
            break;
        }
        if (ok)
        {
          const int a0 = p[-dy * st - dx], b0 = p[dy * st + dx];
          const int e = sgn3(v - a0) + sgn3(v - b0);
          if (e) out = clip3i(0, maxv, v + s.off[e < 0 ? e + 2 : e + 1]);
        }
      }
    }
  }
  P.out.p[comp][(size_t)y * P.out.pitch[comp] + x] = (int16_t)out;
}

void launch_sao(const FrameParams& P, cudaStream_t s)
{
  dim3 block(64, 4);
  dim3 grid((P.w[0] + 63) / 64, (P.h[0] + 3) / 4, P.hdr.chroma_format == HMR_CHROMA_400 ? 1 : 3);
  sao_kernel<<<grid, block, 0, s>>>(P);
}
