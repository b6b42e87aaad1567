// k_sao.cu — sample adaptive offset, reading the deblocked working picture and writing the DPB slot.
//
// Replaces TComSampleAdaptiveOffset::SAOProcess / offsetCTU / offsetBlock (TComSampleAdaptiveOffset.cpp:375-734).
// HM first copies the whole deblocked picture (so that a CTU never sees SAO'd neighbours) and then rewrites the CTUs
// that have SAO on; here source and destination are different surfaces, so the copy IS the kernel: every sample is
// read once from `work` and written once to `out`, with or without an offset.  Merge candidates are already
// resolved on the host (reconstructBlkSAOParams, :348-372).  The per-CTU availability bits reproduce HM's
// skipping of the first/last row/column and of the four corners for the diagonal classes (:501-503,549-551,585-587,626-628).
//
// One thread owns 8 consecutive samples of a row (one 16-byte load of the row, plus the rows above/below and the
// two flanking samples when the CTU's edge class needs them; one 16-byte store).  8 samples never straddle a CTU.
#include "common.cuh"

__device__ __forceinline__ int sgn3(int v) { return (v > 0) - (v < 0); }
__device__ __forceinline__ void unpack8(const uint4 q, int v[8])
{
  v[0] = (int16_t)(q.x & 0xffff); v[1] = (int)q.x >> 16; v[2] = (int16_t)(q.y & 0xffff); v[3] = (int)q.y >> 16;
  v[4] = (int16_t)(q.z & 0xffff); v[5] = (int)q.z >> 16; v[6] = (int16_t)(q.w & 0xffff); v[7] = (int)q.w >> 16;
}

__global__ void __launch_bounds__(256) sao_kernel(const __grid_constant__ FrameParams P)
{
  const int comp = blockIdx.z;
  const int x = 8 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y * blockDim.y + threadIdx.y;
  const int W = P.w[comp], H = P.h[comp];
  if (x >= W || y >= H) return;
  const int st = P.work.pitch[comp];
  const int16_t* __restrict__ p = P.work.p[comp] + (size_t)y * st + x;
  uint4 q = *(const uint4*)p;                                    // rows are padded to 64 samples: reading past W is safe
  if (P.hdr.flags & HMR_FRM_SAO)
  {
    const int cx = comp ? P.csx : 0, cy = comp ? P.csy : 0;
    const int cxs = P.hdr.log2_ctu - cx, cys = P.hdr.log2_ctu - cy;   // CTU size in this component (log2)
    const int ctuX = x >> cxs, ctuY = y >> cys;
    const hmr_ctu* cp = P.ctu + (size_t)ctuY * P.ctus_w + ctuX;
    const int type = cp->sao[comp].type;
    if (type != HMR_SAO_OFF)
    {
      const int bd = comp ? P.hdr.bit_depth_chroma : P.hdr.bit_depth_luma, maxv = (1 << bd) - 1;
      const int o0 = cp->sao[comp].off[0], o1 = cp->sao[comp].off[1], o2 = cp->sao[comp].off[2], o3 = cp->sao[comp].off[3];
      const int nvalid = min(8, W - x);
      int v[8], out[8];
      unpack8(q, v);
      if (type == HMR_SAO_BO)
      {
        const int band = cp->sao[comp].band, sh = bd - 5;
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
          const int k = ((v[j] >> sh) - band) & 31;
          const int o = k == 0 ? o0 : (k == 1 ? o1 : (k == 2 ? o2 : o3));
          out[j] = k < 4 ? clip3i(0, maxv, v[j] + o) : v[j];
        }
      }
      else
      {
        const int bx = ctuX << cxs, by = ctuY << cys;
        const int bw = min(1 << cxs, W - bx), bh = min(1 << cys, H - by);
        const bool firstRow = y == by, lastRow = y == by + bh - 1;
        const int av = cp->avail;
        const bool L = av & HMR_AV_L, R = av & HMR_AV_R, A = av & HMR_AV_A, B = av & HMR_AV_B;
        // neighbours a[j], b[j] of sample j for this class: 10-sample rows, index j+1 = sample j
        int ra[10], rb[10];
        int da, db;                                               // column shift of a / b relative to the sample
        if (type == HMR_SAO_EO_0)
        {
          ra[0] = x > 0 ? p[-1] : 0; ra[9] = x + 8 < W ? p[8] : 0;
#pragma unroll
          for (int j = 0; j < 8; j++) ra[j + 1] = v[j];
#pragma unroll
          for (int j = 0; j < 10; j++) rb[j] = ra[j];
          da = -1; db = 1;
        }
        else
        {
          const bool up = y > 0, dn = y + 1 < H;
          const int16_t* pu = p - st; const int16_t* pd = p + st;
          int t[8];
          uint4 z = make_uint4(0, 0, 0, 0);
          unpack8(up ? *(const uint4*)pu : z, t);
#pragma unroll
          for (int j = 0; j < 8; j++) ra[j + 1] = t[j];
          unpack8(dn ? *(const uint4*)pd : z, t);
#pragma unroll
          for (int j = 0; j < 8; j++) rb[j + 1] = t[j];
          ra[0] = ra[9] = rb[0] = rb[9] = 0;
          if (type == HMR_SAO_EO_90) { da = 0; db = 0; }
          else if (type == HMR_SAO_EO_135)
          {
            if (up && x > 0) ra[0] = pu[-1];
            if (dn && x + 8 < W) rb[9] = pd[8];
            da = -1; db = 1;
          }
          else
          {
            if (up && x + 8 < W) ra[9] = pu[8];
            if (dn && x > 0) rb[0] = pd[-1];
            da = 1; db = -1;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
          const bool firstCol = x + j == bx, lastCol = x + j == bx + bw - 1;
          const bool colOk = !(firstCol && !L) && !(lastCol && !R);
          bool ok;
          switch (type)
          {
            case HMR_SAO_EO_0:  ok = colOk; break;
            case HMR_SAO_EO_90: ok = !(firstRow && !A) && !(lastRow && !B); break;
            case HMR_SAO_EO_135:
              if (firstRow && bh > 1) ok = firstCol ? (av & HMR_AV_AL) != 0 : (A && !(lastCol && !R));
              else if (lastRow)       ok = lastCol ? (av & HMR_AV_BR) != 0 : (B && !(firstCol && !L));
              else                    ok = colOk;
              break;
            default:
              if (firstRow && bh > 1) ok = lastCol ? (av & HMR_AV_AR) != 0 : (A && !(firstCol && !L));
              else if (lastRow)       ok = firstCol ? (av & HMR_AV_BL) != 0 : (B && !(lastCol && !R));
              else                    ok = colOk;
              break;
          }
          const int a = da < 0 ? ra[j] : (da == 0 ? ra[j + 1] : ra[j + 2]);
          const int b = db < 0 ? rb[j] : (db == 0 ? rb[j + 1] : rb[j + 2]);
          const int e = sgn3(v[j] - a) + sgn3(v[j] - b);
          const int o = e == -2 ? o0 : (e == -1 ? o1 : (e == 1 ? o2 : o3));
          out[j] = (ok && e != 0 && j < nvalid) ? clip3i(0, maxv, v[j] + o) : v[j];
        }
      }
      if (P.cu_flags)
      {
        // PCMLFDisableProcess / xPCMRestoration (TComSampleAdaptiveOffset.cpp:743-843): I_PCM (pcm_loop_filter_disabled) and
        // lossless CUs get their pre-filter samples back — SAO never applies to them (deblocking skipped them already)
        const uint8_t* frow = P.cu_flags + (size_t)((y << cy) >> 3) * P.w8;
#pragma unroll
        for (int j = 0; j < 8; j++) if (frow[((x + j) << cx) >> 3] & HMR_CU_NOFILTER) out[j] = v[j];
      }
      q.x = (uint32_t)(out[0] & 0xffff) | ((uint32_t)out[1] << 16);
      q.y = (uint32_t)(out[2] & 0xffff) | ((uint32_t)out[3] << 16);
      q.z = (uint32_t)(out[4] & 0xffff) | ((uint32_t)out[5] << 16);
      q.w = (uint32_t)(out[6] & 0xffff) | ((uint32_t)out[7] << 16);
    }
  }
  *(uint4*)(P.out.p[comp] + (size_t)y * P.out.pitch[comp] + x) = q;
}

void launch_sao(const FrameParams& P, cudaStream_t s)
{
  dim3 block(32, 8);
  dim3 grid((P.w[0] / 8 + 31) / 32, (P.h[0] + 7) / 8, P.hdr.chroma_format == HMR_CHROMA_400 ? 1 : 3);
  sao_kernel<<<grid, block, 0, s>>>(P);
}
