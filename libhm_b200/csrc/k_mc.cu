// k_mc.cu — inter prediction: 8-tap luma / 4-tap chroma fractional-sample interpolation with bi-prediction average.
//
// Replaces TComPrediction::motionCompensation -> xPredInterUni/xPredInterBi -> xPredInterBlk
// (TComPrediction.cpp:514-698), TComInterpolationFilter::filter / filterCopy (TComInterpolationFilter.cpp:94-251)
// and TComYuv::addAvg (TComYuv.cpp:336-391).
//
// One CTA per 16x16-luma tile of a PU (plus the co-located chroma tiles).  The reference window of each list is
// staged in shared memory with clamped coordinates, which is exactly HM's replicated picture border
// (TComPicYuv::extendPicBorder, TComPicYuv.cpp:173-217) without ever materialising it.  Every case of HM
// (copy / H only / V only / H+V, uni / bi) is evaluated as ONE separable pipeline H -> 14-bit intermediate -> V,
// using the identity tap set for a zero fraction; that is bit-identical to HM's special cases because the
// intermediate offset (8192 << s1) is a multiple of the first-stage divisor.
#include "common.cuh"

#define MC_T 16                     // luma tile edge
#define MC_WIN (MC_T + 7)           // 23: tile + 8-tap support
#define MC_LD 24                    // row pitch of the staged window

__constant__ int8_t c_lumaTaps[4][8] = { {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1} };
__constant__ int8_t c_chromaTaps[8][4] = { {0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4}, {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2} };

__global__ void __launch_bounds__(256) mc_kernel(const __grid_constant__ FrameParams P)
{
  __shared__ int16_t s_ref[MC_WIN * MC_LD];
  __shared__ int16_t s_tmp[MC_WIN * MC_T];
  __shared__ int s_pu;
  const int tid = threadIdx.x;
  const uint32_t tile = blockIdx.x;

  if (tid == 0)
  {
    int lo = 0, hi = (int)P.hdr.n_pu;                 // last PU with prefix[pu] <= tile
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.pu_prefix[mid] <= tile) lo = mid; else hi = mid; }
    s_pu = lo;
  }
  __syncthreads();
  const hmr_pu pu = P.pu[s_pu];
  const int t = (int)(tile - P.pu_prefix[s_pu]);
  const int tiles_x = (pu.w + MC_T - 1) / MC_T;
  const int lx0 = pu.x + MC_T * (t % tiles_x), ly0 = pu.y + MC_T * (t / tiles_x);
  const int lw = min(MC_T, pu.x + pu.w - lx0), lh = min(MC_T, pu.y + pu.h - ly0);
  const bool bi = pu.lists == (HMR_PU_L0 | HMR_PU_L1);

  for (int comp = 0; comp < 3; comp++)
  {
    const int cx = comp ? P.csx : 0, cy = comp ? P.csy : 0;
    const int x0 = lx0 >> cx, y0 = ly0 >> cy, w = lw >> cx, h = lh >> cy;
    const int ntaps = comp ? 4 : 8, half = ntaps / 2 - 1;
    const int bd = comp ? P.hdr.bit_depth_chroma : P.hdr.bit_depth_luma;
    const int headroom = max(2, 14 - bd);
    const int maxv = (1 << bd) - 1;
    const int ox = tid % w, oy = tid / w;            // the output sample this thread owns (w*h <= 256)
    const bool owner = tid < w * h;
    int val[2] = {0, 0};

    for (int list = 0; list < 2; list++)
    {
      if (!(pu.lists & (1 << list))) continue;       // uniform across the CTA
      const int slot = list ? (pu.slots >> 4) : (pu.slots & 15);
      const int16_t* __restrict__ ref = P.dpb[slot].p[comp];
      const int rpitch = P.dpb[slot].pitch[comp];
      const int mvx = pu.mv[list][0], mvy = pu.mv[list][1];
      const int sx = 2 + cx, sy = 2 + cy;
      const int ix = x0 + (mvx >> sx) - half, iy = y0 + (mvy >> sy) - half;
      const int fx = mvx & ((1 << sx) - 1), fy = mvy & ((1 << sy) - 1);
      const int8_t* tapx = comp ? c_chromaTaps[fx << (1 - cx)] : c_lumaTaps[fx];
      const int8_t* tapy = comp ? c_chromaTaps[fy << (1 - cy)] : c_lumaTaps[fy];
      const int rows = h + ntaps - 1, cols = w + ntaps - 1;
      const int wmax = P.w[comp] - 1, hmax = P.h[comp] - 1;

      for (int i = tid; i < rows * cols; i += 256)
      {
        const int r = i / cols, c = i - r * cols;
        const int yy = clip3i(0, hmax, iy + r), xx = clip3i(0, wmax, ix + c);
        s_ref[r * MC_LD + c] = ref[(size_t)yy * rpitch + xx];
      }
      __syncthreads();
      // horizontal: isFirst, !isLast  (shift = 6 - headroom, offset = -8192 << shift); result truncated to Pel
      const int s1 = 6 - headroom, o1 = -(8192 << s1);
      for (int i = tid; i < rows * w; i += 256)
      {
        const int r = i / w, c = i - r * w;
        int sum = 0;
        for (int k = 0; k < ntaps; k++) sum += tapx[k] * s_ref[r * MC_LD + c + k];
        s_tmp[r * MC_T + c] = (int16_t)((sum + o1) >> s1);
      }
      __syncthreads();
      // vertical: !isFirst, isLast = !bi
      if (owner)
      {
        int sum = 0;
        for (int k = 0; k < ntaps; k++) sum += tapy[k] * s_tmp[(oy + k) * MC_T + ox];
        const int s2 = bi ? 6 : 6 + headroom;
        const int o2 = bi ? 0 : (1 << (s2 - 1)) + (8192 << 6);
        int v = (int16_t)((sum + o2) >> s2);
        if (!bi) v = clip3i(0, maxv, v);
        val[list] = v;
      }
      __syncthreads();                               // s_ref / s_tmp are reused by the next list / component
    }

    if (owner)
    {
      int v;
      if (bi)
      {
        const int sh = headroom + 1, off = (1 << (sh - 1)) + 2 * 8192;   // TComYuv::addAvg
        v = clip3i(0, maxv, (val[0] + val[1] + off) >> sh);
      }
      else v = (pu.lists & HMR_PU_L0) ? val[0] : val[1];
      P.work.p[comp][(size_t)(y0 + oy) * P.work.pitch[comp] + x0 + ox] = (int16_t)v;
    }
  }
}

void launch_mc(const FrameParams& P, cudaStream_t s)
{
  if (P.hdr.n_mc_tiles == 0) return;
  mc_kernel<<<P.hdr.n_mc_tiles, 256, 0, s>>>(P);
}
