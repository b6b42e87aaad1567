// k_pack.cu — output wire format: the decoded picture as TAppDecoder's `-o` writes it.
//
// Replaces TVideoIOYuv::write / writePlane / scalePlane (Lib/TLibVideoIO/TVideoIOYuv.cpp:70-90, 706-790) for the case the
// decoder uses: conformance (+ default display) window crop, conversion from the internal to the output bit depth
// ((v + round) >> s clipped to [0, 2^out - 1] when going down, v << s when going up), planar Y / Cb / Cr, one byte per
// sample if every output bit depth is <= 8, else two bytes little endian.  The packed picture is what crosses PCIe:
// for 8-bit output that is half of the int16 planes.
#include "common.cuh"

struct PackParams
{
  const int16_t* src[3];
  int pitch[3], w[3], h[3], x0[3], y0[3];     // cropped size and origin per component
  int shift[3], maxv[3];                      // internal - output bit depth; clip maximum when shifting down
  size_t off[3];                              // byte offset of the component in the packed picture
  int twoBytes;
  uint8_t* dst;
};

__global__ void __launch_bounds__(256) pack_kernel(const PackParams Q)
{
  const int c = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= Q.w[c] || y >= Q.h[c]) return;
  int v = Q.src[c][(size_t)(Q.y0[c] + y) * Q.pitch[c] + Q.x0[c] + x];
  const int s = Q.shift[c];
  if (s > 0) v = min(Q.maxv[c], max(0, (v + (1 << (s - 1))) >> s));
  else if (s < 0) v <<= -s;
  const size_t i = (size_t)y * Q.w[c] + x;
  if (Q.twoBytes) ((uint16_t*)(Q.dst + Q.off[c]))[i] = (uint16_t)v;
  else Q.dst[Q.off[c] + i] = (uint8_t)v;
}

size_t launch_pack(const PlaneSet& pic, const int w[3], const int h[3], int csx, int csy, int ncomp, const int bdInternal[3], const int bdOut[3],
                   const int crop[4] /* left, right, top, bottom in luma samples */, uint8_t* d_dst, cudaStream_t s)
{
  PackParams Q;
  Q.twoBytes = 0;
  for (int c = 0; c < ncomp; c++) if (bdOut[c] > 8) Q.twoBytes = 1;
  size_t off = 0;
  int maxw = 1, maxh = 1;
  for (int c = 0; c < 3; c++)
  {
    const int sx = c ? csx : 0, sy = c ? csy : 0;
    Q.src[c] = pic.p[c]; Q.pitch[c] = pic.pitch[c];
    Q.x0[c] = crop[0] >> sx; Q.y0[c] = crop[2] >> sy;
    Q.w[c] = c < ncomp ? w[c] - ((crop[0] + crop[1]) >> sx) : 0;
    Q.h[c] = c < ncomp ? h[c] - ((crop[2] + crop[3]) >> sy) : 0;
    if (Q.w[c] < 0) Q.w[c] = 0;
    if (Q.h[c] < 0) Q.h[c] = 0;
    Q.shift[c] = bdInternal[c] - bdOut[c];
    Q.maxv[c] = (1 << bdOut[c]) - 1;
    Q.off[c] = off;
    off += (size_t)Q.w[c] * Q.h[c] * (Q.twoBytes ? 2 : 1);
    maxw = max(maxw, Q.w[c]); maxh = max(maxh, Q.h[c]);
  }
  Q.dst = d_dst;
  if (d_dst && off)
  {
    dim3 block(64, 4), grid((maxw + 63) / 64, (maxh + 3) / 4, ncomp);
    pack_kernel<<<grid, block, 0, s>>>(Q);
  }
  return off;
}
