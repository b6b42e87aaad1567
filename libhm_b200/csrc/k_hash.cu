// k_hash.cu — decoded-picture hashes that parallelise: 32-bit checksum and CRC-16 partials
// (TComPicYuvMD5.cpp:127-175; SEI decoded picture hash methods 3 and 2).  MD5 (method 1) is a serial chain and
// stays on the host over the planes fetched by hmr_read_plane.
//
// checksum: sum over samples of (low byte ^ mask) + (high byte ^ mask) [if bit depth > 8], mask = (x&0xff)^(y&0xff)^(x>>8)^(y>>8),
//           modulo 2^32 — an order-free sum, done with a grid-stride loop, warp shuffles and one atomicAdd per CTA.
// CRC:      CRC-16/CCITT (poly 0x1021, init 0xFFFF, 16 zero bits appended), MSB first over low byte then high byte of every
//           sample in raster order.  CRC is linear over GF(2): each thread computes the CRC register contribution of a
//           chunk of samples from a zero register, the host-side combine (engine.cu) shifts and XORs the chunk
//           registers in order (x^(8*bytes) mod P by square-and-multiply).
#include "common.cuh"

__global__ void checksum_kernel(const int16_t* __restrict__ plane, int w, int h, int pitch, int bd, uint32_t* out)
{
  uint32_t sum = 0;
  const int total = w * h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
  {
    const int y = i / w, x = i - y * w;
    const uint32_t mask = (uint32_t)((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8));
    const int v = plane[(size_t)y * pitch + x];
    sum += (uint32_t)((v & 0xff) ^ mask);
    if (bd > 8) sum += (uint32_t)((v >> 8) ^ mask);
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, sum);
}

// One thread per image row: CRC register of that row's bytes starting from a zero register.
__global__ void crc_rows_kernel(const int16_t* __restrict__ plane, int w, int h, int pitch, int bd, uint32_t* rowCrc)
{
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= h) return;
  uint32_t crc = 0;
  const int16_t* p = plane + (size_t)y * pitch;
  for (int x = 0; x < w; x++)
  {
    const int v = p[x];
    for (int byte = 0; byte < (bd > 8 ? 2 : 1); byte++)
    {
      const int b = byte ? (v >> 8) & 0xff : v & 0xff;
      for (int bit = 0; bit < 8; bit++)
      {
        const uint32_t msb = (crc >> 15) & 1, in = (uint32_t)(b >> (7 - bit)) & 1;
        crc = (((crc << 1) + in) & 0xffff) ^ (msb * 0x1021);
      }
    }
  }
  rowCrc[y] = crc;
}

void launch_hash(const PlaneSet& pic, const int w[3], const int h[3], const int bd[3], int type, uint32_t* d_out, uint32_t* d_scratch, cudaStream_t s)
{
  if (type == 3)
  {
    cudaMemsetAsync(d_out, 0, 3 * sizeof(uint32_t), s);
    for (int c = 0; c < 3; c++) if (w[c] && h[c]) checksum_kernel<<<296, 256, 0, s>>>(pic.p[c], w[c], h[c], pic.pitch[c], bd[c], d_out + c);
  }
  else
  {
    uint32_t* rows = d_scratch;
    for (int c = 0; c < 3; c++)
    {
      if (h[c] == 0) continue;                               // 4:0:0: no chroma planes
      crc_rows_kernel<<<(h[c] + 127) / 128, 128, 0, s>>>(pic.p[c], w[c], h[c], pic.pitch[c], bd[c], rows);
      rows += h[c];
    }
  }
}
