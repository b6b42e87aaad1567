// k_md5.cu — MD5 of a decoded picture's planes as the SEI decoded-picture-hash defines it (TComPicYuvMD5.cpp:183-205,
// libmd5): per component, the samples in raster order, 1 byte per sample for bit depths <= 8, else 2 bytes little endian.
//
// MD5 is one serial chain per plane (64 dependent steps per 64-byte block, ~900 cycles/block on a GPU thread), so a
// single hash is slower here than on a CPU core — but the chains of different planes and pictures are independent and
// cost the host nothing: one WARP per plane, many pictures in flight on low-priority side streams, verdict delivered
// asynchronously.  A chain is cut into chunks of MD5_CHUNK blocks, one short kernel launch each (state carried in
// global memory), all enqueued at once: no launch runs longer than a few milliseconds, so hashes never hold up a
// hardware work queue or a context time slice that decode kernels are waiting for.  Inside the warp every lane carries the same state; the lanes split only the loading: lane l fetches
// message block (base + l) into registers one batch ahead, parks it in shared memory, and the chain reads its words as
// broadcasts.  The chain itself is LOP3 -> IADD3 -> SHF (rotate) -> IADD per step.
#include <algorithm>
#include "common.cuh"
#include "md5_service.h"

__constant__ uint32_t c_md5K[64] = {
  0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
  0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
  0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
  0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391 };

#define MD5_CHUNK 4096                 // 64-byte blocks per launch (~2 ms)

// Sample cursor of one lane: walks the plane in raster order without divisions.
struct Md5Cursor
{
  const int16_t* plane; int pitch, w; unsigned row, col;
  __device__ __forceinline__ void advance(unsigned samples) { col += samples; while (col >= (unsigned)w) { col -= w; row++; } }
};

// Load the 16 message words of the 64-byte block that starts at the cursor (does not move the cursor).
// `valid` = number of 32-bit words of the block that exist in the plane (the rest reads as zero).
template <bool TWO>
__device__ __forceinline__ void md5_load_block(const Md5Cursor& cur, int valid, uint32_t out[16])
{
  unsigned r = cur.row, c = cur.col;
#pragma unroll
  for (int j = 0; j < 16; j++)
  {
    uint32_t v = 0;
    if (j < valid)
    {
      const int16_t* p = cur.plane + (size_t)r * cur.pitch + c;
      if (TWO) v = *(const uint32_t*)p;                       // 2 samples, widths are even: a pair never straddles a row
      else
      {
        const uint2 q = *(const uint2*)p;                     // 4 samples, widths are multiples of 4
        v = (q.x & 0xff) | ((q.x >> 8) & 0xff00) | ((q.y & 0xff) << 16) | ((q.y << 8) & 0xff000000u);
      }
    }
    out[j] = v;
    c += TWO ? 2 : 4;
    if (c >= (unsigned)cur.w) { c -= cur.w; r++; }
  }
}

// One 64-byte block.  Per step the dependent chain is LOP3 -> IADD3 -> LEA.HI (rotate + add): the message word and the
// round constant are summed ahead of time, off the chain.
__device__ __forceinline__ void md5_block(uint32_t st[4], const uint32_t* __restrict__ m)
{
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = m[i];
#pragma unroll
  for (int i = 0; i < 64; i++)
  {
    int g, s;
    if (i < 16)      { g = i;                s = (i & 3) == 0 ? 7 : (i & 3) == 1 ? 12 : (i & 3) == 2 ? 17 : 22; }
    else if (i < 32) { g = (5 * i + 1) & 15; s = (i & 3) == 0 ? 5 : (i & 3) == 1 ? 9 : (i & 3) == 2 ? 14 : 20; }
    else if (i < 48) { g = (3 * i + 5) & 15; s = (i & 3) == 0 ? 4 : (i & 3) == 1 ? 11 : (i & 3) == 2 ? 16 : 23; }
    else             { g = (7 * i) & 15;     s = (i & 3) == 0 ? 6 : (i & 3) == 1 ? 10 : (i & 3) == 2 ? 15 : 21; }
    const uint32_t mk = x[g] + c_md5K[i];                     // independent of the state
    uint32_t f;
    if (i < 16)      f = (b & c) | (~b & d);
    else if (i < 32) f = (d & b) | (~d & c);
    else if (i < 48) f = b ^ c ^ d;
    else             f = c ^ (b | ~d);
    const uint32_t t = a + f + mk;
    a = d; d = c; c = b;
    b = b + __funnelshift_l(t, t, s);
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

template <bool TWO>
__device__ __forceinline__ void md5_plane(const Md5Job& J, int comp, int chunk, uint32_t* s_msg, int lane)
{
  const int w = J.w[comp];
  const unsigned long long nBytes = (unsigned long long)w * J.h[comp] * (TWO ? 2 : 1);
  const unsigned long long nWords = nBytes >> 2;              // nBytes is a multiple of 4
  const unsigned long long nFull = nBytes >> 6;               // complete 64-byte blocks
  constexpr unsigned SPB = TWO ? 32 : 64;                     // samples per block
  const unsigned long long first = (unsigned long long)chunk * MD5_CHUNK;
  if (first > nFull) return;                                  // this plane finished in an earlier chunk
  const unsigned long long last = min(first + MD5_CHUNK, nFull);   // blocks [first, last) in this launch
  uint32_t st[4];
  uint32_t* state = J.state + comp * 4;
  if (chunk == 0) { st[0] = 0x67452301u; st[1] = 0xefcdab89u; st[2] = 0x98badcfeu; st[3] = 0x10325476u; }
  else { st[0] = state[0]; st[1] = state[1]; st[2] = state[2]; st[3] = state[3]; }
  Md5Cursor cur; cur.plane = J.plane[comp]; cur.pitch = J.pitch[comp]; cur.w = w;
  {
    const unsigned long long smp = (first + lane) * SPB;      // lane l owns block base + l
    cur.row = (unsigned)(smp / (unsigned)w); cur.col = (unsigned)(smp - (unsigned long long)cur.row * w);
  }
  uint32_t nxt[16];
  md5_load_block<TWO>(cur, first + lane < last ? 16 : 0, nxt);
  for (unsigned long long base = first; base < last; base += 32)
  {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; j++) s_msg[lane * 16 + j] = nxt[j];
    __syncwarp();
    cur.advance(32 * SPB);                                    // prefetch the next batch while the chain runs
    md5_load_block<TWO>(cur, base + 32 + lane < last ? 16 : 0, nxt);
    const int cnt = (int)min((unsigned long long)32, last - base);
    for (int k = 0; k < cnt; k++) md5_block(st, s_msg + k * 16);
  }
  if (last < nFull || first + MD5_CHUNK == nFull)             // more full blocks (or exactly the tail) left for the next launch
  {
    if (lane == 0) { state[0] = st[0]; state[1] = st[1]; state[2] = st[2]; state[3] = st[3]; }
    return;
  }
  // tail: remaining bytes (a multiple of 4, < 64) + 0x80 + zero padding + bit length
  __syncwarp();
  const int tailWords = (int)(nWords - nFull * 16);           // 0..15
  s_msg[lane] = 0;
  __syncwarp();
  if (lane == 0)
  {
    Md5Cursor tc; tc.plane = cur.plane; tc.pitch = cur.pitch; tc.w = w;
    const unsigned long long smp = nFull * SPB;
    tc.row = (unsigned)(smp / (unsigned)w); tc.col = (unsigned)(smp - (unsigned long long)tc.row * w);
    uint32_t tw[16];
    md5_load_block<TWO>(tc, tailWords, tw);
#pragma unroll
    for (int j = 0; j < 16; j++) if (j < tailWords) s_msg[j] = tw[j];
    s_msg[tailWords] = 0x80u;
    const unsigned long long bits = nBytes * 8;
    if (tailWords < 14) { s_msg[14] = (uint32_t)bits; s_msg[15] = (uint32_t)(bits >> 32); }
    else                { s_msg[30] = (uint32_t)bits; s_msg[31] = (uint32_t)(bits >> 32); }
  }
  __syncwarp();
  md5_block(st, s_msg);
  if (tailWords >= 14) md5_block(st, s_msg + 16);
  if (lane == 0) { J.out[comp * 4 + 0] = st[0]; J.out[comp * 4 + 1] = st[1]; J.out[comp * 4 + 2] = st[2]; J.out[comp * 4 + 3] = st[3]; }
}

// One tick of the hash service: every in-flight job advances by one chunk (grid = jobs x 3 planes, one warp each).
__global__ void __launch_bounds__(32) md5_tick_kernel(const Md5TickJob* __restrict__ jobs)
{
  __shared__ uint32_t s_msg[32 * 16];
  const int job = blockIdx.x / 3, comp = blockIdx.x % 3, lane = threadIdx.x;
  const Md5Job J = jobs[job].J;
  const int chunk = jobs[job].chunk;
  if (comp >= J.ncomp) return;
  if (J.bd[comp] > 8) md5_plane<true>(J, comp, chunk, s_msg, lane);
  else                md5_plane<false>(J, comp, chunk, s_msg, lane);
}

void launch_md5_tick(const Md5TickJob* jobs, int n, cudaStream_t s)
{
  if (n > 0) md5_tick_kernel<<<3 * n, 32, 0, s>>>(jobs);
}

int md5_chunks(const Md5Job& J)
{
  // the plane with the most blocks decides the number of ticks; +1 so that the tail always has a tick
  unsigned long long maxBlocks = 0;
  for (int c = 0; c < J.ncomp; c++) maxBlocks = std::max(maxBlocks, ((unsigned long long)J.w[c] * J.h[c] * (J.bd[c] > 8 ? 2 : 1)) >> 6);
  return (int)(maxBlocks / MD5_CHUNK) + 1;
}
