// engine.cu — host side of libhmrecon.so: the C ABI of include/hmrecon.h, device memory management
// (DPB slots resident in HBM, pinned staging ring, compact residual buffer) and the per-picture kernel sequence.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <atomic>
#include <thread>
#include <chrono>
#include "common.cuh"
#include "md5_service.h"
#include "hmrecon.h"

#define RING 3
#define MARKERS 64
#define MD5_RING HMR_MD5_MAX_JOBS
#define ALIGN_UP(v, a) (((v) + (a) - 1) / (a) * (a))

struct Section { size_t off, bytes; };
struct Layout
{
  Section hdr, tu, coef, intra, irange, pu, prefix, ctu, bs, qp, cuf, scal, wp, puri;
  size_t total;
};

struct hmr_resident_frame
{
  uint8_t* dev;
  hmr_frame_hdr hdr;
  Layout lay;
  bool hasBs, hasCuf;
  IntraSizes intra;
};

struct FrameEvents { cudaEvent_t ev[HMR_T_COUNT + 1]; bool used[HMR_T_COUNT + 1]; };

struct hmr_engine
{
  int device;
  cudaStream_t stream;
  std::string err;
  // geometry (from the first frame header)
  bool haveGeom;
  int fmt, w[3], h[3], pitch[3], csx, csy, log2ctu, ctusW, ctusH, bdLuma, bdChroma;
  PlaneSet slots[HMR_MAX_SLOTS];
  bool slotAlloc[HMR_MAX_SLOTS];
  PlaneSet work;
  bool workAlloc;
  PlaneSet lastWork;           // where the last picture's pre-SAO state lives (the work picture, or its DPB slot when it had no SAO)
  // staging ring
  struct Stage { uint8_t* host; uint8_t* dev; size_t cap; cudaEvent_t done; bool inflight; } ring[RING];
  int ringPos;
  int16_t* resid; size_t residCap;
  hmr_pu* mcTiles; size_t mcTilesCap; uint8_t* mcTileRef;
  uint8_t* packBuf; size_t packCap;
  unsigned long long* progress; size_t progressCap;
  uint4* intraOps; size_t intraOpsCap; uint16_t* intraTab; size_t intraTabBytes; uint4* intraPrep; size_t intraPrepBytes;
  unsigned long long epoch;
  int stageMask;
  int validation;
  bool timing;
  std::vector<FrameEvents> pending, freeEvents;
  float accMs[HMR_T_COUNT];
  uint32_t accFrames, accLaunches;
  uint32_t* dHash; uint32_t* dHashRows; size_t hashRowsCap;
  void* flushBuf; size_t flushCap;
  int coopLimit;
  cudaEvent_t tBegin, tEnd, tJoin;
  bool timerInit;
  size_t planeSetBytes;
  // stream markers (hmr_marker_*)
  cudaEvent_t markers[MARKERS]; bool markerInit; uint64_t markerNext;
  // asynchronous MD5 (hmr_md5_*): side stream, ring of scratch pictures
  bool auxInit;
  struct Md5Slot { PlaneSet pic; bool alloc; cudaEvent_t copied; uint32_t* dState; uint32_t* hOut; uint64_t job; bool busy; std::atomic<int> done; } md5[MD5_RING];
  uint64_t md5Next;
};

// Process-wide caches of device and page-locked host buffers.  cudaMalloc / cudaFree / cudaMallocHost take milliseconds,
// serialise on the driver and (cudaFree) synchronise the whole device — poison when 16 decoder threads share one GPU and
// every new bitstream means a new engine.  Buffers of an engine that ends are parked here, keyed by (device, size), and
// the next engine of the same geometry picks them up; nothing is returned to the driver before the process exits.
#include <map>
static std::mutex g_poolLock;
static std::multimap<std::pair<int, size_t>, void*> g_devPool, g_hostPool;

static size_t g_devPooled[16] = {0}, g_hostPooled = 0;       // bytes parked, per device / page-locked
static size_t pool_cap(bool host)
{
  // parked bytes above the cap go back to the driver (HMR_POOL_CAP_MB / HMR_HOST_POOL_CAP_MB; defaults 48 GiB of 180 GiB HBM, 16 GiB pinned)
  static const size_t dev = (getenv("HMR_POOL_CAP_MB") ? (size_t)atoll(getenv("HMR_POOL_CAP_MB")) : (size_t)48 * 1024) << 20;
  static const size_t hst = (getenv("HMR_HOST_POOL_CAP_MB") ? (size_t)atoll(getenv("HMR_HOST_POOL_CAP_MB")) : (size_t)16 * 1024) << 20;
  return host ? hst : dev;
}
// Give every parked buffer of `device` (host: the page-locked ones) back to the driver.  Called when an allocation fails.
static void pool_trim(int device, bool host)
{
  std::vector<void*> victims;
  {
    std::lock_guard<std::mutex> g(g_poolLock);
    auto& pool = host ? g_hostPool : g_devPool;
    for (auto it = pool.begin(); it != pool.end();)
      if (host || it->first.first == device) { victims.push_back(it->second); (host ? g_hostPooled : g_devPooled[device & 15]) -= it->first.second; it = pool.erase(it); }
      else ++it;
  }
  for (void* v : victims) { if (host) cudaFreeHost(v); else cudaFree(v); }
}
static cudaError_t pool_malloc(int device, void** p, size_t bytes)
{
  {
    std::lock_guard<std::mutex> g(g_poolLock);
    auto it = g_devPool.find(std::make_pair(device, bytes));
    if (it != g_devPool.end()) { *p = it->second; g_devPool.erase(it); g_devPooled[device & 15] -= bytes; return cudaSuccess; }
  }
  cudaError_t r = cudaMalloc(p, bytes);
  if (r == cudaErrorMemoryAllocation)
  {
    cudaGetLastError();
    pool_trim(device, false);                       // buffers of other sizes (other resolutions decoded earlier) are in the way
    r = cudaMalloc(p, bytes);
  }
  return r;
}
static void pool_free(int device, void* p, size_t bytes)
{
  if (!p) return;
  {
    std::lock_guard<std::mutex> g(g_poolLock);
    if (g_devPooled[device & 15] + bytes <= pool_cap(false))
    {
      g_devPool.insert(std::make_pair(std::make_pair(device, bytes), p));
      g_devPooled[device & 15] += bytes;
      return;
    }
  }
  cudaFree(p);
}
static cudaError_t pool_malloc_host(void** p, size_t bytes)
{
  {
    std::lock_guard<std::mutex> g(g_poolLock);
    auto it = g_hostPool.find(std::make_pair(0, bytes));
    if (it != g_hostPool.end()) { *p = it->second; g_hostPool.erase(it); g_hostPooled -= bytes; return cudaSuccess; }
  }
  cudaError_t r = cudaMallocHost(p, bytes);
  if (r == cudaErrorMemoryAllocation) { cudaGetLastError(); pool_trim(0, true); r = cudaMallocHost(p, bytes); }
  return r;
}
static void pool_free_host(void* p, size_t bytes)
{
  if (!p) return;
  {
    std::lock_guard<std::mutex> g(g_poolLock);
    if (g_hostPooled + bytes <= pool_cap(true))
    {
      g_hostPool.insert(std::make_pair(std::make_pair(0, bytes), p));
      g_hostPooled += bytes;
      return;
    }
  }
  cudaFreeHost(p);
}

// CUDA streams are pooled the same way: creating one is a handful of ioctls that take ~20 ms when 16 decoder threads open new
// bitstreams against one context (tools/ioctl_trace.c, profiles/r04b_ioctl_trace.log: 3-5 % of a decoder thread at one new engine
// per 33-picture bitstream).  A stream goes back to the pool only after cudaStreamSynchronize, i.e. empty.
static std::vector<cudaStream_t>* const g_streamPool = new std::vector<cudaStream_t>[16];   // never destroyed (engines may end during static destruction)
static cudaError_t pool_stream(int device, cudaStream_t* s)
{
  {
    std::lock_guard<std::mutex> g(g_poolLock);
    if (device >= 0 && device < 16 && !g_streamPool[device].empty()) { *s = g_streamPool[device].back(); g_streamPool[device].pop_back(); return cudaSuccess; }
  }
  return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
}
static void pool_stream_release(int device, cudaStream_t s)
{
  if (!s) return;
  static const bool pooled = getenv("HMR_NO_STREAM_POOL") == NULL;          // A/B switch
  if (pooled && cudaStreamSynchronize(s) == cudaSuccess && device >= 0 && device < 16)
  {
    std::lock_guard<std::mutex> g(g_poolLock);
    if (g_streamPool[device].size() < 256) { g_streamPool[device].push_back(s); return; }
  }
  cudaStreamDestroy(s);
}

static int fail(hmr_engine* e, int code, const std::string& msg) { if (e) e->err = msg; return code; }
#define CK(call) do { cudaError_t _r = (call); if (_r != cudaSuccess) return fail(e, HMR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_r)); } while (0)

static Layout make_layout(const hmr_frame_hdr& h, bool hasBs, bool hasCuf, bool hasScal = false)
{
  Layout L;
  size_t off = 0;
  auto put = [&](Section& s, size_t bytes) { s.off = off; s.bytes = bytes; off = ALIGN_UP(off + bytes, 256); };
  const size_t nbs = (size_t)((h.width + 3) >> 2) * ((h.height + 3) >> 2);
  const size_t nqp = (size_t)((h.width + 7) >> 3) * ((h.height + 7) >> 3);
  put(L.hdr, sizeof(hmr_frame_hdr));
  put(L.tu, sizeof(hmr_tu) * h.n_tu);
  put(L.coef, sizeof(int16_t) * h.n_coef);
  put(L.intra, sizeof(hmr_intra) * h.n_intra);
  put(L.irange, sizeof(hmr_ctu_intra_range) * h.n_ctu);
  put(L.pu, sizeof(hmr_pu) * h.n_pu);
  put(L.prefix, sizeof(uint32_t) * (h.n_pu + 1));
  put(L.ctu, sizeof(hmr_ctu) * h.n_ctu);
  put(L.bs, hasBs ? nbs : 0);
  put(L.qp, nqp);
  put(L.cuf, hasCuf ? nqp : 0);
  put(L.scal, hasScal ? HMR_SCALING_BYTES : 0);
  const bool hasWp = (h.flags & HMR_FRM_WEIGHTED_PRED) != 0;
  put(L.wp, hasWp ? sizeof(hmr_wp) * HMR_WP_ENTRIES : 0);
  put(L.puri, hasWp ? h.n_pu : 0);
  L.total = off;
  return L;
}

static void pack(uint8_t* dst, const Layout& L, const hmr_frame_desc* f)
{
  memcpy(dst + L.hdr.off, f->hdr, L.hdr.bytes);
  if (L.tu.bytes)     memcpy(dst + L.tu.off, f->tu, L.tu.bytes);
  if (L.coef.bytes)   memcpy(dst + L.coef.off, f->coef, L.coef.bytes);
  if (L.intra.bytes)  memcpy(dst + L.intra.off, f->intra, L.intra.bytes);
  if (L.irange.bytes) memcpy(dst + L.irange.off, f->intra_range, L.irange.bytes);
  if (L.pu.bytes)     memcpy(dst + L.pu.off, f->pu, L.pu.bytes);
  memcpy(dst + L.prefix.off, f->pu_tile_prefix, L.prefix.bytes);
  if (L.ctu.bytes)    memcpy(dst + L.ctu.off, f->ctu, L.ctu.bytes);
  if (L.bs.bytes)     memcpy(dst + L.bs.off, f->bs, L.bs.bytes);
  if (L.qp.bytes)     memcpy(dst + L.qp.off, f->qp, L.qp.bytes);
  if (L.cuf.bytes)    memcpy(dst + L.cuf.off, f->cu_flags, L.cuf.bytes);
  if (L.scal.bytes)   memcpy(dst + L.scal.off, f->scaling, L.scal.bytes);
  if (L.wp.bytes)     memcpy(dst + L.wp.off, f->wp, L.wp.bytes);
  if (L.puri.bytes)   memcpy(dst + L.puri.off, f->pu_refidx, L.puri.bytes);
}

static int validate(hmr_engine* e, const hmr_frame_desc* f)
{
  if (!f || !f->hdr) return fail(e, HMR_ERR_ARG, "null frame");
  const hmr_frame_hdr& h = *f->hdr;
  if (h.magic != HMR_MAGIC || h.version != HMR_VERSION) return fail(e, HMR_ERR_FORMAT, "bad magic/version in frame header");
  if (h.width <= 0 || h.height <= 0 || (h.width & 7) || (h.height & 7)) return fail(e, HMR_ERR_FORMAT, "picture size must be a positive multiple of 8");
  if (h.chroma_format > HMR_CHROMA_444) return fail(e, HMR_ERR_FORMAT, "unsupported chroma format");      // 4:0:0 .. 4:4:4 (TypeDef.h ChromaFormat)
  if (h.log2_ctu < 4 || h.log2_ctu > 6) return fail(e, HMR_ERR_FORMAT, "unsupported CTU size");
  if (h.bit_depth_luma < 8 || h.bit_depth_luma > 12 || h.bit_depth_chroma < 8 || h.bit_depth_chroma > 12) return fail(e, HMR_ERR_FORMAT, "bit depth outside 8..12");
  if (h.out_slot >= HMR_MAX_SLOTS) return fail(e, HMR_ERR_FORMAT, "out_slot out of range");
  if (h.tu_first[4] != h.n_tu) return fail(e, HMR_ERR_FORMAT, "tu_first does not cover n_tu");
  if ((h.n_tu && (!f->tu || !f->coef)) || (h.n_intra && !f->intra) || (h.n_pu && !f->pu) || !f->pu_tile_prefix || !f->ctu || !f->intra_range || !f->qp)
    return fail(e, HMR_ERR_ARG, "missing record array");
  if ((h.flags & HMR_FRM_DEBLOCK) && !f->bs) return fail(e, HMR_ERR_ARG, "HMR_FRM_DEBLOCK without a BS map");
  if ((h.flags & HMR_FRM_SCALING_LIST) && !f->scaling) return fail(e, HMR_ERR_ARG, "HMR_FRM_SCALING_LIST without scaling factors");
  if ((h.flags & HMR_FRM_WEIGHTED_PRED) && (!f->wp || (h.n_pu && !f->pu_refidx))) return fail(e, HMR_ERR_ARG, "HMR_FRM_WEIGHTED_PRED without weights");
  // consistency the kernels rely on without re-checking (they index irange[row * ctusW + c], intra[first + k], tiles by prefix)
  const uint32_t ctusW = (uint32_t)((h.width + (1 << h.log2_ctu) - 1) >> h.log2_ctu), ctusH = (uint32_t)((h.height + (1 << h.log2_ctu) - 1) >> h.log2_ctu);
  if (h.n_ctu != ctusW * ctusH) return fail(e, HMR_ERR_FORMAT, "n_ctu does not match the picture size and CTU size");
  if (f->pu_tile_prefix[0] != 0 || f->pu_tile_prefix[h.n_pu] != h.n_mc_tiles) return fail(e, HMR_ERR_FORMAT, "pu_tile_prefix does not end at n_mc_tiles");
  if (h.n_coef & 15) return fail(e, HMR_ERR_FORMAT, "n_coef must be a multiple of 16");
  for (int k = 0; k < 4; k++) if (h.tu_first[k] > h.tu_first[k + 1]) return fail(e, HMR_ERR_FORMAT, "tu_first not ascending");
  for (uint32_t a = 0; a < h.n_ctu; a++)
    for (int c = 0; c < 3; c++)
    {
      const uint32_t first = f->intra_range[a].first[c], count = f->intra_range[a].count[c];
      if (count && (first > h.n_intra || count > h.n_intra - first)) return fail(e, HMR_ERR_FORMAT, "intra_range outside the intra records");
    }
  if (e->validation >= 2)
  {
    const int csx = (h.chroma_format == HMR_CHROMA_420 || h.chroma_format == HMR_CHROMA_422) ? 1 : 0, csy = h.chroma_format == HMR_CHROMA_420 ? 1 : 0;
    auto inside = [&](int comp, unsigned x, unsigned y, unsigned n) {
      const unsigned w = comp ? h.width >> csx : h.width, hh = comp ? h.height >> csy : h.height;
      return comp <= (h.chroma_format == HMR_CHROMA_400 ? 0 : 2) && x + n <= w && y + n <= hh;
    };
    for (uint32_t i = 0; i < h.n_tu; i++)
    {
      const hmr_tu& t = f->tu[i];
      const unsigned n = 1u << t.log2_size;
      if (t.log2_size < 2 || t.log2_size > 5 || !inside(t.comp, t.x, t.y, n)) return fail(e, HMR_ERR_FORMAT, "TU record outside the picture");
      if ((size_t)t.coef_off + (size_t)n * n > h.n_coef) return fail(e, HMR_ERR_FORMAT, "TU coef_off outside the coefficient buffer");
      if (t.luma_off != HMR_NO_OFFSET && (size_t)t.luma_off + (size_t)n * n > h.n_coef) return fail(e, HMR_ERR_FORMAT, "TU luma_off outside the coefficient buffer");
    }
    for (uint32_t i = 0; i < h.n_intra; i++)
    {
      const hmr_intra& r = f->intra[i];
      const unsigned n = 1u << r.log2_size;
      if (r.log2_size < 2 || r.log2_size > 5 || !inside(r.comp, r.x, r.y, n)) return fail(e, HMR_ERR_FORMAT, "intra record outside the picture");
      if (r.resid_off != HMR_NO_OFFSET && (size_t)r.resid_off + (size_t)n * n > h.n_coef) return fail(e, HMR_ERR_FORMAT, "intra resid_off outside the residual buffer");
    }
    for (uint32_t i = 0; i < h.n_pu; i++)
    {
      const hmr_pu& p = f->pu[i];
      if (!p.w || !p.h || (unsigned)p.x + p.w > (unsigned)h.width || (unsigned)p.y + p.h > (unsigned)h.height || !(p.lists & 3)) return fail(e, HMR_ERR_FORMAT, "PU record outside the picture");
      if (f->pu_tile_prefix[i + 1] - f->pu_tile_prefix[i] != (uint32_t)(((p.w + 15) >> 4) * ((p.h + 15) >> 4))) return fail(e, HMR_ERR_FORMAT, "pu_tile_prefix does not match the PU sizes");
    }
  }
  return HMR_OK;
}

static int alloc_planes(hmr_engine* e, PlaneSet& ps)
{
  size_t total = 0, offs[3];
  for (int c = 0; c < 3; c++) { offs[c] = total; total += ALIGN_UP((size_t)e->pitch[c] * e->h[c] * sizeof(int16_t), 512); }
  uint8_t* base = nullptr;
  CK(pool_malloc(e->device, (void**)&base, total));
  CK(cudaMemsetAsync(base, 0, total, e->stream));
  for (int c = 0; c < 3; c++) { ps.p[c] = (int16_t*)(base + offs[c]); ps.pitch[c] = e->pitch[c]; }
  e->planeSetBytes = total;
  return HMR_OK;
}

static void free_geometry(hmr_engine* e)
{
  cudaStreamSynchronize(e->stream);
  for (int s = 0; s < HMR_MAX_SLOTS; s++) if (e->slotAlloc[s]) { pool_free(e->device, e->slots[s].p[0], e->planeSetBytes); e->slotAlloc[s] = false; }
  if (e->workAlloc) { pool_free(e->device, e->work.p[0], e->planeSetBytes); e->workAlloc = false; }
  // digests in flight read their private copies: wait for the verdicts (the service always delivers one, -1 on failure); the
  // deadline only guards against a wedged device — the copies are then leaked rather than handed to the next engine
  bool md5Quiet = true;
  for (int i = 0; i < MD5_RING && e->auxInit; i++)
  {
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(20);
    while (e->md5[i].busy && e->md5[i].done.load() == 0 && std::chrono::steady_clock::now() < deadline) std::this_thread::sleep_for(std::chrono::microseconds(200));
    if (e->md5[i].busy && e->md5[i].done.load() == 0) md5Quiet = false;
  }
  if (!md5Quiet) { fprintf(stderr, "hmrecon: MD5 digests still in flight after 20 s; their buffers are not recycled\n"); for (int i = 0; i < MD5_RING; i++) e->md5[i].alloc = false; }
  for (int i = 0; i < MD5_RING; i++) if (e->md5[i].alloc) { pool_free(e->device, e->md5[i].pic.p[0], e->planeSetBytes); e->md5[i].alloc = false; }
  e->haveGeom = false;
}

static int ensure_geometry(hmr_engine* e, const hmr_frame_hdr& h)
{
  const int csx = (h.chroma_format == HMR_CHROMA_420 || h.chroma_format == HMR_CHROMA_422) ? 1 : 0;
  const int csy = h.chroma_format == HMR_CHROMA_420 ? 1 : 0;
  e->bdLuma = h.bit_depth_luma; e->bdChroma = h.bit_depth_chroma;
  if (e->haveGeom && (e->w[0] != h.width || e->h[0] != h.height || e->fmt != h.chroma_format || e->log2ctu != h.log2_ctu)) free_geometry(e);
  if (!e->haveGeom)
  {
    e->fmt = h.chroma_format; e->csx = csx; e->csy = csy; e->log2ctu = h.log2_ctu;
    for (int c = 0; c < 3; c++)
    {
      const bool absent = c && h.chroma_format == HMR_CHROMA_400;      // 4:0:0: the chroma planes are empty (TComPicYuv allocates none)
      e->w[c] = absent ? 0 : (c ? h.width >> csx : h.width);
      e->h[c] = absent ? 0 : (c ? h.height >> csy : h.height);
      e->pitch[c] = (int)ALIGN_UP((size_t)std::max(e->w[c], 1), 64);      // 128-byte rows
    }
    e->ctusW = (h.width + (1 << h.log2_ctu) - 1) >> h.log2_ctu;
    e->ctusH = (h.height + (1 << h.log2_ctu) - 1) >> h.log2_ctu;
    if (e->coopLimit < 1) return fail(e, HMR_ERR_CUDA, "the device cannot launch the intra wavefront cooperatively");
    const size_t need = (size_t)3 * e->ctusH + 1;           // per (component, row) progress + the job counter of the wavefront
    if (need > e->progressCap)
    {
      pool_free(e->device, e->progress, e->progressCap * sizeof(unsigned long long));
      CK(pool_malloc(e->device, (void**)&e->progress, need * sizeof(unsigned long long)));
      CK(cudaMemsetAsync(e->progress, 0, need * sizeof(unsigned long long), e->stream));
      e->progressCap = need;
    }
    {
      const size_t nctu = (size_t)e->ctusW * e->ctusH;
      pool_free(e->device, e->intraTab, e->intraTabBytes);
      pool_free(e->device, e->intraPrep, e->intraPrepBytes);
      e->intraTabBytes = intra_table_bytes((int)nctu);
      e->intraPrepBytes = 3 * nctu * sizeof(uint4);
      CK(pool_malloc(e->device, (void**)&e->intraTab, e->intraTabBytes));
      CK(pool_malloc(e->device, (void**)&e->intraPrep, e->intraPrepBytes));
    }
    int r = alloc_planes(e, e->work);
    if (r) return r;
    e->workAlloc = true;
    e->lastWork = e->work;
    e->haveGeom = true;
  }
  return HMR_OK;
}

static int ensure_slot(hmr_engine* e, int slot)
{
  if (e->slotAlloc[slot]) return HMR_OK;
  int r = alloc_planes(e, e->slots[slot]);
  if (r) return r;
  e->slotAlloc[slot] = true;
  return HMR_OK;
}

static void fill_params(hmr_engine* e, FrameParams& P, const hmr_frame_hdr& h, const Layout& L, uint8_t* dev, bool hasBs, bool hasCuf)
{
  memset(&P, 0, sizeof(P));
  P.hdr = h;
  for (int c = 0; c < 3; c++) { P.w[c] = e->w[c]; P.h[c] = e->h[c]; }
  P.csx = e->csx; P.csy = e->csy; P.ctus_w = e->ctusW; P.ctus_h = e->ctusH;
  P.w4 = (h.width + 3) >> 2; P.h4 = (h.height + 3) >> 2; P.w8 = (h.width + 7) >> 3;
  P.out = e->slots[h.out_slot];
  // No SAO anywhere in the picture: nothing ever needs the un-offset neighbours, so the picture is reconstructed and
  // deblocked IN PLACE in its DPB slot and the SAO pass (which would be a 2 x S_b copy) disappears.
  P.work = (h.flags & HMR_FRM_SAO) ? e->work : P.out;
  for (int s = 0; s < HMR_MAX_SLOTS; s++) P.dpb[s] = e->slotAlloc[s] ? e->slots[s] : e->slots[h.out_slot];
  P.tu = (const hmr_tu*)(dev + L.tu.off);
  P.coef = (const int16_t*)(dev + L.coef.off);
  P.resid = e->resid;
  P.mc_tiles = e->mcTiles;
  P.intra = (const hmr_intra*)(dev + L.intra.off);
  P.irange = (const hmr_ctu_intra_range*)(dev + L.irange.off);
  P.pu = (const hmr_pu*)(dev + L.pu.off);
  P.pu_prefix = (const uint32_t*)(dev + L.prefix.off);
  P.ctu = (const hmr_ctu*)(dev + L.ctu.off);
  P.bs = hasBs ? dev + L.bs.off : nullptr;
  P.qp = (const int8_t*)(dev + L.qp.off);
  P.cu_flags = hasCuf ? dev + L.cuf.off : nullptr;
  P.scaling = L.scal.bytes ? dev + L.scal.off : nullptr;
  P.wp = L.wp.bytes ? (const hmr_wp*)(dev + L.wp.off) : nullptr;
  P.pu_refidx = L.puri.bytes ? dev + L.puri.off : nullptr;
  P.mc_tile_refidx = e->mcTileRef;
  P.intra_progress = e->progress;
  P.intra_ops = e->intraOps; P.intra_tab = e->intraTab; P.intra_prep = e->intraPrep;
  P.intra_max_rec = P.intra_max_addr = P.intra_res_span = 0;      // 0 = worst case; the callers fill in what the records say
  P.intra_jobs = -1;
  P.epoch = e->epoch;
}

static int fold_timing(hmr_engine* e)
{
  if (e->pending.empty()) return HMR_OK;
  CK(cudaStreamSynchronize(e->stream));
  for (size_t i = 0; i < e->pending.size(); i++)
  {
    FrameEvents& fe = e->pending[i];
    for (int k = 0; k < HMR_T_COUNT; k++)
      if (fe.used[k] && fe.used[k + 1])
      {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, fe.ev[k], fe.ev[k + 1]) == cudaSuccess) e->accMs[k] += ms;
      }
    e->freeEvents.push_back(fe);
  }
  e->pending.clear();
  return HMR_OK;
}

// The per-picture kernel sequence.  `uploaded` = an H2D copy was enqueued just before (timing slot HMR_T_H2D).
static int run_frame(hmr_engine* e, FrameParams& P, FrameEvents* fe)
{
  const hmr_frame_hdr& h = P.hdr;
  if (h.n_coef > e->residCap)
  {
    CK(cudaStreamSynchronize(e->stream));
    pool_free(e->device, e->resid, e->residCap * sizeof(int16_t));
    e->residCap = ALIGN_UP((size_t)h.n_coef * 3 / 2 + 4096, 1 << 20);
    CK(pool_malloc(e->device, (void**)&e->resid, e->residCap * sizeof(int16_t)));
    P.resid = e->resid;
  }
  if (h.n_intra > e->intraOpsCap)
  {
    CK(cudaStreamSynchronize(e->stream));
    pool_free(e->device, e->intraOps, e->intraOpsCap * sizeof(uint4));
    e->intraOpsCap = ALIGN_UP((size_t)h.n_intra * 3 / 2 + 1024, 1 << 16);
    CK(pool_malloc(e->device, (void**)&e->intraOps, e->intraOpsCap * sizeof(uint4)));
    P.intra_ops = e->intraOps;
  }
  if (h.n_mc_tiles > e->mcTilesCap)
  {
    CK(cudaStreamSynchronize(e->stream));
    pool_free(e->device, e->mcTiles, e->mcTilesCap * sizeof(hmr_pu));
    pool_free(e->device, e->mcTileRef, e->mcTilesCap);
    e->mcTilesCap = ALIGN_UP((size_t)h.n_mc_tiles * 3 / 2 + 1024, 1 << 16);
    CK(pool_malloc(e->device, (void**)&e->mcTiles, e->mcTilesCap * sizeof(hmr_pu)));
    CK(pool_malloc(e->device, (void**)&e->mcTileRef, e->mcTilesCap));
    P.mc_tiles = e->mcTiles;
    P.mc_tile_refidx = e->mcTileRef;
  }
  auto mark = [&](int k) { if (fe) { cudaEventRecord(fe->ev[k], e->stream); fe->used[k] = true; } };
  const int m = e->stageMask;
  uint32_t launches = 0;
  mark(HMR_T_MC);
  if ((m & HMR_STAGE_MC) && h.n_mc_tiles) launches += launch_mc(P, e->stream);
  mark(HMR_T_RESID);
  if ((m & HMR_STAGE_RESID) && h.n_tu) launches += launch_resid(P, e->stream);
  mark(HMR_T_INTRA);
  if ((m & HMR_STAGE_INTRA) && h.n_intra) { CK(launch_intra(P, e->stream)); launches += 2; }
  mark(HMR_T_DEBLOCK_V);
  if ((m & HMR_STAGE_DEBLOCK_V) && (h.flags & HMR_FRM_DEBLOCK)) { launch_deblock(P, 0, e->stream); launches++; }
  mark(HMR_T_DEBLOCK_H);
  if ((m & HMR_STAGE_DEBLOCK_H) && (h.flags & HMR_FRM_DEBLOCK)) { launch_deblock(P, 1, e->stream); launches++; }
  mark(HMR_T_SAO);
  if ((m & HMR_STAGE_SAO) && (h.flags & HMR_FRM_SAO)) { launch_sao(P, e->stream); launches++; }
  e->lastWork = P.work;
  mark(HMR_T_COUNT);
  CK(cudaGetLastError());
  e->epoch++;
  e->accFrames++;
  e->accLaunches += launches;
  return HMR_OK;
}

static FrameEvents* grab_events(hmr_engine* e)
{
  if (!e->timing) return nullptr;
  if (e->pending.size() >= 512) fold_timing(e);
  FrameEvents fe;
  if (!e->freeEvents.empty()) { fe = e->freeEvents.back(); e->freeEvents.pop_back(); }
  else for (int k = 0; k <= HMR_T_COUNT; k++) cudaEventCreate(&fe.ev[k]);
  for (int k = 0; k <= HMR_T_COUNT; k++) fe.used[k] = false;
  e->pending.push_back(fe);
  return &e->pending.back();
}

extern "C" {

const char* hmr_version(void) { return "hmrecon-b200 0.1 (records v4)"; }

int hmr_engine_create(hmr_engine** out, int device)
{
  if (!out) return HMR_ERR_ARG;
  *out = nullptr;
  // decoder streams should not share hardware queues (effective if CUDA is not initialised yet); once, before any engine thread
  // can race on the environment
  static std::once_flag envOnce;
  std::call_once(envOnce, [] { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); });
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n)
  {
    fprintf(stderr, "hmrecon: no usable CUDA device %d (count %d): %s\n", device, n, cudaGetErrorString(cudaGetLastError()));
    return HMR_ERR_CUDA;
  }
  hmr_engine* e = new hmr_engine();
  e->device = device;
  e->haveGeom = false; e->workAlloc = false;
  memset(e->slotAlloc, 0, sizeof(e->slotAlloc));
  memset(e->ring, 0, sizeof(e->ring));
  e->ringPos = 0; e->resid = nullptr; e->residCap = 0; e->mcTiles = nullptr; e->mcTilesCap = 0; e->mcTileRef = nullptr; e->packBuf = nullptr; e->packCap = 0; e->progress = nullptr; e->progressCap = 0; e->epoch = 1;
  e->intraOps = nullptr; e->intraOpsCap = 0; e->intraTab = nullptr; e->intraTabBytes = 0; e->intraPrep = nullptr; e->intraPrepBytes = 0;
  e->stageMask = HMR_STAGE_ALL; e->timing = false; e->validation = 1;
  memset(e->accMs, 0, sizeof(e->accMs)); e->accFrames = e->accLaunches = 0;
  e->timerInit = false;
  e->planeSetBytes = 0; e->markerInit = false; e->markerNext = 0; e->auxInit = false; e->md5Next = 0;
  for (int i = 0; i < MD5_RING; i++) { e->md5[i].alloc = false; e->md5[i].busy = false; e->md5[i].dState = nullptr; e->md5[i].hOut = nullptr; e->md5[i].job = 0; e->md5[i].done.store(0); }
  e->dHash = nullptr; e->dHashRows = nullptr; e->hashRowsCap = 0; e->flushBuf = nullptr; e->flushCap = 0;
  if (cudaSetDevice(device) != cudaSuccess || pool_stream(device, &e->stream) != cudaSuccess)
  {
    fprintf(stderr, "hmrecon: cannot initialise device %d: %s\n", device, cudaGetErrorString(cudaGetLastError()));
    delete e;
    return HMR_ERR_CUDA;
  }
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  e->coopLimit = coop ? intra_max_coresident_blocks(device) : 0;
  for (int i = 0; i < RING; i++) cudaEventCreateWithFlags(&e->ring[i].done, cudaEventDisableTiming);
  *out = e;
  return HMR_OK;
}

void hmr_engine_destroy(hmr_engine* e)
{
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  free_geometry(e);
  for (int i = 0; i < RING; i++)
  {
    pool_free_host(e->ring[i].host, e->ring[i].cap);
    pool_free(e->device, e->ring[i].dev, e->ring[i].cap);
    cudaEventDestroy(e->ring[i].done);
  }
  pool_free(e->device, e->resid, e->residCap * sizeof(int16_t));
  pool_free(e->device, e->mcTiles, e->mcTilesCap * sizeof(hmr_pu));
  pool_free(e->device, e->mcTileRef, e->mcTilesCap);
  pool_free(e->device, e->packBuf, e->packCap);
  pool_free(e->device, e->progress, e->progressCap * sizeof(unsigned long long));
  pool_free(e->device, e->intraOps, e->intraOpsCap * sizeof(uint4));
  pool_free(e->device, e->intraTab, e->intraTabBytes);
  pool_free(e->device, e->intraPrep, e->intraPrepBytes);
  if (e->dHash) cudaFree(e->dHash);
  if (e->dHashRows) cudaFree(e->dHashRows);
  if (e->flushBuf) cudaFree(e->flushBuf);
  if (e->markerInit) for (int i = 0; i < MARKERS; i++) cudaEventDestroy(e->markers[i]);
  if (e->auxInit)
  {
    for (int i = 0; i < MD5_RING; i++)
    {
      cudaEventDestroy(e->md5[i].copied);
      pool_free(e->device, e->md5[i].dState, 12 * sizeof(uint32_t)); pool_free_host(e->md5[i].hOut, 12 * sizeof(uint32_t));
    }
  }
  fold_timing(e);
  for (size_t i = 0; i < e->freeEvents.size(); i++) for (int k = 0; k <= HMR_T_COUNT; k++) cudaEventDestroy(e->freeEvents[i].ev[k]);
  pool_stream_release(e->device, e->stream);
  delete e;
}

const char* hmr_error_string(const hmr_engine* e) { return e ? e->err.c_str() : "null engine"; }

int hmr_submit_frame(hmr_engine* e, const hmr_frame_desc* f)
{
  if (!e) return HMR_ERR_ARG;
  int r = validate(e, f);
  if (r) return r;
  CK(cudaSetDevice(e->device));
  const hmr_frame_hdr& h = *f->hdr;
  if ((r = ensure_geometry(e, h))) return r;
  if ((r = ensure_slot(e, h.out_slot))) return r;
  const bool hasBs = f->bs != nullptr, hasCuf = f->cu_flags != nullptr;
  const Layout L = make_layout(h, hasBs, hasCuf, (h.flags & HMR_FRM_SCALING_LIST) != 0);

  hmr_engine::Stage& st = e->ring[e->ringPos];
  e->ringPos = (e->ringPos + 1) % RING;
  if (st.inflight) { CK(cudaEventSynchronize(st.done)); st.inflight = false; }
  if (L.total > st.cap)
  {
    pool_free_host(st.host, st.cap);
    pool_free(e->device, st.dev, st.cap);
    st.host = nullptr; st.dev = nullptr;
    st.cap = ALIGN_UP(L.total * 3 / 2, 4 << 20);
    CK(pool_malloc_host((void**)&st.host, st.cap));
    CK(pool_malloc(e->device, (void**)&st.dev, st.cap));
  }
  pack(st.host, L, f);
  FrameEvents* fe = grab_events(e);
  if (fe) { cudaEventRecord(fe->ev[HMR_T_H2D], e->stream); fe->used[HMR_T_H2D] = true; }
  CK(cudaMemcpyAsync(st.dev, st.host, L.total, cudaMemcpyHostToDevice, e->stream));
  FrameParams P;
  fill_params(e, P, h, L, st.dev, hasBs, hasCuf);
  const IntraSizes iz = intra_sizes_host(h, f->intra, f->intra_range);
  P.intra_max_rec = iz.maxRec; P.intra_max_addr = iz.maxAddr; P.intra_res_span = iz.resSpan;
  P.intra_jobs = iz.jobs; memcpy(P.intra_job_mask, iz.jobMask, sizeof(P.intra_job_mask));
  r = run_frame(e, P, fe);
  CK(cudaEventRecord(st.done, e->stream));
  st.inflight = true;
  return r;
}

int hmr_sync(hmr_engine* e)
{
  if (!e) return HMR_ERR_ARG;
  CK(cudaStreamSynchronize(e->stream));
  return HMR_OK;
}

static int read_planeset(hmr_engine* e, const PlaneSet& ps, int comp, int16_t* dst, size_t dstStride, bool async)
{
  if (!e->haveGeom || comp < 0 || comp > 2) return fail(e, HMR_ERR_ARG, "read_plane: bad argument");
  if (e->w[comp] == 0 || e->h[comp] == 0) return HMR_OK;   // an absent component (4:0:0 chroma): nothing to copy
  if (!dst) return fail(e, HMR_ERR_ARG, "read_plane: bad argument");
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpy2DAsync(dst, dstStride * sizeof(int16_t), ps.p[comp], (size_t)ps.pitch[comp] * sizeof(int16_t),
                       (size_t)e->w[comp] * sizeof(int16_t), e->h[comp], cudaMemcpyDeviceToHost, e->stream));
  if (!async) CK(cudaStreamSynchronize(e->stream));
  return HMR_OK;
}

int hmr_read_plane(hmr_engine* e, int slot, int comp, int16_t* dst, size_t dst_stride)
{
  if (!e || slot < 0 || slot >= HMR_MAX_SLOTS || !e->slotAlloc[slot]) return fail(e, HMR_ERR_ARG, "read_plane: slot not allocated");
  return read_planeset(e, e->slots[slot], comp, dst, dst_stride, false);
}
int hmr_read_plane_async(hmr_engine* e, int slot, int comp, int16_t* dst, size_t dst_stride)
{
  if (!e || slot < 0 || slot >= HMR_MAX_SLOTS || !e->slotAlloc[slot]) return fail(e, HMR_ERR_ARG, "read_plane: slot not allocated");
  return read_planeset(e, e->slots[slot], comp, dst, dst_stride, true);
}
int hmr_read_work_plane(hmr_engine* e, int comp, int16_t* dst, size_t dst_stride)
{
  if (!e || !e->workAlloc) return fail(e, HMR_ERR_ARG, "read_work_plane: no picture yet");
  return read_planeset(e, e->lastWork, comp, dst, dst_stride, false);
}

int hmr_write_plane(hmr_engine* e, int slot, int comp, const int16_t* src, size_t src_stride, int width, int height)
{
  if (!e || slot < 0 || slot >= HMR_MAX_SLOTS || comp < 0 || comp > 2 || !src) return fail(e, HMR_ERR_ARG, "write_plane: bad argument");
  if (!e->haveGeom || width != e->w[comp] || height != e->h[comp]) return fail(e, HMR_ERR_ARG, "write_plane: geometry mismatch (submit a frame first)");
  CK(cudaSetDevice(e->device));
  int r = ensure_slot(e, slot);
  if (r) return r;
  CK(cudaMemcpy2DAsync(e->slots[slot].p[comp], (size_t)e->pitch[comp] * sizeof(int16_t), src, src_stride * sizeof(int16_t),
                       (size_t)width * sizeof(int16_t), height, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return HMR_OK;
}

// GF(2) polynomial arithmetic modulo the CRC-16/CCITT polynomial x^16 + x^12 + x^5 + 1
static uint32_t gf_mul(uint32_t a, uint32_t b)
{
  uint32_t r = 0;
  for (int i = 15; i >= 0; i--)
  {
    r = ((r << 1) & 0xffff) ^ (((r >> 15) & 1) * 0x1021);
    if ((b >> i) & 1) r ^= a;
  }
  return r;
}
static uint32_t gf_xpow(uint64_t n)
{
  uint32_t result = 1, base = 2;     // base = x
  while (n) { if (n & 1) result = gf_mul(result, base); base = gf_mul(base, base); n >>= 1; }
  return result;
}

int hmr_picture_hash(hmr_engine* e, int slot, int type, uint32_t out[3])
{
  if (!e || !out || slot < 0 || slot >= HMR_MAX_SLOTS || !e->slotAlloc[slot] || (type != 2 && type != 3)) return fail(e, HMR_ERR_ARG, "picture_hash: bad argument");
  CK(cudaSetDevice(e->device));
  if (!e->dHash) CK(cudaMalloc(&e->dHash, 3 * sizeof(uint32_t)));
  const size_t rows = (size_t)e->h[0] + e->h[1] + e->h[2];
  out[0] = out[1] = out[2] = 0;                             // absent components (4:0:0 chroma) report 0
  if (rows > e->hashRowsCap)
  {
    if (e->dHashRows) cudaFree(e->dHashRows);
    CK(cudaMalloc(&e->dHashRows, rows * sizeof(uint32_t)));
    e->hashRowsCap = rows;
  }
  const int bd[3] = { e->bdLuma, e->bdChroma, e->bdChroma };   // all pictures of a stream share the SPS bit depths
  launch_hash(e->slots[slot], e->w, e->h, bd, type, e->dHash, e->dHashRows, e->stream);
  CK(cudaGetLastError());
  if (type == 3)
  {
    CK(cudaMemcpyAsync(out, e->dHash, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return HMR_OK;
  }
  std::vector<uint32_t> rc(rows);
  CK(cudaMemcpyAsync(rc.data(), e->dHashRows, rows * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  size_t base = 0;
  for (int c = 0; c < 3; c++)
  {
    if (e->h[c] == 0) continue;
    const uint64_t rowBits = (uint64_t)e->w[c] * (bd[c] > 8 ? 16 : 8);
    const uint32_t shiftRow = gf_xpow(rowBits);
    uint32_t crc = 0xffff;
    for (int y = 0; y < e->h[c]; y++) crc = gf_mul(crc, shiftRow) ^ rc[base + y];
    crc = gf_mul(crc, gf_xpow(16));
    out[c] = crc;
    base += e->h[c];
  }
  return HMR_OK;
}

int hmr_set_validation(hmr_engine* e, int level) { if (!e) return HMR_ERR_ARG; e->validation = level; return HMR_OK; }
int hmr_set_stage_mask(hmr_engine* e, int mask) { if (!e) return HMR_ERR_ARG; e->stageMask = mask; return HMR_OK; }
int hmr_enable_timing(hmr_engine* e, int on) { if (!e) return HMR_ERR_ARG; if (!on) fold_timing(e); e->timing = on != 0; return HMR_OK; }

int hmr_get_stage_times(hmr_engine* e, float ms[HMR_T_COUNT], uint32_t* n_frames, uint32_t* n_launches)
{
  if (!e) return HMR_ERR_ARG;
  int r = fold_timing(e);
  if (r) return r;
  if (ms) memcpy(ms, e->accMs, sizeof(e->accMs));
  if (n_frames) *n_frames = e->accFrames;
  if (n_launches) *n_launches = e->accLaunches;
  memset(e->accMs, 0, sizeof(e->accMs));
  e->accFrames = e->accLaunches = 0;
  return HMR_OK;
}

int hmr_upload_frame(hmr_engine* e, const hmr_frame_desc* f, hmr_resident_frame** out)
{
  if (!e || !out) return HMR_ERR_ARG;
  int r = validate(e, f);
  if (r) return r;
  CK(cudaSetDevice(e->device));
  if ((r = ensure_geometry(e, *f->hdr))) return r;
  hmr_resident_frame* rf = new hmr_resident_frame();
  rf->hdr = *f->hdr;
  rf->hasBs = f->bs != nullptr; rf->hasCuf = f->cu_flags != nullptr;
  rf->lay = make_layout(rf->hdr, rf->hasBs, rf->hasCuf, (rf->hdr.flags & HMR_FRM_SCALING_LIST) != 0);
  rf->intra = intra_sizes_host(rf->hdr, f->intra, f->intra_range);
  std::vector<uint8_t> tmp(rf->lay.total);
  pack(tmp.data(), rf->lay, f);
  cudaError_t ce = cudaMalloc(&rf->dev, rf->lay.total);
  if (ce != cudaSuccess) { delete rf; return fail(e, HMR_ERR_NOMEM, cudaGetErrorString(ce)); }
  CK(cudaMemcpy(rf->dev, tmp.data(), rf->lay.total, cudaMemcpyHostToDevice));
  *out = rf;
  return HMR_OK;
}

int hmr_run_resident(hmr_engine* e, const hmr_resident_frame* f)
{
  if (!e || !f) return HMR_ERR_ARG;
  CK(cudaSetDevice(e->device));
  int r = ensure_geometry(e, f->hdr);
  if (r) return r;
  if ((r = ensure_slot(e, f->hdr.out_slot))) return r;
  FrameParams P;
  fill_params(e, P, f->hdr, f->lay, f->dev, f->hasBs, f->hasCuf);
  P.intra_max_rec = f->intra.maxRec; P.intra_max_addr = f->intra.maxAddr; P.intra_res_span = f->intra.resSpan;
  P.intra_jobs = f->intra.jobs; memcpy(P.intra_job_mask, f->intra.jobMask, sizeof(P.intra_job_mask));
  return run_frame(e, P, grab_events(e));
}

int hmr_run_resident_list(hmr_engine* e, hmr_resident_frame* const* frames, int n)
{
  if (!e || (!frames && n > 0)) return HMR_ERR_ARG;
  for (int i = 0; i < n; i++) { int r = hmr_run_resident(e, frames[i]); if (r) return r; }
  return HMR_OK;
}

static int timer_init(hmr_engine* e)
{
  if (e->timerInit) return HMR_OK;
  CK(cudaEventCreate(&e->tBegin)); CK(cudaEventCreate(&e->tEnd)); CK(cudaEventCreateWithFlags(&e->tJoin, cudaEventDisableTiming));
  e->timerInit = true;
  return HMR_OK;
}

int hmr_timer_begin(hmr_engine* e)
{
  if (!e) return HMR_ERR_ARG;
  CK(cudaSetDevice(e->device));
  int r = timer_init(e); if (r) return r;
  CK(cudaEventRecord(e->tBegin, e->stream));
  return HMR_OK;
}

int hmr_timer_join(hmr_engine* e, hmr_engine* other)
{
  if (!e || !other || e->device != other->device) return HMR_ERR_ARG;
  if (e == other) return HMR_OK;
  CK(cudaSetDevice(e->device));
  { hmr_engine* keep = e; e = other; int r = timer_init(e); e = keep; if (r) return r; }
  CK(cudaEventRecord(other->tJoin, other->stream));
  CK(cudaStreamWaitEvent(e->stream, other->tJoin, 0));
  return HMR_OK;
}

int hmr_timer_end(hmr_engine* e, float* ms)
{
  if (!e || !ms || !e->timerInit) return HMR_ERR_ARG;
  CK(cudaSetDevice(e->device));
  CK(cudaEventRecord(e->tEnd, e->stream));
  CK(cudaEventSynchronize(e->tEnd));
  CK(cudaEventElapsedTime(ms, e->tBegin, e->tEnd));
  return HMR_OK;
}

void hmr_free_resident(hmr_engine* e, hmr_resident_frame* f)
{
  if (!f) return;
  if (e) { cudaSetDevice(e->device); cudaStreamSynchronize(e->stream); }
  cudaFree(f->dev);
  delete f;
}

int hmr_host_register(void* p, size_t bytes)
{
  if (!p || !bytes) return HMR_ERR_ARG;
  return cudaHostRegister(p, bytes, cudaHostRegisterPortable) == cudaSuccess ? HMR_OK : HMR_ERR_CUDA;
}
int hmr_host_unregister(void* p) { return (p && cudaHostUnregister(p) == cudaSuccess) ? HMR_OK : HMR_ERR_CUDA; }

int hmr_marker_record(hmr_engine* e, uint64_t* id)
{
  if (!e || !id) return HMR_ERR_ARG;
  CK(cudaSetDevice(e->device));
  if (!e->markerInit)
  {
    for (int i = 0; i < MARKERS; i++) CK(cudaEventCreateWithFlags(&e->markers[i], cudaEventDisableTiming));
    e->markerInit = true;
  }
  const uint64_t m = e->markerNext++;
  cudaEvent_t ev = e->markers[m % MARKERS];
  if (m >= MARKERS) CK(cudaEventSynchronize(ev));            // the marker this slot held before is MARKERS records old
  CK(cudaEventRecord(ev, e->stream));
  *id = m;
  return HMR_OK;
}

int hmr_marker_wait(hmr_engine* e, uint64_t id)
{
  if (!e || id >= e->markerNext) return HMR_ERR_ARG;
  if (id + MARKERS < e->markerNext) return HMR_OK;            // overwritten => it completed long ago (see hmr_marker_record)
  CK(cudaEventSynchronize(e->markers[id % MARKERS]));
  return HMR_OK;
}

static int aux_init(hmr_engine* e)
{
  if (e->auxInit) return HMR_OK;
  for (int i = 0; i < MD5_RING; i++)
  {
    CK(cudaEventCreateWithFlags(&e->md5[i].copied, cudaEventDisableTiming));
    CK(pool_malloc(e->device, (void**)&e->md5[i].dState, 12 * sizeof(uint32_t)));   // chaining state between ticks
    CK(pool_malloc_host((void**)&e->md5[i].hOut, 12 * sizeof(uint32_t)));           // digest, written by the device
    e->md5[i].busy = false; e->md5[i].alloc = false;
  }
  e->auxInit = true;
  return HMR_OK;
}

int hmr_md5_submit(hmr_engine* e, int slot, uint64_t* job)
{
  if (!e || !job || slot < 0 || slot >= HMR_MAX_SLOTS || !e->slotAlloc[slot]) return fail(e, HMR_ERR_ARG, "md5_submit: bad argument");
  CK(cudaSetDevice(e->device));
  int r = aux_init(e);
  if (r) return r;
  hmr_engine::Md5Slot& m = e->md5[e->md5Next % MD5_RING];
  if (m.busy) return fail(e, HMR_ERR_BUSY, "md5_submit: ring full, collect results first");
  if (!m.alloc) { if ((r = alloc_planes(e, m.pic))) return r; m.alloc = true; }
  // private copy: the DPB slot may be overwritten long before the chain has walked the picture
  CK(cudaMemcpyAsync(m.pic.p[0], e->slots[slot].p[0], e->planeSetBytes, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaEventRecord(m.copied, e->stream));
  Md5Job J;
  const int bd[3] = { e->bdLuma, e->bdChroma, e->bdChroma };
  for (int c = 0; c < 3; c++) { J.plane[c] = m.pic.p[c]; J.pitch[c] = m.pic.pitch[c]; J.w[c] = e->w[c]; J.h[c] = e->h[c]; J.bd[c] = bd[c]; }
  J.ncomp = e->fmt == HMR_CHROMA_400 ? 1 : 3;
  J.out = m.hOut;
  J.state = m.dState;
  if (!md5_service_submit(e->device, J, m.copied, &m.done)) return fail(e, HMR_ERR_CUDA, "md5_submit: hash service unavailable");
  m.busy = true;
  m.job = e->md5Next++;
  *job = m.job;
  return HMR_OK;
}

int hmr_md5_result(hmr_engine* e, uint64_t job, uint8_t out[48], int wait)
{
  if (!e || !out || !e->auxInit) return HMR_ERR_ARG;
  hmr_engine::Md5Slot& m = e->md5[job % MD5_RING];
  if (!m.busy || m.job != job) return fail(e, HMR_ERR_ARG, "md5_result: unknown or already collected job");
  int st = m.done.load(std::memory_order_acquire);
  if (st == 0 && !wait) return HMR_PENDING;
  while (st == 0) { std::this_thread::sleep_for(std::chrono::microseconds(200)); st = m.done.load(std::memory_order_acquire); }
  m.busy = false;
  if (st < 0) return fail(e, HMR_ERR_CUDA, "md5_result: the hash service failed");
  memcpy(out, m.hOut, 48);
  return HMR_OK;
}

int hmr_read_packed(hmr_engine* e, int slot, const int out_bit_depth[2], const int crop[4], void* dst, size_t capacity, size_t* bytes)
{
  if (!e || slot < 0 || slot >= HMR_MAX_SLOTS || !e->slotAlloc[slot] || !out_bit_depth || !crop || !bytes) return fail(e, HMR_ERR_ARG, "read_packed: bad argument");
  CK(cudaSetDevice(e->device));
  const int ncomp = e->fmt == HMR_CHROMA_400 ? 1 : 3;
  const int bdIn[3] = { e->bdLuma, e->bdChroma, e->bdChroma };
  const int bdOut[3] = { out_bit_depth[0] > 0 ? out_bit_depth[0] : e->bdLuma, out_bit_depth[1] > 0 ? out_bit_depth[1] : e->bdChroma, out_bit_depth[1] > 0 ? out_bit_depth[1] : e->bdChroma };
  for (int c = 0; c < 3; c++) if (bdOut[c] < 1 || bdOut[c] > 16) return fail(e, HMR_ERR_ARG, "read_packed: output bit depth outside 1..16");
  const size_t need = launch_pack(e->slots[slot], e->w, e->h, e->csx, e->csy, ncomp, bdIn, bdOut, crop, nullptr, e->stream);   // size only
  *bytes = need;
  if (!dst) return HMR_OK;
  if (need > capacity) return fail(e, HMR_ERR_ARG, "read_packed: destination too small");
  if (need > e->packCap)
  {
    pool_free(e->device, e->packBuf, e->packCap);
    e->packCap = ALIGN_UP(need, 1 << 20);
    CK(pool_malloc(e->device, (void**)&e->packBuf, e->packCap));
  }
  launch_pack(e->slots[slot], e->w, e->h, e->csx, e->csy, ncomp, bdIn, bdOut, crop, e->packBuf, e->stream);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst, e->packBuf, need, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return HMR_OK;
}

void* hmr_alloc_pinned(size_t bytes) { void* p = nullptr; return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr; }
void  hmr_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int hmr_flush_l2(hmr_engine* e, size_t bytes)
{
  if (!e) return HMR_ERR_ARG;
  CK(cudaSetDevice(e->device));
  if (bytes > e->flushCap)
  {
    if (e->flushBuf) cudaFree(e->flushBuf);
    CK(cudaMalloc(&e->flushBuf, bytes));
    e->flushCap = bytes;
  }
  CK(cudaMemsetAsync(e->flushBuf, 0x5a, bytes, e->stream));
  return HMR_OK;
}

} // extern "C"
