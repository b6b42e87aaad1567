// gpu_sink.cpp — HmFrameSink that feeds the B200 reconstruction engine through its C ABI (include/hmrecon.h).
// Pictures stay resident in HBM; HM's host planes are refreshed only on demand (plane access, MD5 hash check).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <list>
#include <map>
#include <deque>
#include <iostream>
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicYuv.h"
#include <mutex>
#include "hm_emit.h"
#include "hm_fast.h"
#include "hm_waitstats.h"
#include "hmrecon.h"

// Process-wide pool of page-locked plane buffers (cudaMallocHost through the engine's C ABI).  Allocation is slow
// (milliseconds) and serialises on the driver, so buffers are never given back to the driver: a decoder that ends
// returns its planes to the pool and the next decoder (same geometry => same sizes) picks them up.
namespace {
std::mutex g_poolLock;
std::multimap<size_t, void*> g_poolFree;
std::map<void*, size_t> g_poolSize;

void* pooledPinnedAlloc(size_t bytes)
{
  {
    std::lock_guard<std::mutex> g(g_poolLock);
    std::multimap<size_t, void*>::iterator it = g_poolFree.find(bytes);
    if (it != g_poolFree.end()) { void* p = it->second; g_poolFree.erase(it); return p; }
  }
  void* p = hmr_alloc_pinned(bytes);
  if (p) { std::lock_guard<std::mutex> g(g_poolLock); g_poolSize[p] = bytes; }
  return p;
}

void pooledPinnedFree(void* p)
{
  std::lock_guard<std::mutex> g(g_poolLock);
  std::map<void*, size_t>::iterator it = g_poolSize.find(p);
  if (it != g_poolSize.end()) g_poolFree.insert(std::make_pair(it->second, p));
}
}

class GpuSink : public HmFrameSink
{
public:
  // m_verify (HM's CPU reconstruction next to the engine) exists in libHMDecoder_b200_verify.so only; the product has no CPU reconstruction
#ifdef HMDEC_WITH_HM_RECON
#define HMDEC_VERIFY_REQUESTED (getenv("HMDEC_B200_VERIFY") != NULL)
#else
#define HMDEC_VERIFY_REQUESTED false
#endif
  GpuSink() : m_eng(NULL), m_verify(HMDEC_VERIFY_REQUESTED), m_eager(getenv("HMDEC_B200_LAZY_PLANES") == NULL),
              m_gpuMd5(getenv("HMDEC_B200_HOST_MD5") == NULL), m_mismatch(false), m_jobs(0)
  {
    int dev = 0;
    if (const char* d = getenv("HMDEC_B200_DEVICE")) dev = atoi(d);
    if (hmr_engine_create(&m_eng, dev) != HMR_OK || !m_eng)
    {
      fprintf(stderr, "hmdec_b200: cannot create the GPU reconstruction engine on CUDA device %d — there is no CPU fallback\n", dev);
      m_eng = NULL;
    }
    else if (m_eager && !m_verify) hm_fast_set_plane_allocator(pooledPinnedAlloc, pooledPinnedFree);
  }
  ~GpuSink() { drainHashes(true); releaseHostBuffers(); if (m_eng) hmr_engine_destroy(m_eng); }
  bool ok() const { return m_eng != NULL; }

  // A failing engine call never takes the host process down: the sink latches the message (error()), the emitter stops
  // submitting and libHMDec_push_nal_unit answers LIBHMDEC_ERROR from then on.
  bool engineFailed(const char* call)
  {
    if (m_error.empty())
    {
      m_error = std::string(call) + ": " + hmr_error_string(m_eng);
      fprintf(stderr, "hmdec_b200: %s\n", m_error.c_str());
    }
    return false;
  }
  virtual const char* error() const { return m_error.empty() ? NULL : m_error.c_str(); }

  virtual bool frameReady(const hmr_frame_desc& d, TComPic* pic)
  {
    if (!m_error.empty()) return false;
    { HmWaitScope ws(HMW_SUBMIT); if (hmr_submit_frame(m_eng, &d) != HMR_OK) return engineFailed("hmr_submit_frame"); }
    State& s = m_state[pic];
    s.slot = d.hdr->out_slot;
    s.hostStale = true;
    s.copyIssued = false;
    // A picture that will be output is DMA'd into HM's own (page-locked) planes right behind its kernels, so that
    // libHMDEC_get_image_plane finds it there; pictures nobody will look at never leave the device.
    if (!m_verify && m_eager && pic->getSlice(0)->getPicOutputFlag())
    {
      HmWaitScope wsDma(HMW_DMA_ISSUE);
      TComPicYuv* rec = pic->getPicYuvRec();
      bool ok = true;
      for (int c = 0; c < (int)rec->getNumberValidComponents() && ok; c++)      // 4:0:0: the luma plane only
      {
        const ComponentID id = ComponentID(c);
        ok = hm_fast_plane_is_pinned(rec->getBuf(id)) &&
             hmr_read_plane_async(m_eng, s.slot, c, rec->getAddr(id), (size_t)rec->getStride(id)) == HMR_OK;
      }
      if (ok && hmr_marker_record(m_eng, &s.marker) == HMR_OK) s.copyIssued = true;
      else hmr_sync(m_eng);                                   // planes are not page-locked: copy on demand instead
    }
    // the engine clamps coordinates instead of padding: spare HM the per-reference extendPicBorder() (TComSlice.cpp:350-376)
    if (!m_verify) pic->getPicYuvRec()->setBorderExtension(true);   // (HM's own CPU MC in verify mode needs the real border)
    return true;
  }

  virtual void fetchPicture(TComPic* pic)
  {
    std::map<TComPic*, State>::iterator it = m_state.find(pic);
    if (it == m_state.end() || !it->second.hostStale) return;
    if (it->second.copyIssued)
    {
      if (hmr_marker_wait(m_eng, it->second.marker) != HMR_OK) { engineFailed("hmr_marker_wait"); return; }
      it->second.hostStale = false;
      return;
    }
    TComPicYuv* rec = pic->getPicYuvRec();
    std::vector<Pel> keep;
    for (int c = 0; c < (int)rec->getNumberValidComponents(); c++)
    {
      const ComponentID id = ComponentID(c);
      const int w = rec->getWidth(id), h = rec->getHeight(id), st = rec->getStride(id);
      if (m_verify) { keep.resize((size_t)w * h); for (int y = 0; y < h; y++) memcpy(&keep[(size_t)y * w], rec->getAddr(id) + (size_t)y * st, sizeof(Pel) * w); }
      if (hmr_read_plane(m_eng, it->second.slot, c, rec->getAddr(id), (size_t)st) != HMR_OK) { engineFailed("hmr_read_plane"); return; }
      if (m_verify)
        for (int y = 0; y < h; y++)
          if (memcmp(&keep[(size_t)y * w], rec->getAddr(id) + (size_t)y * st, sizeof(Pel) * w))
          {
            fprintf(stderr, "hmdec_b200 VERIFY: POC %d component %d row %d differs from HM's CPU reconstruction\n", pic->getPOC(), c, y);
            abort();
          }
    }
    it->second.hostStale = false;
  }

  virtual bool deviceHash(TComPic* pic, int method, uint32_t out[3])
  {
    std::map<TComPic*, State>::iterator it = m_state.find(pic);
    if (it == m_state.end() || (method != 2 && method != 3)) return false;
    return hmr_picture_hash(m_eng, it->second.slot, method, out) == HMR_OK;
  }

  virtual bool wantHmRecon() const { return m_verify; }

  virtual bool readPacked(TComPic* pic, const int outBitDepth[2], const int crop[4], void* dst, size_t capacity, size_t* bytes)
  {
    std::map<TComPic*, State>::iterator it = m_state.find(pic);
    if (it == m_state.end()) return false;
    return hmr_read_packed(m_eng, it->second.slot, outBitDepth, crop, dst, capacity, bytes) == HMR_OK;
  }

  virtual bool asyncMd5(TComPic* pic, const unsigned char* expected, int ncomp, const std::string& line, bool quiet)
  {
    std::map<TComPic*, State>::iterator it = m_state.find(pic);
    if (it == m_state.end() || m_verify || !m_gpuMd5) return false;
    if (m_jobs >= HMR_MD5_MAX_JOBS - 1) { HmWaitScope ws(HMW_HASH_RING); while (m_jobs >= HMR_MD5_MAX_JOBS - 1) deliverFront(true); }   // the engine's ring is full: the parser waits for the oldest digest
    Pending p;
    p.isJob = true; p.quiet = quiet; p.ncomp = ncomp; p.line = line;
    p.expected.assign(expected, expected + 16 * ncomp);
    { HmWaitScope ws(HMW_HASH_SUBMIT); if (hmr_md5_submit(m_eng, it->second.slot, &p.job) != HMR_OK) return false; }
    m_pending.push_back(p);
    m_jobs++;
    return true;
  }

  virtual void orderedPrint(const std::string& line)
  {
    if (m_pending.empty()) { fputs(line.c_str(), stdout); return; }
    Pending p;
    p.isJob = false; p.quiet = false; p.ncomp = 0; p.job = 0; p.line = line;
    m_pending.push_back(p);
  }

  virtual void drainHashes(bool wait) { while (!m_pending.empty() && deliverFront(wait)) { } }
  virtual bool hashMismatchSeen() const { return m_mismatch; }

  virtual void releaseHostBuffers() { if (m_eng) hmr_sync(m_eng); }

private:
  struct State { int slot; bool hostStale, copyIssued; uint64_t marker; };
  struct Pending { bool isJob, quiet; int ncomp; uint64_t job; std::string line; std::vector<unsigned char> expected; };

  // Deliver the oldest pending line; false = its digest is not there yet (only when !wait).
  bool deliverFront(bool wait)
  {
    Pending& p = m_pending.front();
    if (p.isJob)
    {
      unsigned char got[48];
      const int r = hmr_md5_result(m_eng, p.job, got, wait ? 1 : 0);
      if (r == HMR_PENDING) return false;
      m_jobs--;
      if (r != HMR_OK)
      {
        // the digest was lost (device error): the picture cannot be verified — report it like a mismatch, keep the order
        engineFailed("hmr_md5_result");
        m_mismatch = true;
        if (!p.quiet) printf("%s[MD5:<device error>,(***ERROR***)] \n", p.line.c_str());
        m_pending.pop_front();
        return true;
      }
      TComDigest dg, rx;
      dg.hash.assign(got, got + 16 * p.ncomp);
      rx.hash = p.expected;
      const bool bad = dg != rx;
      if (bad) m_mismatch = true;
      if (!p.quiet)
      {
        printf("%s[MD5:%s,%s] ", p.line.c_str(), digestToString(dg, 16).c_str(), bad ? "(***ERROR***)" : "(OK)");
        if (bad) printf("[rxMD5:%s] ", digestToString(rx, 16).c_str());
        printf("\n");
      }
    }
    else fputs(p.line.c_str(), stdout);
    m_pending.pop_front();
    return true;
  }

  hmr_engine* m_eng;
  bool m_verify;
  bool m_eager, m_gpuMd5, m_mismatch;
  std::string m_error;
  int m_jobs;
  std::map<TComPic*, State> m_state;
  std::deque<Pending> m_pending;
};

HmFrameSink* hm_new_gpu_sink()
{
  GpuSink* s = new GpuSink();
  if (!s->ok()) { delete s; return NULL; }
  return s;
}
