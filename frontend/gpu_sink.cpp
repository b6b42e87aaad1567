// gpu_sink.cpp — HmFrameSink that feeds the B200 reconstruction engine through its C ABI (include/hmrecon.h).
// Pictures stay resident in HBM; HM's host planes are refreshed only on demand (plane access, MD5 hash check).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <list>
#include <map>
#include <iostream>
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicYuv.h"
#include "hm_emit.h"
#include "hmrecon.h"

class GpuSink : public HmFrameSink
{
public:
  GpuSink() : m_eng(NULL), m_verify(getenv("HMDEC_B200_VERIFY") != NULL)
  {
    int dev = 0;
    if (const char* d = getenv("HMDEC_B200_DEVICE")) dev = atoi(d);
    if (hmr_engine_create(&m_eng, dev) != HMR_OK || !m_eng)
    {
      fprintf(stderr, "hmdec_b200: cannot create the GPU reconstruction engine on CUDA device %d — there is no CPU fallback\n", dev);
      m_eng = NULL;
    }
  }
  ~GpuSink() { if (m_eng) hmr_engine_destroy(m_eng); }
  bool ok() const { return m_eng != NULL; }

  virtual void frameReady(const hmr_frame_desc& d, TComPic* pic)
  {
    if (hmr_submit_frame(m_eng, &d) != HMR_OK)
    {
      fprintf(stderr, "hmdec_b200: hmr_submit_frame failed: %s\n", hmr_error_string(m_eng));
      abort();
    }
    State& s = m_state[pic];
    s.slot = d.hdr->out_slot;
    s.hostStale = true;
    // the engine clamps coordinates instead of padding: spare HM the per-reference extendPicBorder() (TComSlice.cpp:350-376)
    if (!m_verify) pic->getPicYuvRec()->setBorderExtension(true);   // (HM's own CPU MC in verify mode needs the real border)
  }

  virtual void fetchPicture(TComPic* pic)
  {
    std::map<TComPic*, State>::iterator it = m_state.find(pic);
    if (it == m_state.end() || !it->second.hostStale) return;
    TComPicYuv* rec = pic->getPicYuvRec();
    std::vector<Pel> keep;
    for (int c = 0; c < 3; c++)
    {
      const ComponentID id = ComponentID(c);
      const int w = rec->getWidth(id), h = rec->getHeight(id), st = rec->getStride(id);
      if (m_verify) { keep.resize((size_t)w * h); for (int y = 0; y < h; y++) memcpy(&keep[(size_t)y * w], rec->getAddr(id) + (size_t)y * st, sizeof(Pel) * w); }
      if (hmr_read_plane(m_eng, it->second.slot, c, rec->getAddr(id), (size_t)st) != HMR_OK)
      {
        fprintf(stderr, "hmdec_b200: hmr_read_plane failed: %s\n", hmr_error_string(m_eng));
        abort();
      }
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

private:
  struct State { int slot; bool hostStale; };
  hmr_engine* m_eng;
  bool m_verify;
  std::map<TComPic*, State> m_state;
};

HmFrameSink* hm_new_gpu_sink()
{
  GpuSink* s = new GpuSink();
  if (!s->ok()) { delete s; return NULL; }
  return s;
}
