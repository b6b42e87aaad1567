// dump_sink.cpp — HmFrameSink that writes the per-frame records to a file, together with golden
// hashes of HM's own CPU reconstruction after each stage.  Tools/tests only (never on the product path).
//
// File layout (little endian):  "HMRDUMP1", then sections { u32 tag; u32 zero; u64 nbytes; payload padded to 8 }.
// A frame is the run of sections from 'HDR ' to 'END '.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <list>
#include <iostream>
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicYuv.h"
#include "hm_emit.h"

#define TAG(a, b, c, d) ((uint32_t)(a) | ((uint32_t)(b) << 8) | ((uint32_t)(c) << 16) | ((uint32_t)(d) << 24))

class DumpSink : public HmFrameSink
{
public:
  // HMDUMP_RECORDS_ONLY=1: the product's fast parse path (no HM reconstruction, coefficient hygiene on), records only —
  // lets a CPU-only test prove that the fast path emits byte-identical records.
  // The product library (no HMDEC_WITH_HM_RECON) cannot reconstruct on the CPU at all: its dumps are always records only.
#ifdef HMDEC_WITH_HM_RECON
#define HMDUMP_RECORDS_ONLY_DEFAULT (getenv("HMDUMP_RECORDS_ONLY") != NULL)
#else
#define HMDUMP_RECORDS_ONLY_DEFAULT true
#endif
  explicit DumpSink(const char* path) : m_fp(fopen(path, "wb")), m_planes(getenv("HMDUMP_PLANES") != NULL), m_recordsOnly(HMDUMP_RECORDS_ONLY_DEFAULT)
  {
    if (!m_fp) { perror(path); abort(); }
    fwrite("HMRDUMP1", 1, 8, m_fp);
    memset(m_gold, 0, sizeof(m_gold));
  }
  ~DumpSink() { if (m_fp) fclose(m_fp); }

  void section(uint32_t tag, const void* p, size_t n)
  {
    uint32_t h[2] = {tag, 0};
    uint64_t len = n;
    fwrite(h, 4, 2, m_fp);
    fwrite(&len, 8, 1, m_fp);
    if (n) fwrite(p, 1, n, m_fp);
    static const char zeros[8] = {0};
    if (n & 7) fwrite(zeros, 1, 8 - (n & 7), m_fp);
  }

  virtual bool frameReady(const hmr_frame_desc& d, TComPic*)
  {
    const hmr_frame_hdr& h = *d.hdr;
    const size_t nbs = (size_t)((h.width + 3) >> 2) * ((h.height + 3) >> 2);
    const size_t nqp = (size_t)((h.width + 7) >> 3) * ((h.height + 7) >> 3);
    section(TAG('H','D','R',' '), d.hdr, sizeof(hmr_frame_hdr));
    section(TAG('T','U',' ',' '), d.tu, sizeof(hmr_tu) * h.n_tu);
    section(TAG('C','O','E','F'), d.coef, sizeof(int16_t) * h.n_coef);
    section(TAG('I','N','T','R'), d.intra, sizeof(hmr_intra) * h.n_intra);
    section(TAG('I','R','N','G'), d.intra_range, sizeof(hmr_ctu_intra_range) * h.n_ctu);
    section(TAG('P','U',' ',' '), d.pu, sizeof(hmr_pu) * h.n_pu);
    section(TAG('P','U','P','F'), d.pu_tile_prefix, sizeof(uint32_t) * (h.n_pu + 1));
    section(TAG('C','T','U',' '), d.ctu, sizeof(hmr_ctu) * h.n_ctu);
    if (d.bs) section(TAG('B','S',' ',' '), d.bs, nbs);
    section(TAG('Q','P',' ',' '), d.qp, nqp);
    if (d.cu_flags) section(TAG('C','U','F','L'), d.cu_flags, nqp);
    if (d.scaling) section(TAG('S','C','A','L'), d.scaling, HMR_SCALING_BYTES);
    if (d.wp) { section(TAG('W','P',' ',' '), d.wp, sizeof(hmr_wp) * HMR_WP_ENTRIES); section(TAG('P','U','R','I'), d.pu_refidx, h.n_pu); }
    if (m_recordsOnly) { section(TAG('E','N','D',' '), NULL, 0); fflush(m_fp); }
    return true;
  }

  virtual void fetchPicture(TComPic*) {}
  virtual bool wantHmRecon() const { return !m_recordsOnly; }

  virtual void hmStage(int stage, TComPic* pic)
  {
    TComPicYuv& rec = *pic->getPicYuvRec();
    TComDigest digest;
    calcMD5(rec, digest);                       // TComPicYuvMD5.cpp:183-205: 3 x 16 bytes
    memset(m_gold[stage], 0, 48);               // absent components (4:0:0 chroma): all-zero digests
    memcpy(m_gold[stage], digest.hash.data(), 16 * rec.getNumberValidComponents());
    if (m_planes)
    {
      for (int c = 0; c < (int)rec.getNumberValidComponents(); c++)
      {
        const ComponentID id = ComponentID(c);
        const int w = rec.getWidth(id), h = rec.getHeight(id), s = rec.getStride(id);
        std::vector<int16_t> buf((size_t)w * h);
        const Pel* p = rec.getAddr(id);
        for (int y = 0; y < h; y++) memcpy(&buf[(size_t)y * w], p + (size_t)y * s, sizeof(int16_t) * w);
        section(TAG('P','0' + stage,'C','0' + c), buf.data(), buf.size() * 2);
      }
    }
    if (stage == 2)
    {
      section(TAG('G','O','L','D'), m_gold, sizeof(m_gold));
      section(TAG('E','N','D',' '), NULL, 0);
      fflush(m_fp);
    }
  }

private:
  FILE* m_fp;
  bool m_planes, m_recordsOnly;
  unsigned char m_gold[3][48];
};

HmFrameSink* hm_new_dump_sink(const char* path) { return new DumpSink(path); }

// Profiling aid (HMDEC_B200_DUMP=null): parse + record emission only.  Nothing is reconstructed, pictures are NOT valid
// and every hash check fails by construction — it exists to time the host side without a GPU.
class NullSink : public HmFrameSink
{
public:
  virtual bool frameReady(const hmr_frame_desc&, TComPic* pic) { pic->getPicYuvRec()->setBorderExtension(true); return true; }   // like the GPU sink: no border padding
  virtual void fetchPicture(TComPic*) {}
  virtual bool wantHmRecon() const { return false; }
};
HmFrameSink* hm_new_null_sink() { return new NullSink(); }
