// hm_threadsafe.cpp — what it takes to run several HM parser instances in one process (one decoder per thread,
// all sharing one CUDA context):
//   * HM's ROM tables (initROM / destroyROM, TComRom.cpp:140-243) are created by every TDecTop::create and freed by
//     every TDecTop::destroy (TDecTop.cpp:106,131).  TDecTop.cpp is compiled with the two calls renamed to the
//     reference-counted versions below.
//   * TComSlice::m_prevTid0POC, TComDataCU::m_pcGlbArlCoeff and g_md5_mismatch are made thread_local at build time
//     (frontend/Makefile); hmdec_b200.cpp saves and restores the first and the last per decoder around every HM call.
//   * The SPS-dependent globals — g_bitDepth, g_maxTrDynamicRange, g_saoMaxOffsetQVal, g_uiMaxCUWidth/Height/Depth, g_uiAddCUDepth
//     (TComRom.cpp:245-252,319,542; written by TDecTop::xActivateParameterSets, TDecTop.cpp:323-333) and the z-scan /
//     raster / pel tables (written by TDecCu::create for every picture, TDecCu.cpp:95-100) — stay process-global (they are
//     read on every parser path; TLS in a dlopen'ed library would cost a call per access).  Instead every entry of the
//     wrapper into HM passes the GEOMETRY GATE below: calls of decoders whose streams agree on those values (the common
//     case: N streams of one format) run concurrently exactly as before, a call for a different geometry waits until the
//     others have left, re-binds the globals and tables to its own values and runs; so an 8-bit and a 10-bit stream, or
//     64x64 and 16x16 CTUs, decode correctly side by side (interleaved NAL by NAL) instead of corrupting each other.
#include <mutex>
#include <condition_variable>
#include <algorithm>
#include <cstdlib>
#include "TLibCommon/TComRom.h"
#include "TLibCommon/TComSlice.h"
#include "TLibCommon/TComChromaFormat.h"
#include "TLibCommon/TComSampleAdaptiveOffset.h"
#include "hm_threadsafe.h"
#include "hm_waitstats.h"
#include <cstdio>

thread_local bool t_hmwActive = true;

HmWaitStats& hm_wait_stats()
{
  static HmWaitStats s;
  static const bool init = [] {
    for (int i = 0; i < HMW_COUNT; i++) { s.ns[i] = 0; s.cpuNs[i] = 0; s.calls[i] = 0; }
    s.on = getenv("HMDEC_B200_STATS") != NULL;
    s.decoders = 0;
    s.skip = s.on ? atoi(getenv("HMDEC_B200_STATS")) : 0;
    if (s.on) atexit([] {
      static const char* name[HMW_COUNT] = { "new_decoder", "free_decoder", "sink submit (validate + pack + enqueue)", "wait for a picture's planes", "wait for hash verdicts", "geometry gate", "hash ring full (inside push)", "issue of the output DMA", "hash job submit", "picture completion (executeLoopFilters)", "push_nal_unit (total)" };
      fprintf(stderr, "hmdec_b200 wait stats (wall clock, all threads; decoders %d.. of %d):\n", s.skip + 1, s.decoders.load());
      for (int i = 0; i < HMW_COUNT; i++) fprintf(stderr, "  %-42s %9.3f s wall, %9.3f s on a CPU, %8lld calls\n", name[i], 1e-9 * (double)s.ns[i].load(), 1e-9 * (double)s.cpuNs[i].load(), s.calls[i].load());
    });
    return true; }();
  (void)init;
  return s;
}

static std::mutex g_romLock;
static int g_romUsers = 0;

Void hm_guarded_initROM()
{
  std::lock_guard<std::mutex> g(g_romLock);
  if (g_romUsers++ == 0) initROM();
}

Void hm_guarded_destroyROM()
{
  std::lock_guard<std::mutex> g(g_romLock);
  if (--g_romUsers == 0) destroyROM();
}

HmGeomKey hm_geom_key_of(TComSPS* sps)
{
  HmGeomKey k;
  if (!sps) return k;
  for (UInt ch = 0; ch < MAX_NUM_CHANNEL_TYPE; ch++)
  {
    k.bitDepth[ch] = sps->getBitDepth(ChannelType(ch));
    k.maxTrDynamicRange[ch] = sps->getUseExtendedPrecision() ? std::max<Int>(15, k.bitDepth[ch] + 6) : 15;
  }
  k.maxCUWidth  = sps->getMaxCUWidth();
  k.maxCUHeight = sps->getMaxCUHeight();
  k.maxCUDepth  = sps->getMaxCUDepth();
  k.addCUDepth  = std::max(0, sps->getLog2MinCodingBlockSize() - (Int)sps->getQuadtreeTULog2MinSize() + (Int)getMaxCUDepthOffset(sps->getChromaFormatIdc(), sps->getQuadtreeTULog2MinSize()));
  k.valid = k.maxCUWidth > 0;
  return k;
}

namespace {
std::mutex g_gateLock;
std::condition_variable g_gateCv;
HmGeomKey g_bound;            // the key the globals hold now (invalid: unknown)
int  g_users = 0;             // calls inside HM under g_bound
bool g_exclusive = false;     // a call with an unknown key is inside, alone
int  g_exclusiveWaiting = 0;  // calls with an unknown key that wait for the others to leave (newcomers queue behind them)
unsigned long g_rebinds = 0;

void bind(const HmGeomKey& k)
{
  for (UInt ch = 0; ch < MAX_NUM_CHANNEL_TYPE; ch++) { g_bitDepth[ch] = k.bitDepth[ch]; g_maxTrDynamicRange[ch] = k.maxTrDynamicRange[ch]; }
  // the truncation limit of sao_offset_abs, a global written by TComSampleAdaptiveOffset::create (TComSampleAdaptiveOffset.cpp:157)
  // and read by the PARSER (TDecSbac.cpp:1796): a 10-bit stream parsed under an 8-bit neighbour's limit loses CABAC sync
  for (UInt c = 0; c < MAX_NUM_COMPONENT; c++) g_saoMaxOffsetQVal[c] = (1u << (std::min<Int>(k.bitDepth[toChannelType(ComponentID(c))], MAX_SAO_TRUNCATED_BITDEPTH) - 5)) - 1;
  g_uiMaxCUWidth = k.maxCUWidth; g_uiMaxCUHeight = k.maxCUHeight; g_uiMaxCUDepth = k.maxCUDepth; g_uiAddCUDepth = k.addCUDepth;
  UInt* p = &g_auiZscanToRaster[0];
  initZscanToRaster(k.maxCUDepth, 1, 0, p);
  initRasterToZscan(k.maxCUWidth, k.maxCUHeight, k.maxCUDepth);
  initRasterToPelXY(k.maxCUWidth, k.maxCUHeight, k.maxCUDepth);
}
}

// HMDEC_B200_NO_GEOM_GATE=1 (A/B switch for tests): stock behaviour, every decoder scribbles over the shared globals.
static bool gateOff() { static const bool off = getenv("HMDEC_B200_NO_GEOM_GATE") != NULL; return off; }

void hm_geom_enter(const HmGeomKey& key)
{
  if (gateOff()) return;
  HmWaitScope ws(HMW_GEOM_GATE);
  std::unique_lock<std::mutex> l(g_gateLock);
  if (!key.valid)
  {
    g_exclusiveWaiting++;
    g_gateCv.wait(l, [] { return g_users == 0; });
    g_exclusiveWaiting--;
    g_exclusive = true;
    g_users = 1;
    return;
  }
  g_gateCv.wait(l, [&] { return !g_exclusive && !g_exclusiveWaiting && (g_users == 0 || g_bound == key); });
  if (!(g_bound == key))
  {
    if (g_bound.valid) g_rebinds++;
    bind(key);
    g_bound = key;
  }
  g_users++;
}

void hm_geom_leave(const HmGeomKey& activeNow)
{
  if (gateOff()) return;
  std::lock_guard<std::mutex> l(g_gateLock);
  if (g_exclusive) { g_exclusive = false; g_bound = activeNow; }   // whatever the call activated is what the globals hold
  if (--g_users == 0) g_gateCv.notify_all();
}

unsigned long hm_geom_rebinds() { std::lock_guard<std::mutex> l(g_gateLock); return g_rebinds; }
