// hm_threadsafe.cpp — what it takes to run several HM parser instances in one process (one decoder per thread,
// all sharing one CUDA context), for streams of identical SPS geometry:
//   * HM's ROM tables (initROM / destroyROM, TComRom.cpp:140-243) are created by every TDecTop::create and freed by
//     every TDecTop::destroy (TDecTop.cpp:106,131).  TDecTop.cpp is compiled with the two calls renamed to the
//     reference-counted versions below.
//   * TComSlice::m_prevTid0POC is made thread_local at build time (frontend/Makefile); hmdec_b200.cpp saves and
//     restores it per decoder around every TDecTop::decode call.
// Remaining shared state (g_bitDepth, g_uiMaxCU*, the z-scan tables) is rewritten with identical values by every
// decoder of the same geometry — which is why concurrent decoders must share the SPS geometry.
#include <mutex>
#include "TLibCommon/TComRom.h"

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
