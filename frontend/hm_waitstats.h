// hm_waitstats.h — HMDEC_B200_STATS: wall-clock time the decoder threads of the process spend in the places where they can BLOCK
// (waiting for the device or for each other), summed over all threads and printed once at exit.  The end-to-end rate is bounded by the
// host; this tells CPU work from waiting.
#ifndef HM_WAITSTATS_H
#define HM_WAITSTATS_H
#include <atomic>
#include <chrono>
#include <time.h>

enum HmWaitKind { HMW_NEW_DECODER = 0, HMW_FREE_DECODER, HMW_SUBMIT, HMW_PLANE_WAIT, HMW_HASH_WAIT, HMW_GEOM_GATE, HMW_HASH_RING, HMW_DMA_ISSUE, HMW_HASH_SUBMIT, HMW_PICTURE_DONE, HMW_PUSH_TOTAL, HMW_COUNT };

struct HmWaitStats
{
  std::atomic<long long> ns[HMW_COUNT];
  std::atomic<long long> cpuNs[HMW_COUNT];       // of which the calling thread was on a CPU (CLOCK_THREAD_CPUTIME_ID): wall - cpu = blocked
  std::atomic<long long> calls[HMW_COUNT];
  std::atomic<int> decoders;     // decoders created so far
  int skip;                      // HMDEC_B200_STATS=<n>: the first n decoders of the process (a harness's warm-up pass) are not counted
  bool on;
};
HmWaitStats& hm_wait_stats();
extern thread_local bool t_hmwActive;     // the decoder this thread is driving is being counted

struct HmWaitScope
{
  HmWaitKind k; std::chrono::steady_clock::time_point t0; long long c0; bool on;
  static long long threadCpuNs() { timespec ts; clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts); return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec; }
  explicit HmWaitScope(HmWaitKind kind) : k(kind), c0(0), on(hm_wait_stats().on && t_hmwActive) { if (on) { t0 = std::chrono::steady_clock::now(); c0 = threadCpuNs(); } }
  ~HmWaitScope()
  {
    if (!on) return;
    HmWaitStats& s = hm_wait_stats();
    s.ns[k] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    s.cpuNs[k] += threadCpuNs() - c0;
    s.calls[k] += 1;
  }
};
#endif
