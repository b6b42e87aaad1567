// md5_service.cu — process-wide asynchronous MD5 service.
//
// An MD5 chain is serial (one warp, ~0.1 s for a 2160p luma plane), but chains are independent.  Giving every job its own
// stream does not scale: a process has at most 32 hardware work queues, so ~15 hashes could run at once — ~100 pictures/s
// where the decoders deliver 500.  Instead ONE host thread per device drives ONE stream with "ticks": each tick is a
// single short launch (k_md5.cu: md5_tick_kernel) in which every job in flight advances by MD5_CHUNK blocks, one warp
// per plane.  The latency of a job is (its number of chunks) x (tick period ~2 ms) whatever the number of jobs, so the
// throughput is bounded only by how many jobs the engines keep in flight (8 each).  No launch runs longer than a few
// milliseconds, nothing ever queues behind a hash, and the host thread sleeps when there is no work.
#include <pthread.h>
#include <sched.h>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "md5_service.h"

namespace {

struct Active { Md5Job J; int chunk, chunks; std::atomic<int>* done; cudaEvent_t ready; };

struct Service
{
  int device;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<Active> pending;
  bool running;                // a worker thread exists
  int failures;                // workers that ended on a CUDA error (diagnostics; the service restarts on the next submit)
  Service() : device(0), running(false), failures(0) {}
};

Service* const g_svc = new Service[16];   // never destroyed: a lingering worker may still hold its mutex at process exit

void worker(Service* sv)
{
  // the thread that happened to submit first may be pinned to one core: this worker must not inherit that
  cpu_set_t all;
  CPU_ZERO(&all);
  for (int c = 0; c < CPU_SETSIZE; c++) CPU_SET(c, &all);
  pthread_setaffinity_np(pthread_self(), sizeof(all), &all);
  cudaSetDevice(sv->device);
  cudaStream_t stream = nullptr;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);                   // lo = lowest priority
  const int MAXJ = 1024;                                        // jobs one tick advances (32 decoder threads x HMR_MD5_MAX_JOBS fit; more wait in `pending`)
  Md5TickJob* table[2] = { nullptr, nullptr };
  cudaEvent_t ev[2];
  bool evMade[2] = { false, false };
  bool ok = cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, lo) == cudaSuccess;
  for (int i = 0; i < 2 && ok; i++)
    ok = cudaMallocHost(&table[i], sizeof(Md5TickJob) * MAXJ) == cudaSuccess && (evMade[i] = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess);   // sleep, never spin: the cores belong to the parsers
  std::vector<Active> active;
  std::vector<std::atomic<int>*> finishing[2];                  // jobs whose last chunk was in tick t (signalled when tick t completes)
  unsigned long long t = 0;
  for (;;)
  {
    {
      std::unique_lock<std::mutex> lk(sv->mu);
      if (!ok)
      {
        // A CUDA call failed.  EVERY job this worker knows about gets its verdict (-1) — the ones still queued, the ones in
        // flight and the ones whose last chunk was already launched — so nobody spins on a `done` flag forever
        // (hmr_md5_result(wait), free_geometry).  The worker releases what it holds and ends; the next submit starts a fresh
        // worker (a sticky device error will make that one fail the same way, job by job, instead of hanging).
        for (size_t i = 0; i < sv->pending.size(); i++) sv->pending[i].done->store(-1);
        sv->pending.clear();
        for (size_t i = 0; i < active.size(); i++) active[i].done->store(-1);
        for (int k = 0; k < 2; k++) for (size_t i = 0; i < finishing[k].size(); i++) finishing[k][i]->store(-1);
        sv->failures++;
        sv->running = false;
        lk.unlock();
        for (int i = 0; i < 2; i++) { if (table[i]) cudaFreeHost(table[i]); if (evMade[i]) cudaEventDestroy(ev[i]); }
        if (stream) cudaStreamDestroy(stream);
        cudaGetLastError();
        return;
      }
      if (active.empty() && sv->pending.empty() && finishing[0].empty() && finishing[1].empty())
      {
        // idle: sleep until the next submit.  The worker never ends on its own: a worker that tore down its stream, events and
        // page-locked tables 200 ms after the last digest could do so while the process was already running its exit handlers
        // (the CUDA runtime's among them) — a crash at exit that also swallowed the still-buffered stdout of the harness.
        // A detached thread parked on a condition variable is harmless at exit (g_svc is never destroyed).
        sv->cv.wait(lk, [&] { return !sv->pending.empty(); });
      }
      while (!sv->pending.empty() && (int)active.size() < MAXJ)
      {
        active.push_back(sv->pending.back());
        sv->pending.pop_back();
        ok = ok && cudaStreamWaitEvent(stream, active.back().ready, 0) == cudaSuccess;   // the planes are complete before the first tick reads them
      }
    }
    const int cur = (int)(t & 1);
    if (!active.empty())
    {
      for (size_t i = 0; i < active.size(); i++) { table[cur][i].J = active[i].J; table[cur][i].chunk = active[i].chunk; }
      launch_md5_tick(table[cur], (int)active.size(), stream);
      ok = ok && cudaGetLastError() == cudaSuccess;
      size_t keep = 0;
      for (size_t i = 0; i < active.size(); i++)
      {
        if (++active[i].chunk >= active[i].chunks) finishing[cur].push_back(active[i].done);
        else active[keep++] = active[i];
      }
      active.resize(keep);
    }
    ok = ok && cudaEventRecord(ev[cur], stream) == cudaSuccess;
    // tick t-1 has certainly been overtaken once its event fires: publish its digests, and its table may be rewritten next round
    if (t > 0)
    {
      const int prev = cur ^ 1;
      ok = ok && cudaEventSynchronize(ev[prev]) == cudaSuccess;
      for (size_t i = 0; i < finishing[prev].size(); i++) finishing[prev][i]->store(ok ? 1 : -1);
      finishing[prev].clear();
    }
    t++;
  }
}

} // namespace

bool md5_service_submit(int device, const Md5Job& J, cudaEvent_t ready, std::atomic<int>* done)
{
  if (device < 0 || device >= 16) return false;
  Service& sv = g_svc[device];
  Active a;
  a.J = J; a.chunk = 0; a.chunks = md5_chunks(J); a.done = done; a.ready = ready;
  done->store(0);
  std::lock_guard<std::mutex> g(sv.mu);
  sv.device = device;
  sv.pending.push_back(a);
  if (!sv.running)
  {
    sv.running = true;
    std::thread(worker, &sv).detach();
  }
  sv.cv.notify_one();
  return true;
}
