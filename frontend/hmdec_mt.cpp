// hmdec_mt.cpp — throughput harness: T decoder threads in ONE process (one libHMDec_* decoder per thread, each with
// its own engine/CUDA stream, all sharing one CUDA context), every thread decoding the given Annex-B stream R times
// through the public entry points exactly as a YUView-style caller would (push NAL, re-push on bNewPicture, drain
// pictures, touch every plane).  Prints one JSON line with the wall time of the steady-state part.
// --overlap-verdict: a thread collects the verdict of a bitstream (every SEI hash verified, nothing unsupported) and frees its
// decoder only after it has pushed the NEXT bitstream through a second decoder: the last pictures' MD5 chains (0.13 s for 2160p
// planes, on the device) then finish while the thread is already parsing again, the way a multi-stream server would pipeline it.
// Default: check and free immediately.
// Several -b: thread t decodes stream t % (number of streams) — BASELINE.json configs[4], distinct bitstreams decoded concurrently.
// --sum-planes: the caller reads EVERY sample of every returned plane (a 64-bit sum per picture, printed as "plane_sum"), not just
// one per plane: proves the planes can be streamed out of the page-locked buffers at the reported rate.
//   hmdec_mt -b in.bin [-b in2.bin ...] [--threads T] [--repeat R] [--no-hash] [--no-planes] [--sum-planes] [--pin FIRSTCORE] [--overlap-verdict]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#include <pthread.h>
#include <sys/resource.h>
#include "libHMDecoder_api.h"
#include "annexb.h"

struct Stream { std::vector<uint8_t> bytes; std::vector<std::pair<size_t, size_t> > nals; };
struct Shared
{
  const std::vector<Stream>* streams;
  bool hash, planes, sumPlanes, syncVerdict;
  int repeat;
  std::atomic<int> ready, failures;
  std::atomic<long> pictures;
  std::atomic<unsigned long long> planeSum;
  std::atomic<bool> go;
};

// Pushes the whole stream through a new decoder; returns the decoder (verdict still to be collected) or NULL with *rc set.
static libHMDec_context* decodePass(Shared& sh, int which, long& pictures, uint64_t& sink, int* rc)
{
  *rc = 0;
  libHMDec_context* dec = libHMDec_new_decoder();
  if (!dec) { *rc = 3; return NULL; }
  libHMDec_set_SEI_Check(dec, sh.hash);
  const std::vector<uint8_t>& s = (*sh.streams)[which].bytes;
  const std::vector<std::pair<size_t, size_t> >& nals = (*sh.streams)[which].nals;
  for (size_t k = 0; k < nals.size();)
  {
    bool newPicture = false, checkOutput = false;
    if (libHMDec_push_nal_unit(dec, &s[nals[k].first], (int)nals[k].second, k + 1 == nals.size(), newPicture, checkOutput) != LIBHMDEC_OK) { *rc = 4; libHMDec_free_decoder(dec); return NULL; }
    if (checkOutput)
      while (libHMDec_picture* pic = libHMDec_get_picture(dec))
      {
        pictures++;
        if (sh.planes)
          for (int c = 0; c < 3; c++)
          {
            const short* p = libHMDEC_get_image_plane(pic, (libHMDec_ColorComponent)c);
            if (!p) continue;
            if (!sh.sumPlanes) { sink += (uint64_t)p[0]; continue; }
            const int w = libHMDEC_get_picture_width(pic, (libHMDec_ColorComponent)c), h = libHMDEC_get_picture_height(pic, (libHMDec_ColorComponent)c);
            const int st = libHMDEC_get_picture_stride(pic, (libHMDec_ColorComponent)c);
            for (int y = 0; y < h; y++, p += st)
            {
              // two accumulators of two 32-bit lanes each: the even and the odd 16-bit samples of every 64-bit word (no carries between
              // lanes for rows of up to 2^20 samples of 12 bits)
              uint64_t even = 0, odd = 0;
              int x = 0;
              for (; x + 4 <= w; x += 4) { uint64_t v; memcpy(&v, p + x, 8); v &= 0x0fff0fff0fff0fffull; even += v & 0x0000ffff0000ffffull; odd += (v >> 16) & 0x0000ffff0000ffffull; }
              sink += (even & 0xffffffffull) + (even >> 32) + (odd & 0xffffffffull) + (odd >> 32);
              for (; x < w; x++) sink += (uint64_t)(p[x] & 0x0fff);
            }
          }
      }
    if (!newPicture) k++;
  }
  return dec;
}

// Verdict of a finished pass (waits for the hash checks still in flight), then frees the decoder.
static int collect(libHMDec_context* dec)
{
  if (!dec) return 0;
  const int rc = (libHMDecB200_hash_mismatch(dec) || libHMDecB200_unsupported(dec)) ? 1 : 0;
  libHMDec_free_decoder(dec);
  return rc;
}

static void worker(Shared* sh, int core, int which)
{
  if (core >= 0)
  {
    cpu_set_t set; CPU_ZERO(&set); CPU_SET(core, &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
  }
  long pics = 0; uint64_t sink = 0;
  int rc = 0;
  // warm-up passes outside the timed region: CUDA context, first allocations (two decoders' worth of buffers), page cache
  libHMDec_context* prev = decodePass(*sh, which, pics, sink, &rc);
  if (rc) sh->failures++;
  if (!sh->syncVerdict)
  {
    libHMDec_context* second = decodePass(*sh, which, pics, sink, &rc);
    if (rc) sh->failures++;
    if (collect(second)) sh->failures++;
  }
  if (collect(prev)) sh->failures++;
  prev = NULL;
  pics = 0; sink = 0;
  sh->ready++;
  while (!sh->go.load()) std::this_thread::yield();
  for (int r = 0; r < sh->repeat; r++)
  {
    libHMDec_context* dec = decodePass(*sh, which, pics, sink, &rc);
    if (rc) sh->failures++;
    if (sh->syncVerdict) { if (collect(dec)) sh->failures++; }
    else { if (collect(prev)) sh->failures++; prev = dec; }
  }
  if (collect(prev)) sh->failures++;
  sh->pictures += pics;
  sh->planeSum += sink;
}

int main(int argc, char** argv)
{
  std::vector<const char*> ins; int threads = 1, repeat = 1, pin = -1; bool hash = true, planes = true, sumPlanes = false, syncVerdict = true; double startAt = 0;
  for (int i = 1; i < argc; i++)
  {
    if (!strcmp(argv[i], "-b") && i + 1 < argc) ins.push_back(argv[++i]);
    else if (!strcmp(argv[i], "--sum-planes")) sumPlanes = true;
    else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--pin") && i + 1 < argc) pin = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--start-at") && i + 1 < argc) startAt = atof(argv[++i]);   // unix time: common start of several processes
    else if (!strcmp(argv[i], "--no-hash")) hash = false;
    else if (!strcmp(argv[i], "--no-planes")) planes = false;
    else if (!strcmp(argv[i], "--overlap-verdict")) syncVerdict = false;
    else { fprintf(stderr, "usage: %s -b in.bin [--threads T] [--repeat R] [--no-hash] [--no-planes] [--pin FIRSTCORE] [--overlap-verdict]\n", argv[0]); return 2; }
  }
  if (ins.empty()) return 2;
  setenv("HMDEC_B200_QUIET", "1", 0);
  std::vector<Stream> streams(ins.size());
  for (size_t i = 0; i < ins.size(); i++)
  {
    if (!readFile(ins[i], streams[i].bytes)) { perror(ins[i]); return 2; }
    splitAnnexB(streams[i].bytes, streams[i].nals);
  }
  Shared sh;
  sh.streams = &streams; sh.hash = hash; sh.planes = planes; sh.sumPlanes = sumPlanes; sh.syncVerdict = syncVerdict; sh.repeat = repeat;
  sh.ready = 0; sh.failures = 0; sh.pictures = 0; sh.planeSum = 0; sh.go = false;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(worker, &sh, pin >= 0 ? pin + t : -1, (int)(t % streams.size()));
  while (sh.ready.load() < threads) std::this_thread::sleep_for(std::chrono::milliseconds(1));
  if (startAt > 0)
    while (std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count() < startAt) std::this_thread::sleep_for(std::chrono::microseconds(200));
  const double wallStart = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
  struct rusage ru0, ru1;
  getrusage(RUSAGE_SELF, &ru0);
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  sh.go = true;
  for (size_t t = 0; t < pool.size(); t++) pool[t].join();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  getrusage(RUSAGE_SELF, &ru1);
  const double user = (ru1.ru_utime.tv_sec - ru0.ru_utime.tv_sec) + 1e-6 * (ru1.ru_utime.tv_usec - ru0.ru_utime.tv_usec);
  const double sys = (ru1.ru_stime.tv_sec - ru0.ru_stime.tv_sec) + 1e-6 * (ru1.ru_stime.tv_usec - ru0.ru_stime.tv_usec);
  printf("{\"threads\": %d, \"repeat\": %d, \"pictures\": %ld, \"seconds\": %.6f, \"fps\": %.3f, \"hash\": %s, \"planes\": \"%s\", \"plane_sum\": %llu, \"streams\": %d, \"failures\": %d, \"verdict\": \"%s\", \"t_start\": %.6f, \"t_end\": %.6f, "
         "\"cpu_user_s\": %.3f, \"cpu_sys_s\": %.3f, \"minor_faults\": %ld, \"ctx_switches_invol\": %ld}\n",
         threads, repeat, sh.pictures.load(), sec, sh.pictures.load() / sec, hash ? "true" : "false", !planes ? "untouched" : (sumPlanes ? "every sample read" : "first sample read"), sh.planeSum.load(), (int)streams.size(), sh.failures.load(), syncVerdict ? "sync" : "overlapped", wallStart, wallStart + sec,
         user, sys, (long)(ru1.ru_minflt - ru0.ru_minflt), (long)(ru1.ru_nivcsw - ru0.ru_nivcsw));
  fflush(stdout);
  return sh.failures.load() ? 1 : 0;
}
