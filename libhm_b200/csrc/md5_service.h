// md5_service.h — process-wide asynchronous MD5 service (md5_service.cu) and its kernel interface (k_md5.cu)
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

#define MD5_CHUNK 4096                 // 64-byte blocks per tick (~2 ms of one warp)

struct Md5Job
{
  const int16_t* plane[3];
  int pitch[3], w[3], h[3], bd[3];
  int ncomp;
  uint32_t* out;                       // 3 x 4 words (digest A,B,C,D per component = digest byte order); page-locked host memory
  uint32_t* state;                     // 3 x 4 words of device memory: chaining value between ticks
};
struct Md5TickJob { Md5Job J; int chunk; int pad; };

void launch_md5_tick(const Md5TickJob* jobs, int n, cudaStream_t s);
int  md5_chunks(const Md5Job& J);

// Hand a job to the service of `device`.  `ready` must have been recorded (on any stream) after the planes were
// written; *done is set to 1 once the digest is in J.out.  Returns false if the service cannot run.
bool md5_service_submit(int device, const Md5Job& J, cudaEvent_t ready, std::atomic<int>* done);
