// LD_PRELOAD interposer: counts mmap / munmap / madvise / mremap calls (what glibc's malloc does for large blocks and when it trims)
// by size class and by the first two callers outside libc (backtrace()), steady state only if MMT_AFTER_S is set (seconds after
// start).  Output MMT_OUT (default /tmp/mmap_trace.txt); resolve with tools/mmap_trace_resolve.py.
#define _GNU_SOURCE
#include <dlfcn.h>
#include <execinfo.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
static void* (*real_mmap)(void*, size_t, int, int, int, off_t);
static int (*real_munmap)(void*, size_t);
static int (*real_madvise)(void*, size_t, int);
static __thread int inside;
#define NB 4096
static struct { int kind; void* pc[4]; size_t n, bytes; } tab[NB];
static double t0; static double after = -1;
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static void note(int kind, size_t bytes)
{
  if (inside) return;
  if (after < 0) { t0 = now(); after = getenv("MMT_AFTER_S") ? atof(getenv("MMT_AFTER_S")) : 0; }
  if (now() - t0 < after) return;
  inside = 1;
  void* bt[8]; int n = backtrace(bt, 8);
  uintptr_t h = kind; for (int i = 2; i < n && i < 6; i++) h = h * 1000003u + (uintptr_t)bt[i];
  size_t k = h % NB;
  for (int t = 0; t < 32; t++, k = (k + 1) % NB)
  {
    if (tab[k].n == 0) { tab[k].kind = kind; for (int i = 0; i < 4; i++) tab[k].pc[i] = i + 2 < n ? bt[i + 2] : 0; }
    if (tab[k].kind == kind && tab[k].pc[0] == (2 < n ? bt[2] : 0) && tab[k].pc[1] == (3 < n ? bt[3] : 0) && tab[k].pc[2] == (4 < n ? bt[4] : 0) && tab[k].pc[3] == (5 < n ? bt[5] : 0))
    { __sync_fetch_and_add(&tab[k].n, 1); __sync_fetch_and_add(&tab[k].bytes, bytes); break; }
  }
  inside = 0;
}
void* mmap(void* a, size_t len, int prot, int flags, int fd, off_t off)
{
  if (!real_mmap) real_mmap = dlsym(RTLD_NEXT, "mmap");
  void* r = real_mmap(a, len, prot, flags, fd, off);
  note(0, len);
  return r;
}
int munmap(void* a, size_t len) { if (!real_munmap) real_munmap = dlsym(RTLD_NEXT, "munmap"); note(1, len); return real_munmap(a, len); }
int madvise(void* a, size_t len, int adv) { if (!real_madvise) real_madvise = dlsym(RTLD_NEXT, "madvise"); note(2, len); return real_madvise(a, len, adv); }
__attribute__((destructor)) static void fini(void)
{
  FILE* f = fopen(getenv("MMT_OUT") ? getenv("MMT_OUT") : "/tmp/mmap_trace.txt", "w");
  FILE* m = fopen("/proc/self/maps", "r"); char line[512];
  while (fgets(line, sizeof line, m)) if (strstr(line, "r-xp")) fprintf(f, "M %s", line);
  for (int k = 0; k < NB; k++) if (tab[k].n) fprintf(f, "C %d %zu %zu %p %p %p %p\n", tab[k].kind, tab[k].n, tab[k].bytes, tab[k].pc[0], tab[k].pc[1], tab[k].pc[2], tab[k].pc[3]);
  fclose(f);
}
