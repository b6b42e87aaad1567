#define _GNU_SOURCE
#include <dlfcn.h>
#include <execinfo.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static void* (*real_malloc)(size_t);
static __thread int inside;
#define NB 4096
static struct { void* pc[3]; size_t n, bytes; unsigned long long ticks; } tab[NB];
static inline unsigned long long rdtsc(void){ unsigned lo,hi; __asm__ volatile("rdtsc":"=a"(lo),"=d"(hi)); return ((unsigned long long)hi<<32)|lo; }
static char boot[1<<16]; static size_t bootp;
void* malloc(size_t sz)
{
  if (!real_malloc) { if (inside) { void* p = boot + bootp; bootp += (sz + 15) & ~15; return p; } inside = 1; real_malloc = dlsym(RTLD_NEXT, "malloc"); inside = 0; }
  unsigned long long t0 = rdtsc();
  void* r = real_malloc(sz);
  unsigned long long dt = rdtsc() - t0;
  if (sz >= 1024 && !inside)
  {
    inside = 1;
    void* bt[5]; int n = backtrace(bt, 5);
    uintptr_t h = 0; for (int i = 1; i < n && i < 4; i++) h = h * 1000003u + (uintptr_t)bt[i];
    size_t k = h % NB;
    for (int t = 0; t < 16; t++, k = (k + 1) % NB)
    {
      if (tab[k].n == 0) { for (int i = 0; i < 3; i++) tab[k].pc[i] = i + 1 < n ? bt[i + 1] : 0; }
      if (tab[k].pc[0] == (1 < n ? bt[1] : 0) && tab[k].pc[1] == (2 < n ? bt[2] : 0) && tab[k].pc[2] == (3 < n ? bt[3] : 0)) { __sync_fetch_and_add(&tab[k].n, 1); __sync_fetch_and_add(&tab[k].bytes, sz); __sync_fetch_and_add(&tab[k].ticks, dt); break; }
    }
    inside = 0;
  }
  return r;
}
__attribute__((destructor)) static void fini(void)
{
  FILE* f = fopen(getenv("MTRACE_OUT") ? getenv("MTRACE_OUT") : "/tmp/mtrace.txt", "w");
  FILE* m = fopen("/proc/self/maps", "r"); char line[512];
  while (fgets(line, sizeof line, m)) if (strstr(line, "r-xp")) fprintf(f, "M %s", line);
  for (int k = 0; k < NB; k++) if (tab[k].n) fprintf(f, "C %zu %zu %p %p %p %llu\n", tab[k].n, tab[k].bytes, tab[k].pc[0], tab[k].pc[1], tab[k].pc[2], tab[k].ticks);
  fclose(f);
}
