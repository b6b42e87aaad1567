// LD_PRELOAD diagnostic: time spent inside malloc for requests >= 512 bytes, per call site (frame above operator new),
// written to $MS_OUT at exit.  Resolve with tools/diag/malloc_sites.py.  Not part of the product.
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <execinfo.h>
#include <x86intrin.h>
extern void* __libc_malloc(size_t);
#define NS 8192
static struct { void* site; unsigned long n, bytes, cycles, worst; } tab[NS];
static volatile int lock;
static __thread int inside;
void* malloc(size_t n)
{
  if (inside || n < 512) return __libc_malloc(n);
  inside = 1;
  void* bt[5]; int d = backtrace(bt, 5);
  void* ra = d > 2 ? bt[2] : (d > 1 ? bt[1] : 0);
  const unsigned long long t0 = __rdtsc();
  void* p = __libc_malloc(n);
  const unsigned long dt = (unsigned long)(__rdtsc() - t0);
  while (__sync_lock_test_and_set(&lock, 1)) ;
  unsigned h = ((uintptr_t)ra >> 2) % NS;
  for (int i = 0; i < NS; i++, h = (h + 1) % NS)
    if (tab[h].site == ra || !tab[h].site) { tab[h].site = ra; tab[h].n++; tab[h].bytes += n; tab[h].cycles += dt; if (dt > tab[h].worst) tab[h].worst = dt; break; }
  __sync_lock_release(&lock);
  inside = 0;
  return p;
}
__attribute__((destructor)) static void fin(void)
{
  FILE* f = fopen(getenv("MS_OUT") ? getenv("MS_OUT") : "/tmp/malloc_sites.txt", "w");
  FILE* m = fopen("/proc/self/maps", "r"); char line[512];
  while (fgets(line, sizeof line, m)) if (strstr(line, "r-xp")) fprintf(f, "M %s", line);
  for (int i = 0; i < NS; i++) if (tab[i].site) fprintf(f, "S %p %lu %lu %lu %lu\n", tab[i].site, tab[i].n, tab[i].bytes, tab[i].cycles, tab[i].worst);
  fclose(f);
}
