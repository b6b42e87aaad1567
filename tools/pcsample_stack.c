// Like pcsample.c, but keeps the top 192 stack words of every sample so that the resolver can find the first return address
// outside libc (who called malloc / free / memset when the PC is inside a static libc routine).  1 kHz SIGPROF, <= 64 K samples.
#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <ucontext.h>
#define MAXS (1<<16)
#define NW 192
static unsigned long *pcs, *stk; static volatile long n;
static void h(int sig, siginfo_t* si, void* uc_) { ucontext_t* uc = (ucontext_t*)uc_; long i = __sync_fetch_and_add(&n, 1); if (i < MAXS) { pcs[i] = uc->uc_mcontext.gregs[REG_RIP]; memcpy(stk + i * NW, (void*)uc->uc_mcontext.gregs[REG_RSP], NW * 8); } }
__attribute__((constructor)) static void init(void) {
  pcs = malloc(sizeof(unsigned long) * MAXS); stk = malloc(sizeof(unsigned long) * MAXS * NW);
  struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_sigaction = h; sa.sa_flags = SA_SIGINFO | SA_RESTART; sigaction(SIGPROF, &sa, 0);
  struct itimerval it = { {0, 1000}, {0, 1000} }; setitimer(ITIMER_PROF, &it, 0);
}
__attribute__((destructor)) static void fini(void) {
  struct itimerval it = { {0, 0}, {0, 0} }; setitimer(ITIMER_PROF, &it, 0);
  const char* out = getenv("PCS_OUT"); if (!out) out = "/tmp/pcsample_stack.txt";
  FILE* f = fopen(out, "w"); FILE* m = fopen("/proc/self/maps", "r"); char line[512];
  while (fgets(line, sizeof line, m)) if (strstr(line, "r-xp")) fprintf(f, "M %s", line);
  long c = n < MAXS ? n : MAXS;
  for (long i = 0; i < c; i++) { fprintf(f, "S %lx", pcs[i]); for (int k = 0; k < NW; k++) fprintf(f, " %lx", stk[i * NW + k]); fprintf(f, "\n"); }
  fclose(f);
}
