#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <ucontext.h>
#include <unistd.h>
#include <sys/syscall.h>
#define MAXS (1<<22)
static unsigned long *pcs, *rets; static int* tids; static volatile long n;
static void h(int sig, siginfo_t* si, void* uc_) { ucontext_t* uc = (ucontext_t*)uc_; long i = __sync_fetch_and_add(&n, 1); if (i < MAXS) { pcs[i] = uc->uc_mcontext.gregs[REG_RIP]; rets[i] = *(unsigned long*)uc->uc_mcontext.gregs[REG_RSP]; tids[i] = (int)syscall(SYS_gettid); } }
__attribute__((constructor)) static void init(void) {
  pcs = malloc(sizeof(unsigned long) * MAXS); rets = malloc(sizeof(unsigned long) * MAXS); tids = malloc(sizeof(int) * MAXS);
  struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_sigaction = h; sa.sa_flags = SA_SIGINFO | SA_RESTART; sigaction(SIGPROF, &sa, 0);
  struct itimerval it = { {0, 1000}, {0, 1000} }; setitimer(ITIMER_PROF, &it, 0);
}
__attribute__((destructor)) static void fini(void) {
  struct itimerval it = { {0, 0}, {0, 0} }; setitimer(ITIMER_PROF, &it, 0);
  const char* out = getenv("PCS_OUT"); if (!out) out = "/tmp/pcsample.txt";
  FILE* f = fopen(out, "w"); FILE* m = fopen("/proc/self/maps", "r"); char line[512];
  while (fgets(line, sizeof line, m)) if (strstr(line, " r-xp ") || strstr(line, "r-xp")) fprintf(f, "M %s", line);
  long c = n < MAXS ? n : MAXS; for (long i = 0; i < c; i++) fprintf(f, "S %lx %d %lx\n", pcs[i], tids[i], rets[i]); fclose(f);
}
