// Force-included (-include) into HM's TComDataCU.cpp by frontend/Makefile: every memset in that file becomes
// hm_fast_memset (hm_fast.cpp), which skips the per-CTU zero fill of the coefficient arrays when the record emitter
// keeps them clean.  The standard headers are pulled in first so that only the calls are renamed.
#include <cstring>
#include <string.h>
#include <stddef.h>
extern "C" void* hm_fast_memset(void* p, int v, size_t n);
#define memset hm_fast_memset
