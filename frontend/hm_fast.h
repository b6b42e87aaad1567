// hm_fast.h — switches of the allocation-/memset-free picture turnover (hm_fast.cpp)
#ifndef HM_FAST_H
#define HM_FAST_H
#include <stddef.h>
class TDecTop;
class TComPic;
// Per thread (= per decoder call): skip HM's whole-CTU coefficient zero fills (parseCoeffNxN zeroes what it parses).
void hm_fast_set_skip_coeff_fill(bool on);
// Process-wide: where the sample planes of HM's picture buffers come from (NULL = malloc, HM's default).
typedef void* (*HmPlaneAlloc)(size_t bytes);
typedef void  (*HmPlaneFree)(void* p);
void hm_fast_set_plane_allocator(HmPlaneAlloc a, HmPlaneFree f);
bool hm_fast_plane_is_pinned(const void* planeBuffer);
// Must run before TDecTop::destroy: hands allocator-owned planes back and frees picture buffers a flush dropped from the DPB list.
void hm_fast_release_decoder(TDecTop* dec);
// warm the per-partition arrays of a CTU ahead of TComDataCU::initCU, a few cache lines at a time (hm_fast.cpp)
struct HmPrefetchCursor
{
  const char* base[48]; unsigned lines[48]; int count, cur; unsigned line;
  HmPrefetchCursor() : count(0), cur(0), line(0) {}
  inline void step(int nLines)
  {
    while (nLines > 0 && cur < count)
    {
      __builtin_prefetch(base[cur] + ((unsigned long)line << 6), 1, 2);
      nLines--;
      if (++line >= lines[cur]) { cur++; line = 0; }
    }
  }
};
void hm_fast_prefetch_begin(HmPrefetchCursor& c, TComPic* pic, unsigned ctuAddr);
#endif
