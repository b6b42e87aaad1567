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
// Motion-field compression of a picture nobody will reference, postponed until somebody asks for its motion data.
void hm_fast_defer_motion_compression(TComPic* pic);
void hm_fast_ensure_motion_compressed(TComPic* pic);
void hm_fast_prefetch_begin(TComPic* pic, unsigned ctuAddr);   // warm the per-partition arrays of a CTU ahead of initCU:
void hm_fast_prefetch_step(int nLines);                        // ... a few cache lines at a time
#endif
