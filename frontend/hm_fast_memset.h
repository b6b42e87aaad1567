// Force-included (-include) into HM's TComDataCU.cpp by frontend/Makefile: every memset in that file becomes
// hm_fast_memset (hm_fast.cpp), which skips the per-CTU zero fill of the coefficient arrays when the record emitter
// keeps them clean.  The standard headers are pulled in first so that only the calls are renamed.
//
// hm_fast_col_part: TMVP reads the collocated picture's motion at 16x16 granularity.  HM materialises that by rewriting
// the whole motion field of every finished picture (TComPic::compressMotion -> TComCUMvField::compress,
// TComMotionInfo.cpp:330-350: each run of N = (16/unitSize)^2 partitions in z-order takes the values of its first one).
// The only reader of another picture's motion, TComDataCU::xGetColMVP (TComDataCU.cpp:3381), gets the same values by
// rounding its partition index down to the run start instead — frontend/Makefile patches that one line — so the
// decoder never runs the rewrite (libHMDEC_get_internal_info still does, on demand, to report what the reference reports).
#include <cstring>
#include <string.h>
#include <stddef.h>
extern "C" void* hm_fast_memset(void* p, int v, size_t n);
#define memset hm_fast_memset
static inline unsigned hm_fast_col_part(unsigned part, int unitSize)
{
  const int scale = 16 / unitSize;                       // 4 * AMVP_DECIMATION_FACTOR / m_unitSize (TComDataCU.cpp:3515)
  return scale > 0 ? part & ~(unsigned)(scale * scale - 1) : part;
}
