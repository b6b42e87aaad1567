// hm_fast.h — switches of the allocation-/memset-free picture turnover (hm_fast.cpp)
#ifndef HM_FAST_H
#define HM_FAST_H
// Per thread (= per decoder call): the emitter clears coded coefficient blocks after use, so HM's per-CTU zero fills are skipped.
void hm_fast_set_clean_coeffs(bool on);
bool hm_fast_clean_coeffs();
#endif
