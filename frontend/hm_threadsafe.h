// hm_threadsafe.h — process-wide guard for the HM globals that depend on the active SPS (see hm_threadsafe.cpp).
#ifndef HM_THREADSAFE_H
#define HM_THREADSAFE_H

class TComSPS;

// What TDecTop::xActivateParameterSets (TDecTop.cpp:323-333) and TDecCu::create (TDecCu.cpp:95-100) write into globals.
struct HmGeomKey
{
  int bitDepth[2], maxTrDynamicRange[2];
  unsigned maxCUWidth, maxCUHeight, maxCUDepth, addCUDepth;
  bool valid;
  HmGeomKey() : maxCUWidth(0), maxCUHeight(0), maxCUDepth(0), addCUDepth(0), valid(false) { bitDepth[0] = bitDepth[1] = maxTrDynamicRange[0] = maxTrDynamicRange[1] = 0; }
  bool operator==(const HmGeomKey& o) const
  {
    return valid && o.valid && bitDepth[0] == o.bitDepth[0] && bitDepth[1] == o.bitDepth[1] && maxTrDynamicRange[0] == o.maxTrDynamicRange[0] &&
           maxTrDynamicRange[1] == o.maxTrDynamicRange[1] && maxCUWidth == o.maxCUWidth && maxCUHeight == o.maxCUHeight &&
           maxCUDepth == o.maxCUDepth && addCUDepth == o.addCUDepth;
  }
};

HmGeomKey hm_geom_key_of(TComSPS* sps);

// Calls into HM are grouped by key: any number of calls with the SAME key run concurrently, calls with different keys
// exclude each other, and the globals are rewritten for the key that takes over.  An invalid key (the SPS the call will
// activate is not known) runs alone.  hm_geom_leave's argument is the key HM has active when the call ends.
void hm_geom_enter(const HmGeomKey& key);
void hm_geom_leave(const HmGeomKey& activeNow);
// statistics for tests: how often the globals were re-bound to another key
unsigned long hm_geom_rebinds();

#endif
