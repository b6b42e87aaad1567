// hm_shim.cpp — compiled with -DdecompressCU=decompressCU_hm (exactly like the copy of TDecCu.cpp that
// frontend/Makefile builds), so that these calls reach HM's ORIGINAL CPU reconstruction, which
// otherwise is replaced by the hook in hm_hooks.cpp.  Used only in verification mode
// (HmFrameSink::wantHmRecon()).
#include "TLibDecoder/TDecCu.h"
void hm_call_original_decompressCU(TDecCu* dec, TComDataCU* ctu) { dec->decompressCU(ctu); }
