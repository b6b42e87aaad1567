// placeholder until the engine is linked (replaced below)
#include <cstdio>
#include "hm_emit.h"
HmFrameSink* hm_new_gpu_sink() { fprintf(stderr, "hmdec_b200: GPU engine not linked into this build\n"); return 0; }
