// hmdec_internals.cpp — dumps what ANY implementation of the libHMDecoder wrapper ABI reports for a bitstream: per output
// picture the POC, geometry, a checksum of the planes and, for every libHMDec_info_type, the complete block list of
// libHMDEC_get_internal_info.  The library is dlopen'ed, so the same binary drives the reference wrapper
// (oracle/_ref/liblibHMDecoderStatic.so) and this repository's drop-in; tests compare the two dumps byte for byte.
//   hmdec_internals <library.so> <in.bin> <out.txt> [--backend N ARG]     (--backend: libHMDecB200_new_decoder_ex, drop-in only)
// HMDEC_INTERNALS_MAX_TLAYER=N: libHMDec_set_max_temporal_layer(N) before the first NAL (both libraries).
// HMDEC_INTERNALS_TIME=1: do not print the block lists, report the time spent inside libHMDEC_get_internal_info instead.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <dlfcn.h>
#include <chrono>
#include "libHMDecoder_api.h"
#include "annexb.h"

#define RESOLVE(name) decltype(&::name) p_##name = (decltype(&::name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

int main(int argc, char** argv)
{
  if (argc < 4) { fprintf(stderr, "usage: %s <library.so> <in.bin> <out.txt> [--backend N ARG]\n", argv[0]); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  RESOLVE(libHMDec_new_decoder) RESOLVE(libHMDec_free_decoder) RESOLVE(libHMDec_set_SEI_Check) RESOLVE(libHMDec_push_nal_unit)
  RESOLVE(libHMDec_get_picture) RESOLVE(libHMDEC_get_POC) RESOLVE(libHMDEC_get_picture_width) RESOLVE(libHMDEC_get_picture_height)
  RESOLVE(libHMDEC_get_picture_stride) RESOLVE(libHMDEC_get_image_plane) RESOLVE(libHMDEC_get_chroma_format)
  RESOLVE(libHMDEC_get_internal_bit_depth) RESOLVE(libHMDEC_get_internal_info) RESOLVE(libHMDEC_clear_internal_info)
  libHMDec_context* dec = NULL;
  if (argc >= 7 && !strcmp(argv[4], "--backend"))
  {
    decltype(&::libHMDecB200_new_decoder_ex) ex = (decltype(&::libHMDecB200_new_decoder_ex))dlsym(lib, "libHMDecB200_new_decoder_ex");
    if (!ex) { fprintf(stderr, "library has no libHMDecB200_new_decoder_ex\n"); return 2; }
    dec = ex(atoi(argv[5]), argv[6]);
  }
  else dec = p_libHMDec_new_decoder();
  if (!dec) { fprintf(stderr, "no decoder\n"); return 3; }
  p_libHMDec_set_SEI_Check(dec, true);
  if (getenv("HMDEC_INTERNALS_MAX_TLAYER")) { RESOLVE(libHMDec_set_max_temporal_layer) p_libHMDec_set_max_temporal_layer(dec, atoi(getenv("HMDEC_INTERNALS_MAX_TLAYER"))); }
  std::vector<uint8_t> stream;
  if (!readFile(argv[2], stream)) { perror(argv[2]); return 2; }
  std::vector<std::pair<size_t, size_t> > nals;
  splitAnnexB(stream, nals);
  FILE* out = fopen(argv[3], "w");
  if (!out) { perror(argv[3]); return 2; }
  long pictures = 0, blocks = 0;
  double secInfo = 0, secType[24] = {0};
  const bool timeOnly = getenv("HMDEC_INTERNALS_TIME") != NULL;
  for (size_t k = 0; k < nals.size();)
  {
    bool newPicture = false, checkOutput = false;
    if (p_libHMDec_push_nal_unit(dec, &stream[nals[k].first], (int)nals[k].second, k + 1 == nals.size(), newPicture, checkOutput) != LIBHMDEC_OK) return 4;
    if (checkOutput)
      while (libHMDec_picture* pic = p_libHMDec_get_picture(dec))
      {
        pictures++;
        fprintf(out, "PIC poc %d chroma %d depth %d %d\n", p_libHMDEC_get_POC(pic), (int)p_libHMDEC_get_chroma_format(pic),
                p_libHMDEC_get_internal_bit_depth(LIBHMDEC_LUMA), p_libHMDEC_get_internal_bit_depth(LIBHMDEC_CHROMA_U));
        for (int c = 0; c < 3; c++)
        {
          const int w = p_libHMDEC_get_picture_width(pic, (libHMDec_ColorComponent)c), h = p_libHMDEC_get_picture_height(pic, (libHMDec_ColorComponent)c);
          const int stride = p_libHMDEC_get_picture_stride(pic, (libHMDec_ColorComponent)c);
          const short* p = p_libHMDEC_get_image_plane(pic, (libHMDec_ColorComponent)c);
          unsigned long long sum = 0;                                    // position-dependent checksum of the visible samples
          if (p) for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) sum = sum * 1000003ull + (unsigned short)p[(size_t)y * stride + x];
          fprintf(out, " plane %d %dx%d stride %d sum %016llx\n", c, w, h, stride, sum);
        }
        for (int type = LIBHMDEC_CTU_SLICE_INDEX; type <= LIBHMDEC_TU_COEFF_ENERGY_CR; type++)
        {
          const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
          std::vector<libHMDec_BlockValue>* v = p_libHMDEC_get_internal_info(dec, pic, (libHMDec_info_type)type);
          { const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); secInfo += dt; secType[type] += dt;
            if (timeOnly && type == 0 && getenv("HMDEC_INTERNALS_TIME")[0] == '2') printf("picture %ld: first query %.3f ms\n", pictures, 1e3 * dt); }
          if (v) blocks += (long)v->size();
          fprintf(out, " type %d n %ld\n", type, v ? (long)v->size() : -1L);
          if (v && !timeOnly) for (size_t i = 0; i < v->size(); i++)
          {
            const libHMDec_BlockValue& b = (*v)[i];
            // value2 is only defined for the motion-vector types (the reference leaves it uninitialised elsewhere)
            const bool mv = type == LIBHMDEC_PU_MV_0 || type == LIBHMDEC_PU_MV_1;
            fprintf(out, "  %u %u %u %u %d %d\n", b.x, b.y, b.w, b.h, b.value, mv ? b.value2 : 0);
          }
        }
        p_libHMDEC_clear_internal_info(dec);
      }
    if (!newPicture) k++;
  }
  fclose(out);
  p_libHMDec_free_decoder(dec);
  printf("%ld pictures, %ld blocks reported, %.3f s inside libHMDEC_get_internal_info\n", pictures, blocks, secInfo);
  if (timeOnly) { printf("ms per type:"); for (int t = 0; t < 24; t++) printf(" %d:%.1f", t, 1e3 * secType[t]); printf("\n"); }
  return 0;
}
