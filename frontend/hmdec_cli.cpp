// hmdec_cli.cpp — Annex-B harness around the libHMDec_* entry points (the loop documented in the
// reference header, libHMDecoder.h:38-77): split the byte stream into NAL units, push them one at a
// time, re-push when bNewPicture comes back, drain pictures when checkOutputPictures is set.
//   hmdec_cli -b in.bin [-o out.yuv [--packed [-d bits]]] [--dump records.hmr] [--no-hash] [--touch-planes]
// -o writes the FULL coded picture (the wrapper API exposes no conformance window), 1 byte/sample for
// 8-bit streams and 2 bytes little-endian otherwise.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include "libHMDecoder_api.h"

#include "annexb.h"

static void writePicture(FILE* f, libHMDec_picture* pic, int bitDepth[2])
{
  for (int c = 0; c < 3; c++)
  {
    libHMDec_ColorComponent cc = (libHMDec_ColorComponent)c;
    int w = libHMDEC_get_picture_width(pic, cc), h = libHMDEC_get_picture_height(pic, cc), s = libHMDEC_get_picture_stride(pic, cc);
    const short* p = libHMDEC_get_image_plane(pic, cc);
    if (!p || w <= 0) continue;
    const bool is8 = bitDepth[c ? 1 : 0] <= 8;
    std::vector<uint8_t> row((size_t)w * 2);
    for (int y = 0; y < h; y++, p += s)
    {
      if (is8) { for (int x = 0; x < w; x++) row[x] = (uint8_t)p[x]; fwrite(row.data(), 1, w, f); }
      else     { for (int x = 0; x < w; x++) { row[2 * x] = (uint8_t)(p[x] & 0xff); row[2 * x + 1] = (uint8_t)((p[x] >> 8) & 0xff); } fwrite(row.data(), 1, (size_t)w * 2, f); }
    }
  }
}

int main(int argc, char** argv)
{
  const char* in = NULL; const char* out = NULL; const char* dump = NULL;
  bool hash = true, touch = false, packed = false; int repeat = 1, outDepth = 0;
  std::vector<uint8_t> packBuf;
  for (int i = 1; i < argc; i++)
  {
    if (!strcmp(argv[i], "-b") && i + 1 < argc) in = argv[++i];
    else if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
    else if (!strcmp(argv[i], "--dump") && i + 1 < argc) dump = argv[++i];
    else if (!strcmp(argv[i], "--no-hash")) hash = false;
    else if (!strcmp(argv[i], "--packed")) packed = true;                       // -o in TAppDecoder's format (cropped / converted on the GPU)
    else if (!strcmp(argv[i], "-d") && i + 1 < argc) outDepth = atoi(argv[++i]);   // output bit depth for --packed (0 = internal)
    else if (!strcmp(argv[i], "--touch-planes")) touch = true;
    else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = atoi(argv[++i]);
    else { fprintf(stderr, "usage: %s -b in.bin [-o out.yuv] [--dump file] [--no-hash] [--touch-planes] [--repeat N]\n", argv[0]); return 2; }
  }
  if (!in) { fprintf(stderr, "missing -b\n"); return 2; }
  std::vector<uint8_t> stream;
  if (!readFile(in, stream)) { perror(in); return 2; }
  std::vector<std::pair<size_t, size_t> > nals;
  splitAnnexB(stream, nals);

  FILE* fo = out ? fopen(out, "wb") : NULL;
  int pictures = 0; bool mismatch = false; const char* unsupported = NULL;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  const bool picTimes = getenv("HMDEC_CLI_PICTURE_TIMES") != NULL;     // host time between picture boundaries (profiling aid)
  double tLast = 0;
  for (int rep = 0; rep < repeat; rep++)
  {
    libHMDec_context* dec = dump ? libHMDecB200_new_decoder_ex(1, dump) : libHMDec_new_decoder();
    if (!dec) { fprintf(stderr, "could not create decoder\n"); return 3; }
    libHMDec_set_SEI_Check(dec, hash);
    for (size_t k = 0; k < nals.size();)
    {
      const bool eof = (k + 1 == nals.size());
      bool newPicture = false, checkOutput = false;
      if (libHMDec_push_nal_unit(dec, &stream[nals[k].first], (int)nals[k].second, eof, newPicture, checkOutput) != LIBHMDEC_OK)
      {
        fprintf(stderr, "push_nal_unit failed at NAL %zu\n", k); return 4;
      }
      if (checkOutput)
      {
        while (libHMDec_picture* pic = libHMDec_get_picture(dec))
        {
          pictures++;
          int bd[2] = { libHMDEC_get_internal_bit_depth(LIBHMDEC_LUMA), libHMDEC_get_internal_bit_depth(LIBHMDEC_CHROMA_U) };
          if (fo && packed)
          {
            // TAppDecoder -o equivalent: cropped, bit-depth converted and packed on the GPU
            long n = libHMDecB200_pack_picture(dec, pic, outDepth, outDepth, NULL, 0);
            if (n < 0) { fprintf(stderr, "pack_picture failed\n"); return 5; }
            if ((size_t)n > packBuf.size()) packBuf.resize((size_t)n);
            if (libHMDecB200_pack_picture(dec, pic, outDepth, outDepth, packBuf.data(), packBuf.size()) != n) { fprintf(stderr, "pack_picture failed\n"); return 5; }
            fwrite(packBuf.data(), 1, (size_t)n, fo);
          }
          else if (fo) writePicture(fo, pic, bd);
          else if (touch) for (int c = 0; c < 3; c++) (void)libHMDEC_get_image_plane(pic, (libHMDec_ColorComponent)c);
        }
      }
      if (newPicture && picTimes) { const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); fprintf(stderr, "picture boundary at NAL %zu: +%.2f ms\n", k, 1e3 * (t - tLast)); tLast = t; }
      if (!newPicture) k++;      // otherwise the same NAL must be pushed again
    }
    mismatch = mismatch || libHMDecB200_hash_mismatch(dec);
    if (libHMDecB200_unsupported(dec)) unsupported = libHMDecB200_unsupported(dec);
    libHMDec_free_decoder(dec);
  }
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (fo) fclose(fo);
  fprintf(stderr, "hmdec_cli: %d pictures in %.3f s (%.2f fps)%s%s%s\n", pictures, sec, pictures / sec,
          mismatch ? "  HASH MISMATCH" : "", unsupported ? "  UNSUPPORTED: " : "", unsupported ? unsupported : "");
  return (mismatch || unsupported) ? 1 : 0;
}
