// libHMDecoder_api.h — declarations of the 16 C entry points this library exports.  They are
// ABI-identical to the reference wrapper (source/App/libHMDecoder/libHMDecoder.h:102-298), so a
// caller built against the reference header links/dlopens this library unchanged.  Written from
// the interface description in SURVEY.md §8(b); the mixed libHMDec_/libHMDEC_ casing is part of
// the ABI, as are the `bool&` out-parameters and the std::vector* return of get_internal_info.
#ifndef LIBHMDECODER_API_B200_H
#define LIBHMDECODER_API_B200_H

#include <vector>
#include <stdint.h>

extern "C" {

typedef enum { LIBHMDEC_OK = 0, LIBHMDEC_ERROR, LIBHMDEC_ERROR_READ_ERROR } libHMDec_error;
typedef void libHMDec_context;
typedef void libHMDec_picture;
typedef enum { LIBHMDEC_LUMA = 0, LIBHMDEC_CHROMA_U, LIBHMDEC_CHROMA_V } libHMDec_ColorComponent;
typedef enum { LIBHMDEC_CHROMA_400 = 0, LIBHMDEC_CHROMA_420, LIBHMDEC_CHROMA_422, LIBHMDEC_CHROMA_444, LIBHMDEC_CHROMA_UNKNOWN } libHMDec_ChromaFormat;

typedef struct { unsigned short x, y, w, h; int value; int value2; } libHMDec_BlockValue;

typedef enum
{
  LIBHMDEC_CTU_SLICE_INDEX = 0,
  LIBHMDEC_CU_PREDICTION_MODE, LIBHMDEC_CU_TRQ_BYPASS, LIBHMDEC_CU_SKIP_FLAG, LIBHMDEC_CU_PART_MODE,
  LIBHMDEC_CU_INTRA_MODE_LUMA, LIBHMDEC_CU_INTRA_MODE_CHROMA, LIBHMDEC_CU_ROOT_CBF,
  LIBHMDEC_PU_MERGE_FLAG, LIBHMDEC_PU_MERGE_INDEX, LIBHMDEC_PU_UNI_BI_PREDICTION,
  LIBHMDEC_PU_REFERENCE_POC_0, LIBHMDEC_PU_MV_0, LIBHMDEC_PU_REFERENCE_POC_1, LIBHMDEC_PU_MV_1,
  LIBHMDEC_TU_CBF_Y, LIBHMDEC_TU_CBF_CB, LIBHMDEC_TU_CBF_CR,
  LIBHMDEC_TU_COEFF_TR_SKIP_Y, LIBHMDEC_TU_COEFF_TR_SKIP_Cb, LIBHMDEC_TU_COEFF_TR_SKIP_Cr,
  LIBHMDEC_TU_COEFF_ENERGY_Y, LIBHMDEC_TU_COEFF_ENERGY_CB, LIBHMDEC_TU_COEFF_ENERGY_CR,
} libHMDec_info_type;

const char*        libHMDec_get_version(void);
libHMDec_context*  libHMDec_new_decoder(void);
libHMDec_error     libHMDec_free_decoder(libHMDec_context* decCtx);
void               libHMDec_set_SEI_Check(libHMDec_context* decCtx, bool check_hash);
void               libHMDec_set_max_temporal_layer(libHMDec_context* decCtx, int max_layer);
libHMDec_error     libHMDec_push_nal_unit(libHMDec_context* decCtx, const void* data8, int length, bool eof, bool& bNewPicture, bool& checkOutputPictures);
libHMDec_picture*  libHMDec_get_picture(libHMDec_context* decCtx);
int                libHMDEC_get_POC(libHMDec_picture* pic);
int                libHMDEC_get_picture_width(libHMDec_picture* pic, libHMDec_ColorComponent c);
int                libHMDEC_get_picture_height(libHMDec_picture* pic, libHMDec_ColorComponent c);
int                libHMDEC_get_picture_stride(libHMDec_picture* pic, libHMDec_ColorComponent c);
short*             libHMDEC_get_image_plane(libHMDec_picture* pic, libHMDec_ColorComponent c);
libHMDec_ChromaFormat libHMDEC_get_chroma_format(libHMDec_picture* pic);
int                libHMDEC_get_internal_bit_depth(libHMDec_ColorComponent c);
std::vector<libHMDec_BlockValue>* libHMDEC_get_internal_info(libHMDec_context* decCtx, libHMDec_picture* pic, libHMDec_info_type type);
libHMDec_error     libHMDEC_clear_internal_info(libHMDec_context* decCtx);

// ---- extensions of this implementation (not in the reference) ----
// Reconstruction back-ends: 0 = GPU engine (libhmrecon.so, default; fails loudly when CUDA is absent),
// 1 = record dump to the file named by `arg` with HM's CPU reconstruction as golden (tools / tests only).
libHMDec_context*  libHMDecB200_new_decoder_ex(int backend, const char* arg);
// true when any SEI decoded-picture-hash check failed so far (the reference only prints it, libHMDecoder.cpp:59,165)
bool               libHMDecB200_hash_mismatch(libHMDec_context* decCtx);
// The picture as TAppDecoder's `-o` would write it (conformance-window crop, output bit depth 0 = internal, planar Y/Cb/Cr,
// 1 byte per sample when both depths <= 8 else 2 bytes LE), packed on the device: only these bytes cross PCIe.
// Returns the number of bytes (dst == NULL: query only) or -1.
long               libHMDecB200_pack_picture(libHMDec_context* decCtx, libHMDec_picture* pic, int outBitDepthLuma, int outBitDepthChroma, void* dst, size_t capacity);
// name of the first bitstream feature the GPU path does not implement, or NULL
const char*        libHMDecB200_unsupported(libHMDec_context* decCtx);

}
#endif
