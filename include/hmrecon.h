/* hmrecon.h — C ABI of the B200 HEVC reconstruction engine (libhmrecon.so).
 *
 * This is the drop-in boundary for the reference's reconstruction hot path.  Each entry point names the
 * reference interface it replaces (paths under /root/reference/source):
 *
 *   hmr_submit_frame     replaces, for one whole picture, the per-CTU  TDecCu::decompressCU
 *                        (Lib/TLibDecoder/TDecCu.cpp:142, called from TDecSlice.cpp:334) and the per-picture
 *                        TComLoopFilter::loopFilterPic + TComSampleAdaptiveOffset::SAOProcess
 *                        (TDecGop::filterPicture, Lib/TLibDecoder/TDecGop.cpp:157-174)
 *   hmr_read_plane       replaces the host-memory plane access of TComPicYuv::getAddr as used by
 *                        libHMDEC_get_image_plane (App/libHMDecoder/libHMDecoder.cpp:402-417)
 *   hmr_read_packed      replaces TVideoIOYuv::write (Lib/TLibVideoIO/TVideoIOYuv.cpp:706-790): crop, bit-depth conversion,
 *                        8/16-bit packing of an output picture, on the device
 *   hmr_picture_hash     replaces calcChecksum / calcCRC (Lib/TLibCommon/TComPicYuvMD5.cpp:127-175)
 *   hmr_md5_submit/result replace calcMD5 (TComPicYuvMD5.cpp:183-205): one serial chain per plane, run asynchronously on
 *                        the device (many pictures in flight), so the SEI MD5 check costs the host nothing
 *
 * Plain C: pointers, sizes and the POD records of hmr_records.h; no C++/torch/CUDA types.  All functions return
 * HMR_OK (0) or a negative error code; hmr_error_string() describes the last error of an engine.
 * An engine owns one CUDA stream; calls on one engine must be serialised by the caller, different engines are
 * independent (one engine per bitstream).  There is no CPU fallback: without a usable CUDA device
 * hmr_engine_create fails with HMR_ERR_CUDA.
 */
#ifndef HMRECON_H
#define HMRECON_H

#include <stddef.h>
#include <stdint.h>
#include "hmr_records.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hmr_engine hmr_engine;
typedef struct hmr_resident_frame hmr_resident_frame;

#define HMR_MD5_MAX_JOBS 24      /* digests one engine keeps in flight (each holds a private copy of its picture) */
enum { HMR_OK = 0, HMR_PENDING = 1, HMR_ERR_CUDA = -1, HMR_ERR_ARG = -2, HMR_ERR_FORMAT = -3, HMR_ERR_NOMEM = -4, HMR_ERR_BUSY = -5 };

/* stage bits for hmr_set_stage_mask (default: all).  Same numbering as the oracle's orc_reconstruct_frame. */
enum { HMR_STAGE_MC = 1, HMR_STAGE_RESID = 2, HMR_STAGE_INTRA = 4, HMR_STAGE_DEBLOCK_V = 8, HMR_STAGE_DEBLOCK_H = 16, HMR_STAGE_SAO = 32, HMR_STAGE_ALL = 63 };
/* kernel indices for hmr_get_stage_times */
enum { HMR_T_H2D = 0, HMR_T_MC, HMR_T_RESID, HMR_T_INTRA, HMR_T_DEBLOCK_V, HMR_T_DEBLOCK_H, HMR_T_SAO, HMR_T_COUNT };

const char* hmr_version(void);

/* device: CUDA ordinal.  The picture geometry is taken from the first submitted frame header. */
int  hmr_engine_create(hmr_engine** out, int device);
void hmr_engine_destroy(hmr_engine* e);
const char* hmr_error_string(const hmr_engine* e);

/* Reconstruct one picture into DPB slot hdr->out_slot.  The records are copied (pinned staging -> one async H2D copy)
 * before the call returns; the kernels run asynchronously on the engine's stream, in submission order. */
int  hmr_submit_frame(hmr_engine* e, const hmr_frame_desc* frame);
/* Wait until everything submitted so far has finished. */
int  hmr_sync(hmr_engine* e);

/* Copy one component of a DPB slot to host memory (synchronises the engine's stream first).
 * dst_stride in samples; the plane is width x height of the component, int16 per sample. */
int  hmr_read_plane(hmr_engine* e, int slot, int comp, int16_t* dst, size_t dst_stride);
/* Same for the engine's working picture (the picture before SAO: state after the last executed stage <= deblock). */
int  hmr_read_work_plane(hmr_engine* e, int comp, int16_t* dst, size_t dst_stride);
/* Asynchronous variant into caller-provided PINNED memory, ordered after the frames submitted so far. */
int  hmr_read_plane_async(hmr_engine* e, int slot, int comp, int16_t* dst_pinned, size_t dst_stride);
/* Upload a plane into a DPB slot (tests: seeding reference pictures). */
int  hmr_write_plane(hmr_engine* e, int slot, int comp, const int16_t* src, size_t src_stride, int width, int height);

/* type: 2 = CRC, 3 = checksum (SEI decoded picture hash methods, SEI.h:118-134); out[3] one value per component. */
int  hmr_picture_hash(hmr_engine* e, int slot, int type, uint32_t out[3]);

/* Asynchronous SEI-MD5 of the picture in `slot` as it is after everything submitted so far.  The engine hashes a private
 * copy on a side stream; at most HMR_MD5_MAX_JOBS jobs may be outstanding (HMR_ERR_BUSY: collect results first).  A 2160p picture's
 * chain takes ~0.13 s whatever the load, and a parser delivers a group of B pictures in less than that: the ring has to hold a burst.
 * hmr_md5_result: out = 3 x 16 digest bytes (Y, Cb, Cr); wait = 0 polls (HMR_PENDING while running). */
int  hmr_md5_submit(hmr_engine* e, int slot, uint64_t* job);
int  hmr_md5_result(hmr_engine* e, uint64_t job, uint8_t out[48], int wait);

/* The picture in `slot` in TAppDecoder's `-o` wire format (replaces TVideoIOYuv::write, Lib/TLibVideoIO/TVideoIOYuv.cpp:706-790):
 * cropped by crop[4] = {left, right, top, bottom} luma samples (conformance + default display window), converted to
 * out_bit_depth[2] = {luma, chroma} (0 = internal; rounding shift + clip when smaller), planar Y, Cb, Cr, one byte per sample
 * when both output depths are <= 8, else two bytes little endian.  Packing runs on the device, only the packed bytes cross
 * PCIe.  *bytes = size of the packed picture; dst == NULL only queries it.  Synchronous. */
int  hmr_read_packed(hmr_engine* e, int slot, const int out_bit_depth[2], const int crop[4], void* dst, size_t capacity, size_t* bytes);

/* Page-lock / unlock caller memory so that hmr_read_plane_async can DMA straight into it (e.g. HM's TComPicYuv planes). */
int  hmr_host_register(void* p, size_t bytes);
int  hmr_host_unregister(void* p);
/* Stream markers: record = "everything enqueued so far"; wait blocks the host until that point has been reached. */
int  hmr_marker_record(hmr_engine* e, uint64_t* id);
int  hmr_marker_wait(hmr_engine* e, uint64_t id);

/* How much of a submitted / uploaded frame is checked before any kernel sees it.  Level 1 (default): header, geometry, array
 * presence, n_ctu / n_mc_tiles / tile-prefix consistency, every CTU's intra range inside n_intra — O(CTUs).  Level 2: in
 * addition every TU, intra and PU record (coefficient / residual offsets inside n_coef, block inside the picture, tile
 * prefix matching the PU sizes) — O(records); for callers that feed records they did not produce themselves (dump files).  A frame that fails
 * is rejected with HMR_ERR_FORMAT and nothing is launched. */
int  hmr_set_validation(hmr_engine* e, int level);

/* ---- measurement / test hooks ---- */
int  hmr_set_stage_mask(hmr_engine* e, int mask);
/* Enable CUDA-event timing of every stage of every submitted frame (adds events only, no syncs). */
int  hmr_enable_timing(hmr_engine* e, int on);
/* Milliseconds per stage accumulated since the last call (synchronises); also returns the number of frames and of
 * kernel launches they contain. */
int  hmr_get_stage_times(hmr_engine* e, float ms[HMR_T_COUNT], uint32_t* n_frames, uint32_t* n_launches);
/* Records resident in device memory: upload once, replay many times (bench `value`: inputs already in HBM). */
int  hmr_upload_frame(hmr_engine* e, const hmr_frame_desc* frame, hmr_resident_frame** out);
int  hmr_run_resident(hmr_engine* e, const hmr_resident_frame* f);
void hmr_free_resident(hmr_engine* e, hmr_resident_frame* f);
/* Enqueue n resident pictures in order (one call instead of n: keeps the host out of the measured loop). */
int  hmr_run_resident_list(hmr_engine* e, hmr_resident_frame* const* frames, int n);
/* Device-side stopwatch over several engines (= several CUDA streams) of one GPU, CUDA events only:
 * begin records the start event on `master`'s stream; join makes `master` wait for everything queued so far on
 * `other`; end records the stop event on `master`, synchronises and returns the elapsed milliseconds. */
int  hmr_timer_begin(hmr_engine* master);
int  hmr_timer_join(hmr_engine* master, hmr_engine* other);
int  hmr_timer_end(hmr_engine* master, float* ms);
/* Allocate / free page-locked host memory (so that callers without a CUDA binding can stage pinned buffers). */
void* hmr_alloc_pinned(size_t bytes);
void  hmr_free_pinned(void* p);
/* Launches a dummy write of `bytes` to scratch device memory on the engine stream (L2 flush between timed iterations). */
int  hmr_flush_l2(hmr_engine* e, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* HMRECON_H */
